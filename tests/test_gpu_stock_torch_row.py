"""The "stock PyTorch on the same B200" comparison row (last row of the table).

The reference service runs the decoder as ordinary PyTorch modules on the GPU with cuDNN autotuning and
TF32 switched on (services/tts/core/synthesizer.py:175-179).  The oracle restates exactly those modules, so moving
it to cuda:0 with the same switches IS that path: cuDNN convs, cuFFT STFT/iSTFT, one elementwise launch per
Snake / add.  It gives (1) a second, GPU-side parity witness at sizes the CPU oracle cannot reach in seconds,
and (2) the one pre-existing "Blackwell path" timed beside ours on the same device, same inputs.

The oracle is only the checker here (tests may import it); nothing in the product does."""
import numpy as np
import pytest
import torch

from conftest import snr_db
from oracle import hift_ref as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def stock(cuda_device):
    from gonova_tts_b200 import random_state_dict

    sd = random_state_dict(0, False)
    old = (torch.backends.cudnn.benchmark, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cudnn.benchmark = True
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    yield sd, R.load_model(sd).to(cuda_device).eval()
    torch.backends.cudnn.benchmark, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def _time_ms(fn, warmup, iters):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


@pytest.mark.parametrize("dtype", ["tf32", "bf16"])
def test_full_size_batch_against_stock_torch_on_the_gpu(lib, cuda_device, stock, dtype):
    """8 x 10 s utterances (the CPU oracle needs minutes for this): every sample of our decode against the stock
    PyTorch TF32 decode of the same mel and source.  Both sides round their operands (TF32 on both for the tf32
    path), so the bounds are the tensor-core bounds of test_gpu_decode.TOL, not the fp32 ones."""
    from gonova_tts_b200 import B200HiFT

    sd, m = stock
    B, T = 8, 500
    mel = R.synthetic_mel(B, T, seed=77).to(cuda_device)
    g = torch.Generator().manual_seed(78)
    s = (torch.rand(B, 1, T * 480, generator=g) * 0.2 - 0.1).to(cuda_device)
    with torch.inference_mode():
        want = m.decode(mel, s).cpu().numpy()
    got = B200HiFT(sd, device=cuda_device, dtype=dtype).decode(mel, s).cpu().numpy()
    err, snr = np.abs(got - want).max(), snr_db(got, want)
    print(f"[parity] vs stock torch TF32 on cuda, B={B} T={T}, {dtype}: max-abs {err:.3e}  SNR {snr:.1f} dB")
    bound = {"tf32": (2e-4, 64.0), "bf16": (1.5e-3, 45.0)}[dtype]      # measured 5.5e-5 / 70.8 dB and 4.8e-4 / 51.3 dB
    assert err <= bound[0] and snr >= bound[1], (dtype, err, snr)


def test_throughput_beside_stock_torch(lib, cuda_device, stock):
    """BASELINE config 3 (64 x T=500): audio-seconds per second of the stock PyTorch path and of ours, same device,
    same resident inputs, CUDA events.  Reported (printed and asserted loosely: the hand-written path must at least
    be several times faster, otherwise something fell back)."""
    from gonova_tts_b200 import B200HiFT

    sd, m = stock
    B, T = 64, 500
    mel = R.synthetic_mel(B, T, seed=5).to(cuda_device)
    g = torch.Generator().manual_seed(6)
    s = (torch.rand(B, 1, T * 480, generator=g) * 0.2 - 0.1).to(cuda_device)
    audio_s = B * T / 50.0

    def stock_step():
        with torch.inference_mode():
            m.decode(mel, s)

    ms_stock = _time_ms(stock_step, 2, 3)
    torch.cuda.empty_cache()
    rows = [f"[stock-torch row] B={B} T={T} decode(x, s):  stock PyTorch TF32 (cuDNN/cuFFT) {ms_stock:8.2f} ms"
            f" = {audio_s / ms_stock * 1e3:8.0f} audio-s/s"]
    ours = {}
    for dtype in ("tf32", "bf16"):
        dec = B200HiFT(sd, device=cuda_device, dtype=dtype)
        wav = torch.empty(B, T * 480, dtype=torch.float32, device=cuda_device)
        ours[dtype] = _time_ms(lambda: dec.decode(mel, s, out=wav), 3, 5)
        rows.append(f"[stock-torch row] B={B} T={T} decode(x, s):  gonova_tts_b200 {dtype:5s}              "
                    f"{ours[dtype]:8.2f} ms = {audio_s / ours[dtype] * 1e3:8.0f} audio-s/s  "
                    f"({ms_stock / ours[dtype]:.1f}x)")
        del dec
    print("\n" + "\n".join(rows))
    assert ours["bf16"] * 3 < ms_stock and ours["tf32"] * 2 < ms_stock, rows
