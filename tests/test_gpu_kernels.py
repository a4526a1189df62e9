"""Per-kernel parity on a B200, every call through the C ABI (libgonova_hift.so).

Integer / byte work (the PCM pack) is compared bit-exactly with the numpy definition in
oracle/tail_ref.py; floating-point kernels are compared with the CPU oracle (oracle/hift_ref.py) or
a plain fp32 torch restatement of the same op, with the tolerance written next to each assert."""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F
from hypothesis import given, settings, strategies as st

from conftest import snr_db
from oracle import hift_ref as R
from oracle import tail_ref as TR

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _vp(t):
    return None if t is None else C.c_void_p(t.data_ptr())


# ------------------------------------------------------------------------------------------------
# streaming tail: bit-exact
# ------------------------------------------------------------------------------------------------
def test_pcm_tail_known_answers_bit_exact(lib, cuda_device):
    from gonova_tts_b200 import pcm_tail, trim_fade_window

    z = np.load(os.path.join(GOLD, "tail_kat.npz"))
    x = torch.from_numpy(z["x"]).to(cuda_device).reshape(1, -1)
    i16, f32 = pcm_tail(x, None, None, limit=10.0, want_i16=True, want_f32=True)   # limit 10: pack rule alone
    np.testing.assert_array_equal(i16.cpu().numpy()[0], z["x_i16"])
    cur = torch.from_numpy(z["cur"]).to(cuda_device)
    prev = torch.from_numpy(z["prev"]).to(cuda_device)
    w = torch.from_numpy(z["w"]).to(cuda_device)
    i16, f32 = pcm_tail(cur, prev, w, 0.99, True, True)
    np.testing.assert_array_equal(f32.cpu().numpy(), z["f_cf"])
    np.testing.assert_array_equal(i16.cpu().numpy(), z["i_cf"])
    i16, f32 = pcm_tail(cur, None, trim_fade_window(cuda_device), 0.99, True, True)
    np.testing.assert_array_equal(f32.cpu().numpy(), z["f_tf"])
    np.testing.assert_array_equal(i16.cpu().numpy(), z["i_tf"])


@settings(max_examples=25, deadline=None)
@given(n=st.integers(1, 3000), rows=st.integers(1, 3), fade=st.sampled_from([0, 1, 7, 480]),
       has_prev=st.booleans(), seed=st.integers(0, 2 ** 31 - 1), scale=st.sampled_from([0.05, 0.5, 2.0]))
def test_pcm_tail_property_bit_exact(lib, cuda_device, n, rows, fade, has_prev, seed, scale):
    """Any (rows, n, fade): ragged sizes exercise the scalar path, multiples of 4 the 128-bit path."""
    from gonova_tts_b200 import pcm_tail

    rng = np.random.default_rng(seed)
    cur = (rng.standard_normal((rows, n)) * scale).astype(np.float32)
    w = TR.fade_window(fade) if fade else None
    prev = (rng.standard_normal((rows, fade)) * scale).astype(np.float32) if (has_prev and fade) else None
    want_f, want_i = TR.pcm_tail(cur, prev, w, 0.99)
    i16, f32 = pcm_tail(torch.from_numpy(cur).to(cuda_device),
                        None if prev is None else torch.from_numpy(prev).to(cuda_device),
                        None if w is None else torch.from_numpy(w).to(cuda_device), 0.99, True, True)
    np.testing.assert_array_equal(f32.cpu().numpy(), want_f)
    np.testing.assert_array_equal(i16.cpu().numpy(), want_i)


def test_pcm_tail_full_size_checksum(lib, cuda_device):
    """BASELINE config 3 size (64 x 240000 samples): int16 equals the numpy pack everywhere."""
    from gonova_tts_b200 import pcm_tail

    g = torch.Generator().manual_seed(11)
    cur = (torch.randn(64, 240000, generator=g) * 0.5)
    i16, _ = pcm_tail(cur.to(cuda_device), None, None, 0.99, True, False)
    _, want = TR.pcm_tail(cur.numpy(), None, None, 0.99)
    got = i16.cpu().numpy()
    assert got.shape == want.shape
    assert int((got != want).sum()) == 0
    assert int(got.astype(np.int64).sum()) == int(want.astype(np.int64).sum())


def test_pcm_tail_empty_and_strided_rows(lib, cuda_device):
    from gonova_tts_b200 import pcm_tail

    big = torch.randn(3, 1000, device=cuda_device)
    view = big[:, 100:612]                       # row stride 1000, offset not 16-byte aligned in general
    i16, f32 = pcm_tail(view, None, None, 0.99, True, True)
    wf, wi = TR.pcm_tail(view.cpu().numpy(), None, None, 0.99)
    np.testing.assert_array_equal(f32.cpu().numpy(), wf)
    np.testing.assert_array_equal(i16.cpu().numpy(), wi)
    e16, e32 = pcm_tail(torch.empty(2, 0, device=cuda_device), None, None, 0.99, True, True)
    assert e16.shape == (2, 0) and e32.shape == (2, 0)


# ------------------------------------------------------------------------------------------------
# STFT / iSTFT head
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,L", [(1, 16), (2, 480), (3, 4800), (1, 100 * 480)])
def test_stft_matches_oracle(lib, cuda_device, B, L):
    g = torch.Generator().manual_seed(L)
    s = torch.randn(B, L, generator=g) * 0.1
    m = R.HiFTGenerator()
    want = m.source_stft(s[:, None]).numpy()
    Fr = L // 4 + 1
    sd = s.to(cuda_device)
    out = torch.empty(B, 18, Fr, device=cuda_device)
    rc = lib.gnv_stft(_vp(sd), B, L, _vp(out), None)
    assert rc == 0
    # fp32 sums of 16 products of |x| ~ 0.1: 2e-6 absolute
    np.testing.assert_allclose(out.cpu().numpy(), want, atol=2e-6)


@pytest.mark.parametrize("B,Fr", [(1, 2), (2, 121), (2, 1201), (1, 12001)])
def test_istft_head_matches_oracle(lib, cuda_device, B, Fr):
    g = torch.Generator().manual_seed(Fr)
    x = torch.randn(B, 18, Fr, generator=g)
    x[:, :9] *= 2.0                                    # exp spans ~e^-6..e^6: the 100 clip fires
    m = R.HiFTGenerator()
    with torch.inference_mode():
        want = torch.clamp(m._istft(torch.exp(x[:, :9]), torch.sin(x[:, 9:])), -0.99, 0.99).numpy()
    xd = x.to(cuda_device)
    out = torch.empty(B, 4 * (Fr - 1), device=cuda_device)
    assert lib.gnv_istft(_vp(xd), B, Fr, C.c_float(0.99), _vp(out), None) == 0
    got = out.cpu().numpy()
    assert got.shape == want.shape
    # magnitudes reach 100, so absolute error scales with them: 1e-5 * 100 ~ 2e-4 before the clamp
    np.testing.assert_allclose(got, want, atol=2e-4)
    assert np.abs(got).max() <= np.float32(0.99)
    if Fr >= 121:
        assert (np.abs(want) >= 0.99).any(), "case should exercise the clamp"


def test_stft_then_istft_head_full_row(lib, cuda_device):
    """A full 10 s row (240000 samples, BASELINE config 3 length) through both kernels.  The head
    maps x -> exp(x[:9]) * exp(i*sin(x[9:])), so phases are confined to +-1 rad and STFT -> head is
    not an identity; the size-independent check is: device STFT equals the oracle STFT, and the
    device head equals the oracle head on log|STFT| / a bounded phase, at full length."""
    g = torch.Generator().manual_seed(99)
    B, L = 2, 240000
    s = (torch.rand(B, L, generator=g) * 2 - 1) * 0.5
    Fr = L // 4 + 1
    m = R.HiFTGenerator()
    spec = torch.empty(B, 18, Fr, device=cuda_device)
    assert lib.gnv_stft(_vp(s.to(cuda_device)), B, L, _vp(spec), None) == 0
    want_spec = m.source_stft(s[:, None])
    np.testing.assert_allclose(spec.cpu().numpy(), want_spec.numpy(), atol=5e-6)
    re, im = want_spec[:, :9], want_spec[:, 9:]
    x = torch.cat([torch.log(torch.sqrt(re * re + im * im).clamp_min(1e-6)), im], 1).contiguous()
    out = torch.empty(B, L, device=cuda_device)
    assert lib.gnv_istft(_vp(x.to(cuda_device)), B, Fr, C.c_float(10.0), _vp(out), None) == 0
    with torch.inference_mode():
        want = m._istft(torch.exp(x[:, :9]), torch.sin(x[:, 9:])).numpy()
    np.testing.assert_allclose(out.cpu().numpy(), want, atol=1e-5)


# ------------------------------------------------------------------------------------------------
# conv layers (every geometry the decoder uses) against torch fp32 on the CPU
# ------------------------------------------------------------------------------------------------
def _conv_case(lib, dev, dtype, flags, transposed, B, Cin, Cout, L, k, stride, pad, dil, act, seed, residual=False):
    from gonova_tts_b200 import _cabi

    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, Cin, L, generator=g)
    bound = 1.0 / np.sqrt(Cin * k)
    if transposed:
        w = (torch.rand(Cin, Cout, k, generator=g) * 2 - 1) * bound
    else:
        w = (torch.rand(Cout, Cin, k, generator=g) * 2 - 1) * bound
    b = (torch.rand(Cout, generator=g) * 2 - 1) * bound
    alpha = 0.5 + 1.5 * torch.rand(Cout, generator=g)
    if transposed:
        y = F.conv_transpose1d(x, w, b, stride=stride, padding=pad)
    else:
        y = F.conv1d(x, w, b, stride=stride, padding=pad, dilation=dil)
    res = None
    if residual:
        res = torch.randn(y.shape, generator=g)
        y = y + res
    if act == "snake":
        a = alpha.view(1, -1, 1)
        y = y + (1.0 / (a + 1e-9)) * torch.sin(y * a) ** 2
    elif act == "lrelu":
        y = F.leaky_relu(y, 0.1)
    elif act == "elu":
        y = F.elu(y)
    Lout = y.shape[2]
    xd = x.to(dev)
    resd = None if res is None else res.to(dev)
    out = torch.empty(B, Cout, Lout, device=dev)
    rc = lib.gnv_conv1d(0, _cabi.DTYPE[dtype], flags, int(transposed), _vp(xd), B, Cin, L,
                        C.c_void_p(w.data_ptr()), C.c_void_p(b.data_ptr()), Cout, k, stride, pad, dil,
                        _cabi.ACT[act], C.c_void_p(alpha.data_ptr()), C.c_float(0.1), _vp(resd), _vp(out), Lout, None)
    assert rc == 0, _cabi.last_error(None)
    return out.cpu().numpy(), y.numpy()


LAYER_GEOMS = [
    # name, transposed, Cin, Cout, L, k, stride, pad, dil, act
    ("conv_pre", False, 80, 512, 50, 7, 1, 3, 1, "lrelu"),
    ("f0_conv0", False, 80, 512, 37, 3, 1, 1, 1, "elu"),
    ("f0_conv1", False, 512, 512, 37, 3, 1, 1, 1, "elu"),
    ("rb256_k3_d5", False, 256, 256, 200, 3, 1, 5, 5, "snake"),
    ("rb256_k7_d3", False, 256, 256, 200, 7, 1, 9, 3, "snake"),
    ("rb128_k11_d5", False, 128, 128, 333, 11, 1, 25, 5, "snake"),
    ("rb64_k11_d1", False, 64, 64, 601, 11, 1, 5, 1, "none"),
    ("rb64_k3_d3", False, 64, 64, 129, 3, 1, 3, 3, "snake"),
    ("conv_post", False, 64, 18, 601, 7, 1, 3, 1, "none"),
    ("ups0", True, 512, 256, 25, 16, 8, 4, 1, "snake"),
    ("ups1", True, 256, 128, 100, 11, 5, 3, 1, "snake"),
    ("ups2", True, 128, 64, 200, 7, 3, 2, 1, "none"),
    ("sdown0", False, 18, 256, 1201, 30, 15, 7, 1, "snake"),
    ("sdown1", False, 18, 128, 1201, 6, 3, 1, 1, "none"),
    ("sdown2", False, 18, 64, 301, 1, 1, 0, 1, "snake"),
]


@pytest.mark.parametrize("geom", LAYER_GEOMS, ids=[g[0] for g in LAYER_GEOMS])
def test_conv_fp32_cuda_cores_matches_torch(lib, cuda_device, geom):
    name, tr, Cin, Cout, L, k, s, p, d, act = geom
    got, want = _conv_case(lib, cuda_device, "fp32", 0, tr, 2, Cin, Cout, L, k, s, p, d, act, seed=len(name),
                           residual=(act == "none"))
    # same fp32 products in a different summation order: K <= 5632 terms of magnitude <= ~1
    np.testing.assert_allclose(got, want, atol=3e-5, rtol=1e-5)


@pytest.mark.parametrize("geom", [g for g in LAYER_GEOMS if not g[0].startswith("sdown")],
                         ids=[g[0] for g in LAYER_GEOMS if not g[0].startswith("sdown")])
@pytest.mark.parametrize("dtype", ["tf32", "bf16"])
def test_conv_tcgen05_matches_torch(lib, cuda_device, geom, dtype):
    name, tr, Cin, Cout, L, k, s, p, d, act = geom
    got, want = _conv_case(lib, cuda_device, dtype, 0, tr, 2, Cin, Cout, L, k, s, p, d, act, seed=len(name) + 7,
                           residual=(act == "none"))
    err = np.abs(got - want).max()
    snr = snr_db(got, want)
    # operands rounded to 10 (tf32) / 7 (bf16) mantissa bits, fp32 accumulate; the activation output
    # is stored in the operand type of the next conv, so bf16 adds one more rounding of the result.
    # The rounding alone (operands and stored activation rounded on the CPU, fp32 accumulate) gives, over these
    # geometries: tf32 max-abs <= 1.8e-3 / SNR >= 68.5 dB, bf16 <= 1.43e-2 / >= 50.5 dB.  Bounds at < 2x that.
    print(f"[parity] layer {name} {dtype}: max-abs {err:.3e} SNR {snr:.1f} dB")
    if dtype == "tf32":
        assert err < 3e-3 and snr > 64.0, (err, snr)
    else:
        assert err < 2.5e-2 and snr > 46.0, (err, snr)


def test_conv_tcgen05_equals_cuda_core_kernel_on_same_operands(lib, cuda_device):
    """Same rounded operands through the tcgen05 kernel and the CUDA-core kernel: only the summation
    order differs, so the two agree to fp32 accumulation noise.  This isolates descriptor / TMA /
    TMEM addressing mistakes from rounding."""
    from gonova_tts_b200 import _cabi

    for dtype in ("tf32", "bf16"):
        a, _ = _conv_case(lib, cuda_device, dtype, 0, False, 3, 128, 128, 700, 7, 1, 9, 3, "none", seed=5)
        b, _ = _conv_case(lib, cuda_device, dtype, _cabi.FLAG_SIMT_CONV, False, 3, 128, 128, 700, 7, 1, 9, 3, "none",
                          seed=5)
        np.testing.assert_allclose(a, b, atol=2e-5, rtol=1e-5)
        a, _ = _conv_case(lib, cuda_device, dtype, 0, True, 2, 256, 128, 131, 11, 5, 3, 1, "none", seed=6)
        b, _ = _conv_case(lib, cuda_device, dtype, _cabi.FLAG_SIMT_CONV, True, 2, 256, 128, 131, 11, 5, 3, 1, "none",
                          seed=6)
        np.testing.assert_allclose(a, b, atol=2e-5, rtol=1e-5)


def test_conv_rejects_bad_geometry(lib, cuda_device):
    from gonova_tts_b200 import _cabi

    x = torch.zeros(1, 8, 16, device=cuda_device)
    w = torch.zeros(8, 8, 3)
    out = torch.zeros(1, 8, 16, device=cuda_device)
    rc = lib.gnv_conv1d(0, 0, 0, 0, _vp(x), 1, 8, 16, C.c_void_p(w.data_ptr()), None, 8, 3, 1, 1, 1, 0, None,
                        C.c_float(0), None, _vp(out), 15, None)
    assert rc != 0 and "Lout" in _cabi.last_error(None)


def test_mulaw_bit_exact(lib, cuda_device):
    from gonova_tts_b200 import mulaw_encode

    x = np.arange(-32768, 32768, dtype=np.int32).astype(np.int16)                  # every int16, 128-bit path
    got = mulaw_encode(torch.from_numpy(x).to(cuda_device)).cpu().numpy()
    np.testing.assert_array_equal(got, TR.mulaw_encode(x))
    rng = np.random.default_rng(3)
    for n in (1, 7, 8, 9, 4803, 240000):                                           # ragged tails
        y = rng.integers(-32768, 32768, size=n, dtype=np.int64).astype(np.int16)
        got = mulaw_encode(torch.from_numpy(y).to(cuda_device)).cpu().numpy()
        np.testing.assert_array_equal(got, TR.mulaw_encode(y))
    assert mulaw_encode(torch.empty(0, dtype=torch.int16, device=cuda_device)).numel() == 0
