"""The CFM flow decoder on a B200 against its CPU oracle (oracle/flow_ref.py; parity unpinned like the
vocoder's: the engine is not vendored).  Both sides hold the same seeded weights and get the SAME initial noise z."""
import numpy as np
import pytest
import torch

from conftest import snr_db
from oracle import flow_ref as FR

pytestmark = pytest.mark.gpu

# (max-abs, SNR dB) against the fp32 oracle, at about 2x what the kernels deliver (measured on B200: one estimator call
# tf32 4.6e-3 / 53.9 dB of the velocity, bf16 3.1e-2 / 36.6 dB; ten Euler steps tf32 <= 4.1e-3 / >= 61.3 dB of the mel, bf16
# <= 3.0e-2 / >= 42.1 dB).  Operands are rounded to 10 (tf32) / 7 (bf16) mantissa bits through ~270 GEMMs per estimator call.
TOL = {"tf32": (1e-2, 48.0), "bf16": (6e-2, 31.0)}
TOL10 = {"tf32": (1e-2, 55.0), "bf16": (6e-2, 36.0)}


@pytest.fixture(scope="module")
def sd():
    return FR.random_state_dict(0)


@pytest.fixture(scope="module")
def oracle(sd):
    return FR.load_estimator(sd)


@pytest.fixture(scope="module")
def flows(lib, cuda_device, sd):
    from gonova_tts_b200 import B200Flow

    cache = {}

    def get(dtype):
        if dtype not in cache:
            cache[dtype] = B200Flow(sd, device=cuda_device, dtype=dtype)
        return cache[dtype]

    return get


@pytest.mark.parametrize("dtype", ["tf32", "bf16"])
def test_one_euler_step_is_one_estimator_call(flows, oracle, cuda_device, dtype):
    """n_timesteps = 1: mel = z + 1 * ((1 + cfg) v_cond - cfg v_uncond) at t = 0: one estimator call on the doubled batch."""
    B, T = 2, 37
    z, mu, mask, spks, cond = FR.synthetic_inputs(B, T, seed=3)
    want = FR.solve_euler(oracle, z, mu, mask, spks, cond, n_timesteps=1).numpy()
    dev = cuda_device
    got = flows(dtype).decode(z.to(dev), mu.to(dev), spks.to(dev), cond.to(dev), n_timesteps=1).cpu().numpy()
    err, snr = np.abs(got - want).max(), snr_db(got - z.numpy(), want - z.numpy())
    print(f"[parity] flow 1 step {dtype}: max-abs {err:.3e}  SNR of the velocity {snr:.1f} dB")
    assert got.shape == (B, 80, T)
    assert err <= TOL[dtype][0] and snr >= TOL[dtype][1], (err, snr)


@pytest.mark.parametrize("dtype", ["tf32", "bf16"])
def test_ten_steps_with_guidance_and_ragged_lengths(flows, oracle, cuda_device, dtype):
    B, T = 3, 64
    lengths = [64, 41, 17]
    z, mu, mask, spks, cond = FR.synthetic_inputs(B, T, seed=5, lengths=lengths)
    want = FR.solve_euler(oracle, z * mask, mu, mask, spks, cond).numpy()
    dev = cuda_device
    flow = flows(dtype)
    got = flow.decode(z.to(dev), mu.to(dev), spks.to(dev), cond.to(dev), lengths=lengths).cpu().numpy()
    for b, n in enumerate(lengths):
        assert np.all(got[b, :, n:] == 0)                                  # masked frames come back as zeros
        err, snr = np.abs(got[b, :, :n] - want[b, :, :n]).max(), snr_db(got[b, :, :n], want[b, :, :n])
        print(f"[parity] flow 10 steps {dtype} row {b} ({n} frames): max-abs {err:.3e}  SNR {snr:.1f} dB")
        assert err <= TOL10[dtype][0] and snr >= TOL10[dtype][1], (b, err, snr)
    # every row equals that utterance decoded alone (the ragged batch is masking, not approximation)
    alone = flow.decode(z[1:2, :, :41].to(dev).contiguous(), mu[1:2, :, :41].to(dev).contiguous(), spks[1:2].to(dev),
                        cond[1:2, :, :41].to(dev).contiguous()).cpu().numpy()
    assert snr_db(got[1:2, :, :41], alone) >= 50.0


def test_upstream_call_shape_and_determinism(flows, cuda_device):
    """`decoder(mu=..., mask=..., spks=..., cond=..., n_timesteps=10)` -> (mel, None), noise from the module's fixed buffer."""
    flow = flows("bf16")
    B, T = 1, 50
    _, mu, mask, spks, cond = FR.synthetic_inputs(B, T, seed=9)
    dev = cuda_device
    a, none = flow(mu=mu.to(dev), mask=mask.to(dev), spks=spks.to(dev), cond=cond.to(dev), n_timesteps=10)
    b, _ = flow(mu=mu.to(dev), mask=mask.to(dev), spks=spks.to(dev), cond=cond.to(dev), n_timesteps=10)
    assert none is None and a.shape == (B, 80, T) and torch.isfinite(a).all() and torch.equal(a, b)
    assert flow.launches(10) > 1000
    with pytest.raises(ValueError, match="shape"):
        flow.decode(torch.zeros(1, 80, 7, device=dev), mu.to(dev), spks.to(dev), cond.to(dev))


def test_large_batch_walks_every_tile_path(flows, oracle, cuda_device):
    """2 B T = 19 000 rows: 75 pair tiles of flow_blk_kernel on 74 CTA pairs (one pair takes a second tile: accumulator
    hand-over, scratch reuse), six q/k/v tiles per resident A tile, attention over four 128-key tiles and two 256-query
    groups, ragged lengths across tile boundaries.  One Euler step against the oracle, row by row."""
    B, T = 19, 500
    lengths = [500, 500, 257, 256, 255, 3, 500, 129, 400, 500, 500, 13, 500, 385, 500, 500, 64, 500, 500]
    z, mu, mask, spks, cond = FR.synthetic_inputs(B, T, seed=11, lengths=lengths)
    want = FR.solve_euler(oracle, z * mask, mu, mask, spks, cond, n_timesteps=1).numpy()
    dev = cuda_device
    got = flows("bf16").decode(z.to(dev), mu.to(dev), spks.to(dev), cond.to(dev), lengths=lengths, n_timesteps=1).cpu().numpy()
    zn = (z * mask).numpy()
    worst = 1e9
    for b, n in enumerate(lengths):
        assert np.all(got[b, :, n:] == 0)
        snr = snr_db(got[b, :, :n] - zn[b, :, :n], want[b, :, :n] - zn[b, :, :n])
        worst = min(worst, snr)
        assert np.abs(got[b, :, :n] - want[b, :, :n]).max() <= TOL["bf16"][0] and snr >= TOL["bf16"][1], (b, n, snr)
    print(f"[parity] flow 1 step bf16, 19 x 500 ragged: worst row SNR of the velocity {worst:.1f} dB")


def test_empty_utterance_in_a_batch(flows, cuda_device):
    """lengths[b] = 0: that row comes back as zeros (the attention kernel has no key tile to visit) and its neighbours are
    what they are without it."""
    B, T = 3, 40
    z, mu, mask, spks, cond = FR.synthetic_inputs(B, T, seed=13, lengths=[40, 40, 40])
    dev = cuda_device
    flow = flows("bf16")
    full = flow.decode(z.to(dev), mu.to(dev), spks.to(dev), cond.to(dev), lengths=[40, 40, 40], n_timesteps=2).cpu().numpy()
    got = flow.decode(z.to(dev), mu.to(dev), spks.to(dev), cond.to(dev), lengths=[40, 0, 40], n_timesteps=2).cpu().numpy()
    assert np.all(got[1] == 0) and np.isfinite(got).all()
    assert np.array_equal(got[0], full[0]) and np.array_equal(got[2], full[2])


def test_fused_blocks_equal_one_launch_per_projection(flows, cuda_device, tmp_path):
    """GONOVA_FLOW_FUSED=0 (read once per process, hence a child process) runs the bf16 estimator with one conv_tc2 launch per
    projection, separate LayerNorm launches and the mma.sync attention: the same arithmetic up to bf16 rounding of different
    intermediates (the fused path never rounds the feed-forward's hidden activation through HBM differently: both round it to
    bf16 once).  The two must agree far inside the tolerance against the oracle."""
    import os
    import subprocess
    import sys

    B, T = 2, 100
    z, mu, mask, spks, cond = FR.synthetic_inputs(B, T, seed=17)
    dev = cuda_device
    got = flows("bf16").decode(z.to(dev), mu.to(dev), spks.to(dev), cond.to(dev), n_timesteps=3).cpu().numpy()
    out = tmp_path / "unfused.npy"
    code = (
        "import sys, numpy as np, torch\n"
        f"sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})\n"
        "from oracle import flow_ref as FR\n"
        "from gonova_tts_b200 import B200Flow\n"
        f"z, mu, mask, spks, cond = FR.synthetic_inputs({B}, {T}, seed=17)\n"
        "f = B200Flow(FR.random_state_dict(0), device='cuda:0', dtype='bf16')\n"
        "d = 'cuda:0'\n"
        "y = f.decode(z.to(d), mu.to(d), spks.to(d), cond.to(d), n_timesteps=3).cpu().numpy()\n"
        f"np.save({str(out)!r}, y)\n"
    )
    # each switch takes one fusion away: everything / the chained out-projection + feed-forward / its q/k/v tail and the fused
    # ResNet blocks / the tcgen05 attention (back to mma.sync)
    for name, extra in (("one launch per projection", {"GONOVA_FLOW_FUSED": "0"}),
                        ("out-proj and feed-forward as two launches", {"GONOVA_FLOW_OUTFF": "0"}),
                        ("no q/k/v tail, ResNet convs unfused", {"GONOVA_FLOW_QKV_TAIL": "0", "GONOVA_FLOW_RESNET_FUSED": "0"}),
                        ("mma.sync attention", {"GONOVA_FLOW_ATTN_TC": "0"})):
        env = dict(os.environ, **extra)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        ref = np.load(out)
        snr = snr_db(got, ref)
        print(f"[parity] flow 3 steps bf16: default against '{name}': max-abs {np.abs(got - ref).max():.3e}  SNR {snr:.1f} dB")
        assert snr >= 38.0, name


@pytest.mark.parametrize("B,T,lengths", [(1, 3, None), (3, 129, [129, 1, 128]), (2, 1031, [1031, 700])])
def test_odd_and_long_shapes(flows, oracle, cuda_device, B, T, lengths):
    """Three frames; lengths 1 / 128 / 129 around the 128-key attention tile; 20 s utterances (nine key tiles, five 256-row
    tiles per utterance, T not a multiple of 8: the transposed-V pitch is padded)."""
    z, mu, mask, spks, cond = FR.synthetic_inputs(B, T, seed=21, lengths=lengths)
    want = FR.solve_euler(oracle, z * mask, mu, mask, spks, cond, n_timesteps=1).numpy()
    dev = cuda_device
    got = flows("bf16").decode(z.to(dev), mu.to(dev), spks.to(dev), cond.to(dev), lengths=lengths, n_timesteps=1).cpu().numpy()
    zn = (z * mask).numpy()
    snr = snr_db(got - zn, want - zn)
    print(f"[parity] flow 1 step bf16 B={B} T={T} lengths={lengths}: max-abs {np.abs(got - want).max():.3e}  SNR of the velocity {snr:.1f} dB")
    assert np.isfinite(got).all() and np.abs(got - want).max() <= TOL["bf16"][0] and snr >= TOL["bf16"][1]
