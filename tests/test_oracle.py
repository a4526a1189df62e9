"""The CPU oracle checked against independent restatements and the committed known-answer vectors.

PARITY UNPINNED (see oracle/hift_ref.py): the reference has no vectors for this path, so the oracle is
anchored on closed-form identities (STFT/iSTFT as explicit DFT sums, convolutions as explicit loops),
on an fp64 re-run, and on tests/golden/*.npz produced by tests/golden/make_golden.py."""
import os

import numpy as np
import pytest
import torch

from oracle import hift_ref as R
from oracle import tail_ref as TR
from gonova_tts_b200.weights import random_state_dict

from conftest import snr_db

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def plain_model():
    return R.load_model(random_state_dict(0, False))


def test_hyperparameters_and_shapes(plain_model):
    m = plain_model
    assert m.scale == 480 and m.sampling_rate == 24000 and m.audio_limit == 0.99
    n_params = sum(p.numel() for p in m.parameters())
    assert n_params == 20_806_557 + sum(  # weight-norm keeps g and v: g adds one scalar per dim-0 slice
        p.numel() for n, p in m.named_parameters() if n.endswith("original0"))
    mel = R.synthetic_mel(1, 10)
    s = R.synthetic_source(m, mel)
    taps = {}
    with torch.inference_mode():
        wav = m.decode(mel, s, taps=taps)
    assert wav.shape == (1, 4800) and s.shape == (1, 1, 4800)
    assert taps["s_stft"].shape == (1, 18, 1201)
    assert taps["stage0"].shape == (1, 256, 80) and taps["stage1"].shape == (1, 128, 400)
    assert taps["stage2"].shape == (1, 64, 1201) and taps["conv_post"].shape == (1, 18, 1201)


def test_stft_matches_closed_form(plain_model):
    g = torch.Generator().manual_seed(3)
    s = torch.randn(1, 1, 960, generator=g) * 0.1
    got = plain_model.source_stft(s)[0].numpy()
    want = R.stft_direct(s[0, 0].numpy())
    assert got.shape == want.shape == (18, 241)
    np.testing.assert_allclose(got, want, atol=2e-6)


def test_istft_matches_closed_form(plain_model):
    g = torch.Generator().manual_seed(4)
    x = torch.randn(1, 18, 61, generator=g)
    mag = torch.exp(x[:, :9])
    ph = torch.sin(x[:, 9:])
    got = plain_model._istft(mag, ph)[0].numpy()
    want = R.istft_direct(mag[0].numpy(), ph[0].numpy())
    assert got.shape == want.shape == (240,)
    np.testing.assert_allclose(got, want, atol=5e-6)


def test_stft_istft_round_trip(plain_model):
    g = torch.Generator().manual_seed(5)
    s = torch.randn(1, 480, generator=g) * 0.1
    re, im = plain_model._stft(s)
    spec = torch.complex(re, im)
    back = torch.istft(spec, 16, 4, 16, window=plain_model.stft_window)
    np.testing.assert_allclose(back.numpy(), s.numpy(), atol=1e-6)


def test_conv_layers_match_explicit_loops():
    # dilated Conv1d and ConvTranspose1d exactly as the decoder configures them, against index loops
    g = torch.Generator().manual_seed(6)
    x = torch.randn(1, 3, 11, generator=g, dtype=torch.float64)
    w = torch.randn(4, 3, 3, generator=g, dtype=torch.float64)
    d = 3
    pad = R.get_padding(3, d)
    got = torch.nn.functional.conv1d(x, w, padding=pad, dilation=d)[0].numpy()
    want = np.zeros((4, 11))
    for co in range(4):
        for t in range(11):
            for ci in range(3):
                for k in range(3):
                    q = t - pad + k * d
                    if 0 <= q < 11:
                        want[co, t] += x[0, ci, q].item() * w[co, ci, k].item()
    np.testing.assert_allclose(got, want, atol=1e-12)
    wt = torch.randn(3, 2, 7, generator=g, dtype=torch.float64)       # ups[2]: k7 s3 p2
    got = torch.nn.functional.conv_transpose1d(x, wt, stride=3, padding=2)[0].numpy()
    want = np.zeros((2, 33))
    for ci in range(3):
        for q in range(11):
            for k in range(7):
                p = q * 3 - 2 + k
                if 0 <= p < 33:
                    want[:, p] += x[0, ci, q].item() * wt[ci, :, k].numpy()
    np.testing.assert_allclose(got, want, atol=1e-12)


def test_snake_and_trim_fade():
    sn = R.Snake(2)
    with torch.no_grad():
        sn.alpha.copy_(torch.tensor([0.5, 2.0]))
    x = torch.tensor([[[0.3, -1.2], [0.7, 2.0]]])
    want = x + (1.0 / (sn.alpha.view(1, 2, 1) + 1e-9)) * torch.sin(sn.alpha.view(1, 2, 1) * x) ** 2
    np.testing.assert_allclose(sn(x).detach().numpy(), want.detach().numpy(), rtol=1e-6)
    w = R.trim_fade_window()
    assert w.shape == (960,) and float(w[:480].abs().max()) == 0.0
    assert abs(float(w[480])) < 1e-7 and abs(float(w[-1]) - 1.0) < 1e-7
    np.testing.assert_array_equal(w.numpy(), TR.trim_fade_window())


def test_sinegen_phase_is_running_sum():
    sg = R.SineGen(24000, 8, 0.1, 0.003, 10.0)
    f0 = torch.full((1, 1, 480), 120.0)
    ph = torch.zeros(1, 9, 1)
    noise = torch.zeros(1, 9, 480)
    sine, uv = sg(f0, phase_vec=ph, noise=noise)
    n = np.arange(1, 481)
    for h in range(9):
        want = 0.1 * np.sin(2 * np.pi * ((120.0 * (h + 1) / 24000.0 * n) % 1.0))
        np.testing.assert_allclose(sine[0, h].numpy(), want, atol=2e-4)
    assert float(uv.min()) == 1.0
    sine0, uv0 = sg(torch.full((1, 1, 8), 5.0), phase_vec=ph, noise=torch.ones(1, 9, 8))
    assert float(uv0.max()) == 0.0
    np.testing.assert_allclose(sine0.numpy(), 0.1 / 3, rtol=1e-6)     # unvoiced: noise_amp * noise only


def test_cache_source_overwrites_the_stochastic_source(plain_model):
    mel = R.synthetic_mel(1, 8)
    s = R.synthetic_source(plain_model, mel, seed=9)
    wav_a, s_a = plain_model.inference(mel, cache_source=s, generator=torch.Generator().manual_seed(1))
    wav_b, s_b = plain_model.inference(mel, cache_source=s, generator=torch.Generator().manual_seed(2))
    assert torch.equal(s_a, s) and torch.equal(wav_a, wav_b)
    with torch.inference_mode():
        assert torch.equal(wav_a, plain_model.decode(mel, s))


@pytest.mark.parametrize("corners", [False, True])
@pytest.mark.parametrize("T", [8, 32])
def test_golden_vectors(T, corners):
    z = np.load(os.path.join(GOLD, f"decode_T{T}_{'corners' if corners else 'plain'}.npz"))
    m = R.load_model(random_state_dict(0, corners))
    mel = R.synthetic_mel(2, T, seed=1234 + T)
    s = R.synthetic_source(m, mel, seed=4321 + T)
    np.testing.assert_allclose(s.numpy()[:, 0, :64], z["s_head"], atol=1e-6)
    taps = {}
    with torch.inference_mode():
        wav = m.decode(mel, s, taps=taps)
        f0 = m.f0_predictor(mel)
    # thread count / oneDNN blocking may reorder fp32 sums; the vectors are pinned to ~1e-5
    np.testing.assert_allclose(f0.numpy(), z["f0"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(taps["stage0"].numpy().mean(axis=2), z["stage0_mean"], atol=1e-5)
    assert snr_db(taps["conv_post"].numpy()[:, :, ::7], z["conv_post"]) > 80.0
    assert snr_db(wav.numpy(), z["wav"]) > (60.0 if corners else 80.0)
    if corners:
        assert float(np.abs(z["wav"]).max()) == pytest.approx(0.99)       # clamp fires
        assert float(z["f0"].max()) > 10.0 and float(z["f0"].min()) < 10.0   # voiced and unvoiced


def test_fp32_oracle_agrees_with_fp64(plain_model):
    sd = random_state_dict(0, False)
    m64 = R.load_model(sd, dtype=torch.float64)
    mel = R.synthetic_mel(1, 16)
    s = R.synthetic_source(plain_model, mel)
    with torch.inference_mode():
        a = plain_model.decode(mel, s)
        b = m64.decode(mel.double(), s.double())
    assert float((a.double() - b).abs().max()) < 1e-5
    assert snr_db(a.numpy(), b.numpy()) > 100.0
