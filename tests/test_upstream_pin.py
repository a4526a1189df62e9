"""The hook that turns "parity unpinned" into "pinned" the day the engine is importable.

The arithmetic of the path lives in the third-party `chatterbox` package (imported at
services/tts/core/synthesizer.py:167), which is neither vendored nor installable in this environment, and the
reference has no golden vectors for it.  oracle/hift_ref.py therefore restates the published algorithm and is pinned
only against itself.  With `chatterbox` present, this file compares the oracle with upstream's own
HiFTGenerator on the same state dict and inputs, and can regenerate tests/golden/ FROM UPSTREAM:

    python -m pytest tests/test_upstream_pin.py -q                      # compare
    GONOVA_REGEN_GOLDEN=1 python -m pytest tests/test_upstream_pin.py   # + rewrite tests/golden/decode_*.npz from upstream

Without the package every test here is skipped (visible in the pytest summary as `s`)."""
import os

import numpy as np
import pytest
import torch

chatterbox = pytest.importorskip("chatterbox", reason="the engine the service wraps is not installed: parity stays unpinned")

from gonova_tts_b200 import random_state_dict  # noqa: E402
from oracle import hift_ref as R  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _upstream_hift():
    """Upstream's vocoder exactly as S3Token2Wav.__init__ builds it (chatterbox/models/s3gen/s3gen.py)."""
    from chatterbox.models.s3gen.f0_predictor import ConvRNNF0Predictor
    from chatterbox.models.s3gen.hifigan import HiFTGenerator

    return HiFTGenerator(
        sampling_rate=24000, upsample_rates=[8, 5, 3], upsample_kernel_sizes=[16, 11, 7],
        source_resblock_kernel_sizes=[7, 7, 11], source_resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]],
        f0_predictor=ConvRNNF0Predictor()).eval()


@pytest.mark.parametrize("corners", [False, True])
def test_oracle_equals_upstream_hiftgenerator(corners):
    sd = random_state_dict(0, corners)
    up = _upstream_hift()
    missing, unexpected = up.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected                    # the oracle's / product's key names ARE upstream's
    assert all("stft_window" in k or "window" in k for k in missing), missing
    mine = R.load_model(sd)
    for T in (8, 32, 116):
        mel = R.synthetic_mel(2, T, seed=1234 + T)
        s = R.synthetic_source(mine, mel, seed=4321 + T)
        with torch.inference_mode():
            want = up.decode(x=mel, s=s)
            got = mine.decode(mel, s)
            f0_up = up.f0_predictor(mel)
            f0_me = mine.f0_predictor(mel)
        np.testing.assert_allclose(got.numpy(), want.numpy(), atol=2e-6, rtol=0)
        np.testing.assert_allclose(f0_me.numpy(), f0_up.numpy().reshape(f0_me.shape), atol=1e-4, rtol=1e-5)


def test_trim_fade_equals_upstream():
    from chatterbox.models.s3gen.s3gen import S3Token2Wav

    up = S3Token2Wav.__new__(S3Token2Wav)
    torch.nn.Module.__init__(up)
    n_trim = 24000 // 50
    tf = torch.zeros(2 * n_trim)
    tf[n_trim:] = (torch.cos(torch.linspace(torch.pi, 0, n_trim)) + 1) / 2
    np.testing.assert_array_equal(R.trim_fade_window().numpy(), tf.numpy())


@pytest.mark.skipif(not os.environ.get("GONOVA_REGEN_GOLDEN"), reason="set GONOVA_REGEN_GOLDEN=1 to rewrite tests/golden/")
@pytest.mark.parametrize("corners", [False, True])
@pytest.mark.parametrize("T", [8, 32])
def test_regenerate_golden_from_upstream(T, corners):
    sd = random_state_dict(0, corners)
    up = _upstream_hift()
    up.load_state_dict(sd, strict=False)
    mine = R.load_model(sd)
    mel = R.synthetic_mel(2, T, seed=1234 + T)
    s = R.synthetic_source(mine, mel, seed=4321 + T)
    path = os.path.join(GOLD, f"decode_T{T}_{'corners' if corners else 'plain'}.npz")
    old = dict(np.load(path))
    with torch.inference_mode():
        old["wav"] = up.decode(x=mel, s=s).numpy()
        old["f0"] = up.f0_predictor(mel).numpy().reshape(old["f0"].shape) if "f0" in old else None
    np.savez_compressed(path, **{k: v for k, v in old.items() if v is not None}, source=np.array("upstream chatterbox"))


def test_flow_oracle_equals_upstream_estimator():
    """oracle/flow_ref.py against upstream's ConditionalDecoder(causal) + CausalConditionalCFM.solve_euler."""
    from chatterbox.models.s3gen.decoder import ConditionalDecoder
    from chatterbox.models.s3gen.flow_matching import CausalConditionalCFM
    from oracle import flow_ref as FR

    sd = FR.random_state_dict(0)
    est = ConditionalDecoder(in_channels=320, out_channels=80, causal=True, channels=[256], dropout=0.0, attention_head_dim=64,
                             n_blocks=4, num_mid_blocks=12, num_heads=8, act_fn="gelu").eval()
    missing, unexpected = est.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    mine = FR.load_estimator(sd)
    z, mu, mask, spks, cond = FR.synthetic_inputs(2, 40, seed=1)
    t = torch.full((2,), 0.37)
    with torch.inference_mode():
        want = est(z, mask, mu, t, spks, cond)
        got = mine(z, mask, mu, t, spks, cond)
    np.testing.assert_allclose(got.numpy(), want.numpy(), atol=2e-5, rtol=1e-4)
    assert CausalConditionalCFM is not None


def test_flow_front_oracle_equals_upstream_encoder():
    """oracle/flow_enc_ref.py (tokens -> mu) against upstream's UpsampleConformerEncoder as S3Token2Mel builds it."""
    from chatterbox.models.s3gen.transformer.upsample_encoder import UpsampleConformerEncoder
    from oracle import flow_enc_ref as ER

    sd = ER.random_state_dict(0)
    enc = UpsampleConformerEncoder(output_size=512, attention_heads=8, linear_units=2048, num_blocks=6, dropout_rate=0.1,
                                   positional_dropout_rate=0.1, attention_dropout_rate=0.1, normalize_before=True,
                                   input_layer="linear", pos_enc_layer_type="rel_pos_espnet", selfattention_layer_type="rel_selfattn",
                                   input_size=512, use_cnn_module=False, macaron_style=False).eval()
    esd = {k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}
    missing, unexpected = enc.load_state_dict(esd, strict=False)
    assert not unexpected, unexpected                    # the oracle's key names ARE upstream's
    assert not [k for k in missing if "pe" not in k], missing
    mine = ER.load_front(sd)
    x = torch.randn(1, 37, 512, generator=torch.Generator().manual_seed(3))
    with torch.inference_mode():
        want, _ = enc(x, torch.tensor([37]))
        got = mine.encoder(x)
    np.testing.assert_allclose(got.numpy(), want.numpy(), atol=5e-5, rtol=1e-4)
