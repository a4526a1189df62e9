"""CPU checks of the flow front's oracle (oracle/flow_enc_ref.py) and of the host-side tables that mirror it."""
import math

import torch

from oracle import flow_enc_ref as ER


def test_rel_shift_matches_its_definition():
    T = 7
    x = torch.randn(2, 3, T, 2 * T - 1)
    y = ER.RelPositionMultiHeadedAttention.rel_shift(x)
    assert y.shape == (2, 3, T, T)
    for i in range(T):
        for j in range(T):
            assert torch.equal(y[:, :, i, j], x[:, :, i, T - 1 - i + j])


def test_position_table_layout():
    """Row r of pos_emb is the sinusoid of relative position (T-1) - r: what the CUDA table kernel assumes."""
    T, d = 9, 512
    pe = ER.EspnetRelPositionalEncoding(d).position_encoding(T)[0]
    assert pe.shape == (2 * T - 1, d)
    div = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * -(math.log(10000.0) / d))
    for r in (0, 3, T - 1, T, 2 * T - 2):
        rel = float(T - 1 - r)
        assert torch.allclose(pe[r, 0::2], torch.sin(rel * div), atol=1e-6)
        assert torch.allclose(pe[r, 1::2], torch.cos(rel * div), atol=1e-6)


def test_upsample_is_two_causal_three_tap_convs():
    """The identity the CUDA path uses for Upsample1D: row 2m reads (W0+W1, W2+W3, W4), row 2m+1 (W0, W1+W2, W3+W4) over
    x[m-2], x[m-1], x[m]."""
    torch.manual_seed(0)
    up = ER.Upsample1D(8, 8, 2)
    x = torch.randn(1, 8, 11)
    want = up(x)
    w, b = up.conv.weight, up.conv.bias
    even = torch.stack([w[..., 0] + w[..., 1], w[..., 2] + w[..., 3], w[..., 4]], dim=-1)
    odd = torch.stack([w[..., 0], w[..., 1] + w[..., 2], w[..., 3] + w[..., 4]], dim=-1)
    xp = torch.nn.functional.pad(x, (2, 0))
    ye = torch.nn.functional.conv1d(xp, even, b)
    yo = torch.nn.functional.conv1d(xp, odd, b)
    got = torch.stack([ye, yo], dim=-1).reshape(1, 8, 22)
    assert torch.allclose(got, want, atol=1e-5)


def test_encode_shapes_masking_and_determinism():
    m = ER.make_front(0)
    tokens, token_len, emb = ER.synthetic_tokens(3, 12, seed=2, lengths=[12, 5, 0])
    with torch.inference_mode():
        mu = m.encode(tokens, token_len)
        spks = m.speaker(emb)
        again = m.encode(tokens, token_len)
    assert mu.shape == (3, 80, 24) and spks.shape == (3, 80)
    assert torch.equal(mu, again)
    assert not mu[1, :, 10:].any() and not mu[2].any()
    assert torch.isfinite(mu).all() and mu[0].abs().max() > 0.1
    # an utterance does not depend on its neighbours or on the padding after it
    with torch.inference_mode():
        alone = m.encode(tokens[1:2, :5])
    assert torch.equal(alone[0], mu[1, :, :10])


def test_shape_table_matches_the_oracle_state_dict():
    from gonova_tts_b200.flow_front import front_layer_shapes, random_front_state_dict

    sd = ER.random_state_dict(0)
    assert {k: tuple(v.shape) for k, v in sd.items()} == dict(front_layer_shapes())
    ER.load_front(random_front_state_dict(3))
