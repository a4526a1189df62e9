"""CPU oracle of the step BEFORE the vocoder: the conditional-flow-matching decoder that turns the encoder's
`mu` into mel frames.  TEST INFRASTRUCTURE ONLY (tests/, smoke and bench's CPU legs may import it; the product may not).

PARITY UNPINNED, like oracle/hift_ref.py and for the same reason: the arithmetic lives in the third-party `chatterbox`
package (imported at services/tts/core/synthesizer.py:167; not vendored, not pinned, not installable here), reached from
`self.model.generate(...)` (synthesizer.py:344-350) through `S3Gen.inference -> flow_inference -> flow.decoder(...)`.
This file restates, from the published architecture of that engine's flow decoder (CosyVoice-2 lineage):

  CausalConditionalCFM.forward / solve_euler     n_timesteps Euler steps on a cosine time grid, classifier-free guidance
                                                 with a doubled batch (conditioned row + zeroed-condition row), cfg 0.7
  ConditionalDecoder(causal=True).forward        the estimator v(x, t | mu, spks, cond): a 1-D U-Net with ONE down
      in_channels 320 = x(80) + mu(80) + spks(80) + cond(80), channels [256], 4 transformer blocks per level, 12 mid blocks,
      8 heads x 64, GELU feed-forward; causal 3-tap convs, LayerNorm over channels, Mish
  SinusoidalPosEmb(320) + TimestepEmbedding(320 -> 1024 -> 1024, SiLU)

State-dict names follow upstream (`down_blocks.0.0.block1.block.0.weight`, `mid_blocks.3.1.2.attn1.to_q.weight`, ...) so a real
`flow.decoder.estimator.*` checkpoint loads unchanged the day it is available; tests/test_upstream_pin.py compares this
file with upstream's module when `chatterbox` is importable.  Scope: the estimator and the ODE loop — the token
embedding and the Conformer encoder in front of them (run once per utterance, ~5 % of the flow's FLOPs) are not restated."""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F
from torch import nn

MEL = 80
IN_CHANNELS = 4 * MEL          # x, mu, spks, cond
CHANNELS = 256
TIME_DIM = 4 * CHANNELS
N_BLOCKS = 4
N_MID = 12
HEADS = 8
HEAD_DIM = 64
FF_MULT = 4
CFG_RATE = 0.7
N_TIMESTEPS = 10


class SinusoidalPosEmb(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.dim = dim

    def forward(self, t: torch.Tensor, scale: float = 1000.0) -> torch.Tensor:
        half = self.dim // 2
        emb = math.log(10000) / (half - 1)
        emb = torch.exp(torch.arange(half, dtype=t.dtype, device=t.device) * -emb)
        emb = scale * t[:, None] * emb[None, :]
        return torch.cat([emb.sin(), emb.cos()], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels: int, time_embed_dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class CausalConv1d(nn.Conv1d):
    """Conv1d with all of its padding on the left: output frame t sees input frames t-k+1 .. t."""

    def __init__(self, cin: int, cout: int, k: int):
        super().__init__(cin, cout, k)
        self.left = k - 1

    def forward(self, x):
        return super().forward(F.pad(x, (self.left, 0)))


class _Transpose(nn.Module):
    def forward(self, x):
        return x.transpose(1, 2)


class CausalBlock1D(nn.Module):
    def __init__(self, dim: int, dim_out: int):
        super().__init__()
        self.block = nn.Sequential(CausalConv1d(dim, dim_out, 3), _Transpose(), nn.LayerNorm(dim_out), _Transpose(), nn.Mish())

    def forward(self, x, mask):
        return self.block(x * mask) * mask


class CausalResnetBlock1D(nn.Module):
    def __init__(self, dim: int, dim_out: int, time_emb_dim: int):
        super().__init__()
        self.mlp = nn.Sequential(nn.Mish(), nn.Linear(time_emb_dim, dim_out))
        self.block1 = CausalBlock1D(dim, dim_out)
        self.block2 = CausalBlock1D(dim_out, dim_out)
        self.res_conv = nn.Conv1d(dim, dim_out, 1)

    def forward(self, x, mask, time_emb):
        h = self.block1(x, mask)
        h = h + self.mlp(time_emb).unsqueeze(-1)
        h = self.block2(h, mask)
        return h + self.res_conv(x * mask)


class Attention(nn.Module):
    """Self-attention of the diffusers kind the upstream block uses: q/k/v without bias, inner dim heads x head_dim, output
    projection with bias; additive key mask."""

    def __init__(self, dim: int, heads: int, dim_head: int):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.scale = heads, dim_head ** -0.5
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_k = nn.Linear(dim, inner, bias=False)
        self.to_v = nn.Linear(dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, dim), nn.Identity()])

    def forward(self, x, key_bias):
        B, T, _ = x.shape
        q = self.to_q(x).view(B, T, self.heads, -1).transpose(1, 2)
        k = self.to_k(x).view(B, T, self.heads, -1).transpose(1, 2)
        v = self.to_v(x).view(B, T, self.heads, -1).transpose(1, 2)
        s = torch.matmul(q, k.transpose(-1, -2)) * self.scale + key_bias[:, None, None, :]
        o = torch.matmul(torch.softmax(s, dim=-1), v)
        return self.to_out[0](o.transpose(1, 2).reshape(B, T, -1))


class _GELU(nn.Module):
    def __init__(self, dim: int, inner: int):
        super().__init__()
        self.proj = nn.Linear(dim, inner)

    def forward(self, x):
        return F.gelu(self.proj(x))


class FeedForward(nn.Module):
    def __init__(self, dim: int, mult: int = FF_MULT):
        super().__init__()
        self.net = nn.ModuleList([_GELU(dim, dim * mult), nn.Identity(), nn.Linear(dim * mult, dim)])

    def forward(self, x):
        return self.net[2](self.net[0](x))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, heads: int, dim_head: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, heads, dim_head)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, key_bias):
        x = self.attn1(self.norm1(x), key_bias) + x
        return self.ff(self.norm3(x)) + x


class ConditionalDecoder(nn.Module):
    """v = estimator(x, mask, mu, t, spks, cond): all of [B, 80, T] except t [B], spks [B, 80], mask [B, 1, T]."""

    def __init__(self):
        super().__init__()
        self.time_embeddings = SinusoidalPosEmb(IN_CHANNELS)
        self.time_mlp = TimestepEmbedding(IN_CHANNELS, TIME_DIM)

        def level(cin):
            return nn.ModuleList([CausalResnetBlock1D(cin, CHANNELS, TIME_DIM),
                                  nn.ModuleList([BasicTransformerBlock(CHANNELS, HEADS, HEAD_DIM) for _ in range(N_BLOCKS)])])

        down = level(IN_CHANNELS)
        down.append(CausalConv1d(CHANNELS, CHANNELS, 3))                 # the last (only) level does not downsample
        self.down_blocks = nn.ModuleList([down])
        self.mid_blocks = nn.ModuleList([level(CHANNELS) for _ in range(N_MID)])
        up = level(2 * CHANNELS)
        up.append(CausalConv1d(CHANNELS, CHANNELS, 3))
        self.up_blocks = nn.ModuleList([up])
        self.final_block = CausalBlock1D(CHANNELS, CHANNELS)
        self.final_proj = nn.Conv1d(CHANNELS, MEL, 1)

    def _level(self, resnet, blocks, x, mask, t, key_bias):
        x = resnet(x, mask, t)
        x = x.transpose(1, 2)
        for blk in blocks:
            x = blk(x, key_bias)
        return x.transpose(1, 2)

    def forward(self, x, mask, mu, t, spks, cond):
        t = self.time_mlp(self.time_embeddings(t))
        x = torch.cat([x, mu, spks[:, :, None].expand(-1, -1, x.shape[-1]), cond], dim=1)
        key_bias = (1.0 - mask[:, 0, :]) * -1.0e10                      # full (non-streaming) attention over the valid frames
        resnet, blocks, down = self.down_blocks[0]
        x = self._level(resnet, blocks, x, mask, t, key_bias)
        skip = x
        x = down(x * mask)
        for resnet, blocks in self.mid_blocks:
            x = self._level(resnet, blocks, x, mask, t, key_bias)
        resnet, blocks, up = self.up_blocks[0]
        x = self._level(resnet, blocks, torch.cat([x, skip], dim=1), mask, t, key_bias)
        x = up(x * mask)
        x = self.final_block(x, mask)
        return self.final_proj(x * mask) * mask


def cosine_t_span(n: int = N_TIMESTEPS) -> torch.Tensor:
    t = torch.linspace(0, 1, n + 1)
    return 1 - torch.cos(t * 0.5 * torch.pi)


@torch.inference_mode()
def solve_euler(est: ConditionalDecoder, z: torch.Tensor, mu: torch.Tensor, mask: torch.Tensor, spks: torch.Tensor,
                cond: torch.Tensor, n_timesteps: int = N_TIMESTEPS, cfg_rate: float = CFG_RATE, taps: Optional[list] = None):
    """CausalConditionalCFM.solve_euler: the conditioned and the unconditioned estimate come from ONE estimator call on a
    doubled batch (second half: mu, spks, cond zeroed)."""
    B = z.shape[0]
    t_span = cosine_t_span(n_timesteps).to(z.dtype).to(z.device)       # (device-agnostic: the bench's stock-PyTorch row runs this on cuda)
    x = z.clone()
    t, dt = t_span[0], t_span[1] - t_span[0]
    zeros = torch.zeros_like
    for step in range(1, n_timesteps + 1):
        x_in = torch.cat([x, x], 0)
        v = est(x_in, torch.cat([mask, mask], 0), torch.cat([mu, zeros(mu)], 0), t.expand(2 * B).to(z.dtype),
                torch.cat([spks, zeros(spks)], 0), torch.cat([cond, zeros(cond)], 0))
        v = (1.0 + cfg_rate) * v[:B] - cfg_rate * v[B:]
        if taps is not None:
            taps.append(v.clone())
        x = x + dt * v
        t = t + dt
        if step < n_timesteps:
            dt = t_span[step + 1] - t
    return x


class CausalConditionalCFM(nn.Module):
    """The module the engine keeps at `s3gen.flow.decoder`: the estimator, the fixed noise buffer, and
    forward(mu, mask, n_timesteps, temperature, spks, cond) -> (mel, None)."""

    def __init__(self, estimator: ConditionalDecoder, noise_seed: int = 0, max_frames: int = 50 * 300):
        super().__init__()
        self.estimator = estimator
        g = torch.Generator().manual_seed(noise_seed)
        self.rand_noise = torch.randn(1, MEL, max_frames, generator=g)

    @torch.inference_mode()
    def forward(self, mu, mask, n_timesteps: int = N_TIMESTEPS, temperature: float = 1.0, spks=None, cond=None):
        z = self.rand_noise[:, :, :mu.size(2)].to(mu.device).to(mu.dtype) * temperature
        z = z.expand(mu.shape[0], -1, -1)
        return solve_euler(self.estimator, z, mu, mask, spks, cond, n_timesteps=n_timesteps), None


def make_estimator(seed: int = 0) -> ConditionalDecoder:
    torch.manual_seed(seed)
    m = ConditionalDecoder().eval()
    for p in m.parameters():
        p.requires_grad_(False)
    return m


def random_state_dict(seed: int = 0) -> Dict[str, torch.Tensor]:
    return {k: v.clone() for k, v in make_estimator(seed).state_dict().items()}


def load_estimator(sd: Dict[str, torch.Tensor], dtype=torch.float32) -> ConditionalDecoder:
    m = ConditionalDecoder()
    m.load_state_dict(sd, strict=True)
    m = m.to(dtype).eval()
    for p in m.parameters():
        p.requires_grad_(False)
    return m


def synthetic_inputs(B: int, T: int, seed: int = 0, lengths=None):
    """(z, mu, mask, spks, cond): z ~ N(0, 1) like upstream's fixed noise buffer, mu log-mel-like, a prompt-shaped cond."""
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(B, MEL, T, generator=g)
    mu = torch.randn(B, MEL, T, generator=g) * 0.5 - 1.0
    spks = F.normalize(torch.randn(B, MEL, generator=g), dim=1)
    cond = torch.zeros(B, MEL, T)
    n_prompt = min(T // 3, 150)
    cond[:, :, :n_prompt] = torch.randn(B, MEL, n_prompt, generator=g) * 0.5 - 1.0
    mask = torch.ones(B, 1, T)
    if lengths is not None:
        for b, n in enumerate(lengths):
            mask[b, :, n:] = 0
    return z, mu * mask, mask, spks, cond * mask
