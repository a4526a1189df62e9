"""CPU oracle of the FRONT of the flow step ("tokens -> mel"): token embedding, the upsampling Conformer
encoder and `encoder_proj`, i.e. everything that turns speech tokens into the `mu` / `spks` the CFM decoder
(oracle/flow_ref.py) conditions on.  TEST INFRASTRUCTURE ONLY (tests/, smoke and bench's CPU legs may import it).

PARITY UNPINNED, like the other oracle files and for the same reason: the arithmetic lives in the third-party `chatterbox`
package (services/tts/core/synthesizer.py:167, not vendored / pinned / installable here), reached from
`self.model.generate(...)` (synthesizer.py:344-350) through `S3Gen.inference -> flow_inference -> flow.inference(...)`.
Restated from the published architecture of that engine's flow front (CosyVoice-2 lineage):

  CausalMaskedDiffWithXvec.inference      F.normalize(x-vector) -> spk_embed_affine_layer (192 -> 80);
                                          input_embedding(clamp(token, 0)) * mask -> encoder -> encoder_proj (512 -> 80) = mu;
                                          cond = [prompt mel | zeros]; decoder(mu, mask, spks, cond, 10 steps); crop the prompt
  UpsampleConformerEncoder                embed = Linear + LayerNorm(1e-5), x * sqrt(512), ESPnet relative positions
                                          PreLookaheadLayer (conv k4 looking 3 tokens ahead, LeakyReLU, causal conv k3, + x)
                                          6 x ConformerEncoderLayer -> Upsample1D (nearest x2, left pad 4, conv k5)
                                          -> up_embed -> 4 x ConformerEncoderLayer -> after_norm
  ConformerEncoderLayer (pre-norm, no macaron, no conv module, LayerNorm eps 1e-12)
                                          x + RelPositionMultiHeadedAttention(norm_mha(x)); x + FFN_swish(norm_ff(x))
  RelPositionMultiHeadedAttention         8 x 64; scores = ((q + u) k^T + rel_shift((q + v) p^T)) / 8, p = linear_pos(pos_emb)

State-dict names follow upstream's `flow.*` (`input_embedding.weight`, `encoder.encoders.3.self_attn.linear_pos.weight`,
`encoder.up_layer.conv.weight`, `encoder_proj.bias`, ...).  Upstream asserts a batch of ONE utterance; `encode()` here
defines a batch as "every utterance alone, at its own length" (rows past a length are zero), which is what the CUDA path
must reproduce for ragged batches."""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F
from torch import nn

VOCAB = 6561
DIM = 512
HEADS = 8
HEAD_DIM = 64
FF = 2048
N_BLOCKS = 6
N_UP_BLOCKS = 4
PRE_LOOKAHEAD = 3
UP_STRIDE = 2
SPK_DIM = 192
MEL = 80


class EspnetRelPositionalEncoding(nn.Module):
    """x * sqrt(d); pos_emb [1, 2T-1, d]: row r holds the sinusoid of relative position (T-1) - r."""

    def __init__(self, d_model: int):
        super().__init__()
        self.d_model = d_model
        self.xscale = math.sqrt(d_model)

    def position_encoding(self, size: int) -> torch.Tensor:
        d = self.d_model
        position = torch.arange(0, size, dtype=torch.float32).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * -(math.log(10000.0) / d))
        pos = torch.zeros(size, d)
        neg = torch.zeros(size, d)
        pos[:, 0::2] = torch.sin(position * div_term)
        pos[:, 1::2] = torch.cos(position * div_term)
        neg[:, 0::2] = torch.sin(-1 * position * div_term)
        neg[:, 1::2] = torch.cos(-1 * position * div_term)
        pos = torch.flip(pos, [0]).unsqueeze(0)
        neg = neg[1:].unsqueeze(0)
        return torch.cat([pos, neg], dim=1)

    def forward(self, x: torch.Tensor):
        return x * self.xscale, self.position_encoding(x.size(1)).to(x.dtype).to(x.device)


class LinearNoSubsampling(nn.Module):
    def __init__(self, idim: int, odim: int):
        super().__init__()
        self.out = nn.Sequential(nn.Linear(idim, odim), nn.LayerNorm(odim, eps=1e-5), nn.Identity())
        self.pos_enc = EspnetRelPositionalEncoding(odim)

    def forward(self, x):
        return self.pos_enc(self.out(x))


class PreLookaheadLayer(nn.Module):
    def __init__(self, channels: int, pre_lookahead_len: int):
        super().__init__()
        self.pre_lookahead_len = pre_lookahead_len
        self.conv1 = nn.Conv1d(channels, channels, pre_lookahead_len + 1)
        self.conv2 = nn.Conv1d(channels, channels, 3)

    def forward(self, inputs):                     # [B, T, C]
        y = inputs.transpose(1, 2)
        y = F.pad(y, (0, self.pre_lookahead_len))
        y = F.leaky_relu(self.conv1(y))
        y = F.pad(y, (2, 0))
        y = self.conv2(y)
        return y.transpose(1, 2) + inputs


class Upsample1D(nn.Module):
    def __init__(self, channels: int, out_channels: int, stride: int):
        super().__init__()
        self.stride = stride
        self.conv = nn.Conv1d(channels, out_channels, stride * 2 + 1)

    def forward(self, x):                          # [B, C, T]
        y = F.interpolate(x, scale_factor=float(self.stride), mode="nearest")
        y = F.pad(y, (self.stride * 2, 0))
        return self.conv(y)


class RelPositionMultiHeadedAttention(nn.Module):
    def __init__(self, n_head: int, n_feat: int):
        super().__init__()
        self.h = n_head
        self.d_k = n_feat // n_head
        self.linear_q = nn.Linear(n_feat, n_feat)
        self.linear_k = nn.Linear(n_feat, n_feat)
        self.linear_v = nn.Linear(n_feat, n_feat)
        self.linear_out = nn.Linear(n_feat, n_feat)
        self.linear_pos = nn.Linear(n_feat, n_feat, bias=False)
        self.pos_bias_u = nn.Parameter(torch.empty(self.h, self.d_k))
        self.pos_bias_v = nn.Parameter(torch.empty(self.h, self.d_k))
        nn.init.xavier_uniform_(self.pos_bias_u)
        nn.init.xavier_uniform_(self.pos_bias_v)

    @staticmethod
    def rel_shift(x: torch.Tensor) -> torch.Tensor:
        """[B, H, T, 2T-1] indexed by row r of pos_emb -> [B, H, T, T] indexed by key: out[i, j] = x[i, (T-1) - i + j]."""
        b, h, t, n = x.shape
        zero_pad = torch.zeros((b, h, t, 1), dtype=x.dtype, device=x.device)
        x_padded = torch.cat([zero_pad, x], dim=-1).view(b, h, n + 1, t)
        return x_padded[:, :, 1:].view_as(x)[:, :, :, : n // 2 + 1]

    def forward(self, x, pos_emb):                 # one utterance at its own length: no padding mask
        b, t, _ = x.shape
        q = self.linear_q(x).view(b, t, self.h, self.d_k)
        k = self.linear_k(x).view(b, t, self.h, self.d_k).transpose(1, 2)
        v = self.linear_v(x).view(b, t, self.h, self.d_k).transpose(1, 2)
        p = self.linear_pos(pos_emb).view(1, -1, self.h, self.d_k).transpose(1, 2)
        q_u = (q + self.pos_bias_u).transpose(1, 2)
        q_v = (q + self.pos_bias_v).transpose(1, 2)
        ac = torch.matmul(q_u, k.transpose(-2, -1))
        bd = torch.matmul(q_v, p.transpose(-2, -1))
        if ac.shape != bd.shape:
            bd = self.rel_shift(bd)
        attn = torch.softmax((ac + bd) / math.sqrt(self.d_k), dim=-1)
        y = torch.matmul(attn, v).transpose(1, 2).reshape(b, t, self.h * self.d_k)
        return self.linear_out(y)


class PositionwiseFeedForward(nn.Module):
    def __init__(self, idim: int, hidden: int):
        super().__init__()
        self.w_1 = nn.Linear(idim, hidden)
        self.w_2 = nn.Linear(hidden, idim)

    def forward(self, x):
        return self.w_2(F.silu(self.w_1(x)))


class ConformerEncoderLayer(nn.Module):
    def __init__(self, size: int):
        super().__init__()
        self.self_attn = RelPositionMultiHeadedAttention(HEADS, size)
        self.feed_forward = PositionwiseFeedForward(size, FF)
        self.norm_ff = nn.LayerNorm(size, eps=1e-12)
        self.norm_mha = nn.LayerNorm(size, eps=1e-12)

    def forward(self, x, pos_emb):
        x = x + self.self_attn(self.norm_mha(x), pos_emb)
        return x + self.feed_forward(self.norm_ff(x))


class UpsampleConformerEncoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.embed = LinearNoSubsampling(DIM, DIM)
        self.pre_lookahead_layer = PreLookaheadLayer(DIM, PRE_LOOKAHEAD)
        self.encoders = nn.ModuleList([ConformerEncoderLayer(DIM) for _ in range(N_BLOCKS)])
        self.up_layer = Upsample1D(DIM, DIM, UP_STRIDE)
        self.up_embed = LinearNoSubsampling(DIM, DIM)
        self.up_encoders = nn.ModuleList([ConformerEncoderLayer(DIM) for _ in range(N_UP_BLOCKS)])
        self.after_norm = nn.LayerNorm(DIM, eps=1e-5)

    def forward(self, xs):                         # [1, L, 512] -> [1, 2L, 512]
        xs, pos_emb = self.embed(xs)
        xs = self.pre_lookahead_layer(xs)
        for layer in self.encoders:
            xs = layer(xs, pos_emb)
        xs = self.up_layer(xs.transpose(1, 2)).transpose(1, 2)
        xs, pos_emb = self.up_embed(xs)
        for layer in self.up_encoders:
            xs = layer(xs, pos_emb)
        return self.after_norm(xs)


class FlowFront(nn.Module):
    """The parameters of upstream's flow module in front of `decoder`, under upstream's names."""

    def __init__(self):
        super().__init__()
        self.input_embedding = nn.Embedding(VOCAB, DIM)
        self.spk_embed_affine_layer = nn.Linear(SPK_DIM, MEL)
        self.encoder = UpsampleConformerEncoder()
        self.encoder_proj = nn.Linear(DIM, MEL)

    def speaker(self, embedding: torch.Tensor) -> torch.Tensor:      # [B, 192] -> [B, 80]
        return self.spk_embed_affine_layer(F.normalize(embedding, dim=1))

    def encode_one(self, token: torch.Tensor) -> torch.Tensor:       # [L] int -> mu [80, 2L]
        x = self.input_embedding(torch.clamp(token, min=0)).unsqueeze(0)
        h = self.encoder_proj(self.encoder(x))
        return h[0].transpose(0, 1).contiguous()

    def encode(self, tokens: torch.Tensor, token_len: Optional[torch.Tensor] = None) -> torch.Tensor:
        """tokens [B, L] (prompt tokens followed by the utterance's) -> mu [B, 80, 2L]; utterance b is encoded alone at
        token_len[b], frames past 2 * token_len[b] are zero."""
        B, L = tokens.shape
        mu = torch.zeros(B, MEL, UP_STRIDE * L, device=tokens.device)
        for b in range(B):
            n = L if token_len is None else int(token_len[b])
            if n > 0:
                mu[b, :, : UP_STRIDE * n] = self.encode_one(tokens[b, :n])
        return mu


def make_front(seed: int = 0) -> FlowFront:
    """Seeded default PyTorch init; LayerNorm affine parameters are perturbed so that they are exercised."""
    torch.manual_seed(seed)
    m = FlowFront()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, nn.LayerNorm):
                mod.weight.add_(0.1 * torch.randn(mod.weight.shape, generator=g))
                mod.bias.add_(0.1 * torch.randn(mod.bias.shape, generator=g))
    return m.eval()


def random_state_dict(seed: int = 0) -> Dict[str, torch.Tensor]:
    return {k: v.detach().clone() for k, v in make_front(seed).state_dict().items()}


def load_front(sd: Dict[str, torch.Tensor]) -> FlowFront:
    m = FlowFront()
    m.load_state_dict({k: v.float() for k, v in sd.items()})
    return m.eval()


def synthetic_tokens(B: int, L: int, seed: int = 0, lengths=None):
    """(tokens [B, L] int32 with a few negative ids (upstream clamps them to 0), token_len [B] int32, x-vectors [B, 192])."""
    g = torch.Generator().manual_seed(seed)
    tokens = torch.randint(0, VOCAB, (B, L), generator=g, dtype=torch.int32)
    tokens[:, ::17] = -1
    if lengths is None:
        token_len = torch.full((B,), L, dtype=torch.int32)
    else:
        token_len = torch.tensor(lengths, dtype=torch.int32)
    emb = torch.randn(B, SPK_DIM, generator=g)
    return tokens, token_len, emb


def flow_inference(front: FlowFront, cfm, token, prompt_token, prompt_feat, embedding, n_timesteps: int = 10,
                   finalize: bool = True):
    """Upstream's `flow.inference` for one utterance with `finalize=True`: token [1, n], prompt_token [1, m], prompt_feat
    [1, mel_len1, 80], embedding [1, 192] -> mel [1, 80, 2 (m + n) - mel_len1].  `cfm` is oracle/flow_ref's
    CausalConditionalCFM (the estimator and the engine's fixed noise buffer).  `finalize=False` (more tokens will follow):
    upstream drops the last pre_lookahead_len * token_mel_ratio = 6 encoder frames before the decoder."""
    spks = front.speaker(embedding)
    tok = torch.cat([prompt_token, token], dim=1)
    mu = front.encode(tok)
    if not finalize:
        mu = mu[:, :, : mu.shape[2] - PRE_LOOKAHEAD * UP_STRIDE]
    T = mu.shape[2]
    mel_len1 = prompt_feat.shape[1]
    cond = torch.zeros(1, MEL, T)
    cond[:, :, :mel_len1] = prompt_feat.transpose(1, 2)
    mask = torch.ones(1, 1, T)
    feat, _ = cfm(mu, mask, n_timesteps=n_timesteps, spks=spks, cond=cond)
    return feat[:, :, mel_len1:]
