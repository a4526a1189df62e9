"""B200FlowFront — speech tokens -> `mu` / `spks` (token embedding, upsampling Conformer encoder, encoder_proj, speaker
projection), the front of the flow step; and B200FlowInference, upstream's `flow.inference(...)` call on top
of it and of B200Flow.

Reference boundary: services/tts/core/synthesizer.py:344-350 `model.generate(...)` -> S3Gen.inference -> flow_inference ->
`self.flow.inference(token=..., token_len=..., prompt_token=..., prompt_token_len=..., prompt_feat=..., prompt_feat_len=...,
embedding=..., finalize=...)` (upstream CausalMaskedDiffWithXvec.inference).  All arithmetic runs in libgonova_hift.so
(`gnv_flow_enc_*`, include/gonova_hift.h): every Linear / Conv1d on the tcgen05 implicit-GEMM kernel, LayerNorm and the
relative-position attention in small CUDA kernels.  No fallback."""
from __future__ import annotations

import ctypes as C
import threading
from typing import Dict, Optional

import torch

from . import _cabi
from .flow import B200Flow, N_TIMESTEPS

VOCAB, DIM, FF, SPK_DIM, MEL = 6561, 512, 2048, 192, 80
N_BLOCKS, N_UP_BLOCKS = 6, 4
PRE_LOOKAHEAD_LEN, TOKEN_MEL_RATIO = 3, 2


def front_layer_shapes():
    """(name, shape) of every tensor of the flow module in front of `decoder`, upstream names."""
    out = [("input_embedding.weight", (VOCAB, DIM)), ("spk_embed_affine_layer.weight", (MEL, SPK_DIM)),
           ("spk_embed_affine_layer.bias", (MEL,))]

    def lin(name, o, i, bias=True):
        out.append((f"{name}.weight", (o, i)))
        if bias:
            out.append((f"{name}.bias", (o,)))

    def norm(name):
        out.extend([(f"{name}.weight", (DIM,)), (f"{name}.bias", (DIM,))])

    def layer(pre):
        for n in ("linear_q", "linear_k", "linear_v", "linear_out"):
            lin(f"{pre}.self_attn.{n}", DIM, DIM)
        lin(f"{pre}.self_attn.linear_pos", DIM, DIM, bias=False)
        out.extend([(f"{pre}.self_attn.pos_bias_u", (8, 64)), (f"{pre}.self_attn.pos_bias_v", (8, 64))])
        lin(f"{pre}.feed_forward.w_1", FF, DIM)
        lin(f"{pre}.feed_forward.w_2", DIM, FF)
        norm(f"{pre}.norm_ff")
        norm(f"{pre}.norm_mha")

    for emb in ("encoder.embed", "encoder.up_embed"):
        lin(f"{emb}.out.0", DIM, DIM)
        norm(f"{emb}.out.1")
    out.extend([("encoder.pre_lookahead_layer.conv1.weight", (DIM, DIM, 4)), ("encoder.pre_lookahead_layer.conv1.bias", (DIM,)),
                ("encoder.pre_lookahead_layer.conv2.weight", (DIM, DIM, 3)), ("encoder.pre_lookahead_layer.conv2.bias", (DIM,)),
                ("encoder.up_layer.conv.weight", (DIM, DIM, 5)), ("encoder.up_layer.conv.bias", (DIM,))])
    for i in range(N_BLOCKS):
        layer(f"encoder.encoders.{i}")
    for i in range(N_UP_BLOCKS):
        layer(f"encoder.up_encoders.{i}")
    norm("encoder.after_norm")
    lin("encoder_proj", MEL, DIM)
    return out


def random_front_state_dict(seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded synthetic weights of the front's architecture (no checkpoint exists here): N(0, 1) embedding rows,
    U(+-1/sqrt(fan_in)) weights and biases, LayerNorm weight 1 / bias 0, position biases U(+-0.3)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    fan = {}
    for name, shape in front_layer_shapes():
        if name == "input_embedding.weight":
            sd[name] = torch.randn(shape, generator=g)
        elif ".norm_" in name or ".out.1." in name or "after_norm" in name:
            sd[name] = torch.ones(shape) if name.endswith("weight") else torch.zeros(shape)
        elif "pos_bias" in name:
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * 0.3
        elif name.endswith("weight"):
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            fan[name[:-6]] = fan_in
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) / fan_in ** 0.5
        else:
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) / fan[name[:-4]] ** 0.5
    return sd


class B200FlowFront:
    """`encode(tokens, token_len, embedding)` -> (mu [B, 80, 2L], spks [B, 80]).  `state_dict` = upstream's `flow.*` weights
    with the prefix stripped (the `decoder.*` entries are ignored here: they belong to B200Flow)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda:0", dtype: str = "bf16", prefix: str = ""):
        if dtype not in ("bf16", "tf32"):
            raise ValueError("dtype must be 'bf16' or 'tf32'")
        self.device = torch.device(device)
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise RuntimeError("B200FlowFront needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", self.device.index if self.device.index is not None else torch.cuda.current_device())
        self.dtype = dtype
        self._lib = _cabi.load()
        sd = {k[len(prefix):]: v for k, v in state_dict.items() if k.startswith(prefix)} if prefix else dict(state_dict)
        names = sorted(k for k in sd if not k.startswith("decoder."))
        arr = (_cabi.GnvWeight * len(names))()
        keep = []
        for i, n in enumerate(names):
            t = sd[n].detach().to("cpu", torch.float32).contiguous()
            keep.append(t)
            arr[i].name = n.encode()
            arr[i].data = C.cast(C.c_void_p(t.data_ptr()), C.POINTER(C.c_float))
            arr[i].ndim = max(1, t.dim())
            for d in range(t.dim()):
                arr[i].shape[d] = t.shape[d]
        h = C.c_void_p()
        rc = self._lib.gnv_flow_enc_create(arr, len(names), self.device.index, _cabi.DTYPE[dtype], 0, C.byref(h))
        _cabi.check(rc, None, "gnv_flow_enc_create")
        self._h = h
        self._ws: Optional[torch.Tensor] = None
        self._lock = threading.Lock()

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self._lib.gnv_flow_enc_destroy(h)
            except Exception:
                pass
            self._h = None

    def workspace_bytes(self, B: int, L: int) -> int:
        n = C.c_size_t()
        _cabi.check(self._lib.gnv_flow_enc_workspace_bytes(self._h, B, L, C.byref(n)), None, "gnv_flow_enc_workspace_bytes")
        return n.value

    def _workspace(self, B: int, L: int) -> torch.Tensor:
        need = self.workspace_bytes(B, L) + 1024
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def _args(self, tokens, token_len, embedding):
        if tokens.dim() != 2:
            raise ValueError("tokens must be [B, L]")
        B, L = tokens.shape
        if tokens.device != self.device:
            raise RuntimeError(f"tokens are on {tokens.device}, the encoder is on {self.device}")
        tokens = tokens.to(torch.int32).contiguous()
        if token_len is not None:
            token_len = torch.as_tensor(token_len, dtype=torch.int32, device=self.device).contiguous()
            if token_len.shape != (B,):
                raise ValueError("token_len must have shape [B]")
        if embedding is not None:
            if embedding.device != self.device:
                raise RuntimeError(f"embedding is on {embedding.device}, the encoder is on {self.device}")
            embedding = embedding.to(torch.float32).contiguous()
            if tuple(embedding.shape) != (B, SPK_DIM):
                raise ValueError(f"embedding must be [B, {SPK_DIM}]")
        return B, L, tokens, token_len, embedding

    @torch.no_grad()
    def encode(self, tokens: torch.Tensor, token_len=None, embedding: Optional[torch.Tensor] = None):
        """tokens [B, L] integer (prompt tokens followed by the utterance's), token_len [B] or None, embedding [B, 192] or
        None -> (mu [B, 80, 2L] fp32, spks [B, 80] fp32 or None)."""
        B, L, tokens, token_len, embedding = self._args(tokens, token_len, embedding)
        mu = torch.empty(B, MEL, TOKEN_MEL_RATIO * L, dtype=torch.float32, device=self.device)
        spks = torch.empty(B, MEL, dtype=torch.float32, device=self.device) if embedding is not None else None
        with self._lock:
            ws = self._workspace(B, L)
            base = ws.data_ptr()
            off = (-base) % 1024
            rc = self._lib.gnv_flow_encode(self._h, C.c_void_p(tokens.data_ptr()),
                                           None if token_len is None else C.c_void_p(token_len.data_ptr()),
                                           None if embedding is None else C.c_void_p(embedding.data_ptr()), B, L,
                                           C.c_void_p(mu.data_ptr()), None if spks is None else C.c_void_p(spks.data_ptr()),
                                           C.c_void_p(base + off), ws.numel() - off,
                                           C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
            _cabi.check(rc, None, "gnv_flow_encode")
        return mu, spks

    @torch.no_grad()
    def profile(self, tokens: torch.Tensor, token_len=None, embedding: Optional[torch.Tensor] = None):
        """One encode with per-launch device times (gnv_flow_encode_profile) -> [(name, kind, ms, algorithmic_flops), ...]."""
        B, L, tokens, token_len, embedding = self._args(tokens, token_len, embedding)
        mu = torch.empty(B, MEL, TOKEN_MEL_RATIO * L, dtype=torch.float32, device=self.device)
        spks = torch.empty(B, MEL, dtype=torch.float32, device=self.device) if embedding is not None else None
        cap = 512
        ms = (C.c_float * cap)(); kinds = (C.c_int32 * cap)(); flops = (C.c_double * cap)()
        names = C.create_string_buffer(cap * _cabi.LAUNCH_NAME_LEN)
        n = C.c_int()
        with self._lock:
            ws = self._workspace(B, L)
            base = ws.data_ptr()
            off = (-base) % 1024
            rc = self._lib.gnv_flow_encode_profile(self._h, C.c_void_p(tokens.data_ptr()),
                                                   None if token_len is None else C.c_void_p(token_len.data_ptr()),
                                                   None if embedding is None else C.c_void_p(embedding.data_ptr()), B, L,
                                                   C.c_void_p(mu.data_ptr()), None if spks is None else C.c_void_p(spks.data_ptr()),
                                                   C.c_void_p(base + off), ws.numel() - off,
                                                   C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream), cap, ms, kinds,
                                                   flops, names, C.byref(n))
            _cabi.check(rc, None, "gnv_flow_encode_profile")
        rows = []
        for i in range(n.value):
            raw = names.raw[i * _cabi.LAUNCH_NAME_LEN:(i + 1) * _cabi.LAUNCH_NAME_LEN]
            rows.append((raw.split(b"\0", 1)[0].decode(), int(kinds[i]), float(ms[i]), float(flops[i])))
        return rows

    def launches(self) -> int:
        n = C.c_int()
        _cabi.check(self._lib.gnv_flow_enc_launches(self._h, C.byref(n)), None, "gnv_flow_enc_launches")
        return n.value


class B200FlowInference(torch.nn.Module):
    """Upstream's flow module as the engine calls it: `inference(token, token_len, prompt_token, prompt_token_len,
    prompt_feat, prompt_feat_len, embedding, finalize)` -> (mel [1, 80, frames of the new tokens], None); tokens -> mel
    entirely on this library (B200FlowFront, then B200Flow's ten Euler steps from the engine's fixed noise buffer)."""

    def __init__(self, state_dict: Optional[Dict[str, torch.Tensor]] = None, device="cuda:0", dtype: str = "bf16",
                 prefix: str = "", noise_seed: int = 0, front: Optional[B200FlowFront] = None,
                 decoder: Optional[B200Flow] = None):
        super().__init__()
        if front is None or decoder is None:
            if state_dict is None:
                raise ValueError("give the flow module's state dict, or both parts")
            sd = {k[len(prefix):]: v for k, v in state_dict.items() if k.startswith(prefix)} if prefix else dict(state_dict)
            front = front or B200FlowFront(sd, device=device, dtype=dtype)
            decoder = decoder or B200Flow(sd, device=device, dtype=dtype, prefix="decoder.estimator.", noise_seed=noise_seed)
        self.front = front
        self.decoder = decoder
        self.device = self.front.device
        self.pre_lookahead_len = PRE_LOOKAHEAD_LEN
        self.token_mel_ratio = TOKEN_MEL_RATIO

    @torch.no_grad()
    def inference(self, token, token_len, prompt_token, prompt_token_len, prompt_feat, prompt_feat_len, embedding,
                  finalize: bool = True, n_timesteps: int = N_TIMESTEPS):
        if token.shape[0] != 1:
            raise ValueError("inference() takes one utterance, like upstream (use front.encode / decoder.decode for batches)")
        dev = self.device
        tok = torch.cat([prompt_token.to(dev), token.to(dev)], dim=1)
        mu, spks = self.front.encode(tok, None, embedding.to(dev))
        if not finalize:
            mu = mu[:, :, : mu.shape[2] - self.pre_lookahead_len * self.token_mel_ratio].contiguous()
        T = mu.shape[2]
        mel_len1 = prompt_feat.shape[1]
        cond = torch.zeros(1, MEL, T, dtype=torch.float32, device=dev)
        cond[:, :, :mel_len1] = prompt_feat.to(dev, torch.float32).transpose(1, 2)
        mask = torch.ones(1, 1, T, dtype=torch.float32, device=dev)
        feat, _ = self.decoder(mu, mask, n_timesteps=n_timesteps, spks=spks, cond=cond)
        return feat[:, :, mel_len1:].float(), None
