"""B200Flow — the CFM flow decoder (tokens' encoder output `mu` -> mel), the step before the vocoder.

Reference boundary: services/tts/core/synthesizer.py:344-350 `model.generate(...)` -> S3Gen.inference -> flow_inference ->
`self.decoder(mu=..., mask=..., spks=..., cond=..., n_timesteps=10)` (upstream CausalConditionalCFM.forward); this class
mirrors that call.  All arithmetic runs in libgonova_hift.so (`gnv_flow_*`, include/gonova_hift.h): the estimator's convs and
projections on the tcgen05 implicit-GEMM kernel, LayerNorm / attention / CFG + Euler in small CUDA kernels.  No fallback."""
from __future__ import annotations

import ctypes as C
import threading
from typing import Dict, Optional

import torch

from . import _cabi

N_TIMESTEPS = 10
CFG_RATE = 0.7


def flow_layer_shapes():
    """(name, shape) of every tensor of the estimator's state dict, upstream names (ConditionalDecoder: in_channels 320,
    channels [256], 4 transformer blocks per level, 12 mid blocks, 8 heads x 64, feed-forward x4)."""
    out = [("time_mlp.linear_1.weight", (1024, 320)), ("time_mlp.linear_1.bias", (1024,)),
           ("time_mlp.linear_2.weight", (1024, 1024)), ("time_mlp.linear_2.bias", (1024,))]

    def level(pre, cin):
        out.extend([(f"{pre}.0.mlp.1.weight", (256, 1024)), (f"{pre}.0.mlp.1.bias", (256,))])
        for blk, ci in (("block1", cin), ("block2", 256)):
            out.extend([(f"{pre}.0.{blk}.block.0.weight", (256, ci, 3)), (f"{pre}.0.{blk}.block.0.bias", (256,)),
                        (f"{pre}.0.{blk}.block.2.weight", (256,)), (f"{pre}.0.{blk}.block.2.bias", (256,))])
        out.extend([(f"{pre}.0.res_conv.weight", (256, cin, 1)), (f"{pre}.0.res_conv.bias", (256,))])
        for j in range(4):
            tp = f"{pre}.1.{j}"
            out.extend([(f"{tp}.norm1.weight", (256,)), (f"{tp}.norm1.bias", (256,)),
                        (f"{tp}.attn1.to_q.weight", (512, 256)), (f"{tp}.attn1.to_k.weight", (512, 256)),
                        (f"{tp}.attn1.to_v.weight", (512, 256)), (f"{tp}.attn1.to_out.0.weight", (256, 512)),
                        (f"{tp}.attn1.to_out.0.bias", (256,)), (f"{tp}.norm3.weight", (256,)), (f"{tp}.norm3.bias", (256,)),
                        (f"{tp}.ff.net.0.proj.weight", (1024, 256)), (f"{tp}.ff.net.0.proj.bias", (1024,)),
                        (f"{tp}.ff.net.2.weight", (256, 1024)), (f"{tp}.ff.net.2.bias", (256,))])

    level("down_blocks.0", 320)
    out.extend([("down_blocks.0.2.weight", (256, 256, 3)), ("down_blocks.0.2.bias", (256,))])
    for i in range(12):
        level(f"mid_blocks.{i}", 256)
    level("up_blocks.0", 512)
    out.extend([("up_blocks.0.2.weight", (256, 256, 3)), ("up_blocks.0.2.bias", (256,)),
                ("final_block.block.0.weight", (256, 256, 3)), ("final_block.block.0.bias", (256,)),
                ("final_block.block.2.weight", (256,)), ("final_block.block.2.bias", (256,)),
                ("final_proj.weight", (80, 256, 1)), ("final_proj.bias", (80,))])
    return out


def random_flow_state_dict(seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded synthetic weights of the estimator's architecture (no checkpoint exists here): U(+-1/sqrt(fan_in)) weights
    and biases like torch's default initialisers, LayerNorm weight 1 / bias 0."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    fan = {}
    for name, shape in flow_layer_shapes():
        is_ln = ".block.2." in name or ".norm1." in name or ".norm3." in name
        if is_ln:
            sd[name] = torch.ones(shape) if name.endswith("weight") else torch.zeros(shape)
            continue
        if name.endswith("weight"):
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            fan[name[:-6]] = fan_in
            bound = 1.0 / fan_in ** 0.5
        else:
            bound = 1.0 / fan[name[:-4]] ** 0.5
        sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    return sd


class B200Flow(torch.nn.Module):
    """`forward(mu, mask, n_timesteps, temperature, spks, cond)` like upstream's CausalConditionalCFM (returns (mel, None)).
    `state_dict` = the estimator's weights under upstream's names (`flow.decoder.estimator.*` with the prefix stripped)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda:0", dtype: str = "bf16", prefix: str = "",
                 noise_seed: int = 0, max_frames: int = 50 * 300):
        super().__init__()
        if dtype not in ("bf16", "tf32"):
            raise ValueError("dtype must be 'bf16' or 'tf32'")
        self.device = torch.device(device)
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise RuntimeError("B200Flow needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", self.device.index if self.device.index is not None else torch.cuda.current_device())
        self.dtype = dtype
        self._lib = _cabi.load()
        sd = {k[len(prefix):]: v for k, v in state_dict.items() if k.startswith(prefix)} if prefix else dict(state_dict)
        names = sorted(sd)
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
        rc = self._lib.gnv_flow_create(arr, len(names), self.device.index, _cabi.DTYPE[dtype], 0, C.byref(h))
        _cabi.check(rc, None, "gnv_flow_create")
        self._h = h
        self._ws: Optional[torch.Tensor] = None
        self._lock = threading.Lock()
        # upstream keeps ONE fixed noise buffer (`self.rand_noise = torch.randn([1, 80, 50 * 300])`) and slices it per call
        g = torch.Generator().manual_seed(noise_seed)
        self.rand_noise = torch.randn(1, 80, max_frames, generator=g).to(self.device)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self._lib.gnv_flow_destroy(h)
            except Exception:
                pass
            self.__dict__["_h"] = None          # (not nn.Module.__setattr__: it may be half torn down at interpreter exit)

    def workspace_bytes(self, B: int, T: int) -> int:
        n = C.c_size_t()
        _cabi.check(self._lib.gnv_flow_workspace_bytes(self._h, B, T, C.byref(n)), None, "gnv_flow_workspace_bytes")
        return n.value

    def _workspace(self, B: int, T: int) -> torch.Tensor:
        need = self.workspace_bytes(B, T) + 1024
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def _f32(self, t: torch.Tensor, name: str, shape) -> torch.Tensor:
        if t.device != self.device:
            raise RuntimeError(f"{name} is on {t.device}, the flow decoder is on {self.device}")
        t = t.to(torch.float32).contiguous()
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name} must have shape {tuple(shape)}, got {tuple(t.shape)}")
        return t

    @torch.no_grad()
    def decode(self, z: torch.Tensor, mu: torch.Tensor, spks: torch.Tensor, cond: torch.Tensor, lengths=None,
               n_timesteps: int = N_TIMESTEPS, cfg_rate: float = CFG_RATE) -> torch.Tensor:
        """solve_euler from the initial noise z: all [B, 80, T] but spks [B, 80]; lengths [B] frames or None -> mel [B, 80, T]."""
        B, Cm, T = mu.shape
        if Cm != 80:
            raise ValueError("mu must be [B, 80, T]")
        z = self._f32(z, "z", (B, 80, T))
        mu = self._f32(mu, "mu", (B, 80, T))
        cond = self._f32(cond, "cond", (B, 80, T))
        spks = self._f32(spks, "spks", (B, 80))
        if lengths is not None:
            lengths = torch.as_tensor(lengths, dtype=torch.int32, device=self.device).contiguous()
            if lengths.shape != (B,):
                raise ValueError("lengths must have shape [B]")
        mel = torch.empty(B, 80, T, dtype=torch.float32, device=self.device)
        with self._lock:
            ws = self._workspace(B, T)
            base = ws.data_ptr()
            off = (-base) % 1024
            rc = self._lib.gnv_flow_decode(self._h, C.c_void_p(z.data_ptr()), C.c_void_p(mu.data_ptr()),
                                           C.c_void_p(spks.data_ptr()), C.c_void_p(cond.data_ptr()),
                                           None if lengths is None else C.c_void_p(lengths.data_ptr()), B, T, int(n_timesteps),
                                           C.c_float(cfg_rate), C.c_void_p(mel.data_ptr()), C.c_void_p(base + off),
                                           ws.numel() - off, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
            _cabi.check(rc, None, "gnv_flow_decode")
        return mel

    @torch.no_grad()
    def forward(self, mu, mask, n_timesteps: int = N_TIMESTEPS, temperature: float = 1.0, spks=None, cond=None, **_):
        """Upstream's call shape: `decoder(mu=h, mask=mask, spks=embedding, cond=conds, n_timesteps=10)` -> (mel, None)."""
        B, _, T = mu.shape
        if T > self.rand_noise.shape[2]:
            raise ValueError("utterance longer than the noise buffer")
        z = (self.rand_noise[:, :, :T] * temperature).expand(B, -1, -1).contiguous()
        lengths = mask.reshape(B, T).sum(dim=1).to(torch.int32)
        return self.decode(z, mu, spks, cond, lengths=lengths, n_timesteps=n_timesteps), None

    @torch.no_grad()
    def profile(self, z, mu, spks, cond, n_timesteps: int = 1, cfg_rate: float = CFG_RATE):
        """One decode with per-launch device times (gnv_flow_profile) -> [(name, kind, ms, algorithmic_flops), ...]."""
        B, _, T = mu.shape
        z = self._f32(z, "z", (B, 80, T)); mu = self._f32(mu, "mu", (B, 80, T))
        cond = self._f32(cond, "cond", (B, 80, T)); spks = self._f32(spks, "spks", (B, 80))
        mel = torch.empty(B, 80, T, dtype=torch.float32, device=self.device)
        cap = 4096 * n_timesteps
        ms = (C.c_float * cap)(); kinds = (C.c_int32 * cap)(); flops = (C.c_double * cap)()
        names = C.create_string_buffer(cap * _cabi.LAUNCH_NAME_LEN)
        n = C.c_int()
        with self._lock:
            ws = self._workspace(B, T)
            base = ws.data_ptr()
            off = (-base) % 1024
            rc = self._lib.gnv_flow_profile(self._h, C.c_void_p(z.data_ptr()), C.c_void_p(mu.data_ptr()),
                                            C.c_void_p(spks.data_ptr()), C.c_void_p(cond.data_ptr()), None, B, T,
                                            int(n_timesteps), C.c_float(cfg_rate), C.c_void_p(mel.data_ptr()),
                                            C.c_void_p(base + off), ws.numel() - off,
                                            C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream), cap, ms, kinds, flops,
                                            names, C.byref(n))
            _cabi.check(rc, None, "gnv_flow_profile")
        rows = []
        for i in range(n.value):
            raw = names.raw[i * _cabi.LAUNCH_NAME_LEN:(i + 1) * _cabi.LAUNCH_NAME_LEN]
            rows.append((raw.split(b"\0", 1)[0].decode(), int(kinds[i]), float(ms[i]), float(flops[i])))
        return rows

    def launches(self, n_timesteps: int = N_TIMESTEPS) -> int:
        n = C.c_int()
        _cabi.check(self._lib.gnv_flow_launches(self._h, n_timesteps, C.byref(n)), None, "gnv_flow_launches")
        return n.value
