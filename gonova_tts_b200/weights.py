"""Weights of the decoder: folding of the upstream `mel2wav.*` state dict, and seeded synthetic
weights of the same architecture (no checkpoint exists in this environment).

The engine the service loads (`ChatterboxTTS.from_pretrained`, reference
services/tts/core/synthesizer.py:185) keeps its vocoder under `s3gen.mel2wav`; its convolutions
are weight-normalised, stored either as `<m>.parametrizations.weight.original0/original1`
(torch >= 2.1) or as `<m>.weight_g/<m>.weight_v`.  `fold_state_dict` turns either form into the
effective tensors the C ABI takes (GnvWeight names, include/gonova_hift.h)."""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch

UPSAMPLE_RATES = (8, 5, 3)
UPSAMPLE_KERNELS = (16, 11, 7)
RESBLOCK_KERNELS = (3, 7, 11)
SOURCE_RESBLOCK_KERNELS = (7, 7, 11)
SOURCE_DOWN = ((30, 15, 7), (6, 3, 1), (1, 1, 0))   # (kernel, stride, padding)
BASE_CHANNELS = 512
MEL_CHANNELS = 80
N_FFT = 16


def layer_specs() -> List[Tuple[str, str, Tuple[int, ...], bool]]:
    """(module path, kind, weight shape, weight_normed) for every parametrised module."""
    specs: List[Tuple[str, str, Tuple[int, ...], bool]] = []
    specs.append(("m_source.l_linear", "linear", (1, 9), False))
    specs.append(("conv_pre", "conv", (BASE_CHANNELS, MEL_CHANNELS, 7), True))
    for i, (u, k) in enumerate(zip(UPSAMPLE_RATES, UPSAMPLE_KERNELS)):
        specs.append((f"ups.{i}", "convT", (BASE_CHANNELS >> i, BASE_CHANNELS >> (i + 1), k), True))
    for i, (k, s, p) in enumerate(SOURCE_DOWN):
        specs.append((f"source_downs.{i}", "conv", (BASE_CHANNELS >> (i + 1), N_FFT + 2, k), False))

    def resblock(prefix: str, ch: int, k: int):
        for d in range(3):
            specs.append((f"{prefix}.convs1.{d}", "conv", (ch, ch, k), True))
        for d in range(3):
            specs.append((f"{prefix}.convs2.{d}", "conv", (ch, ch, k), True))
        for d in range(3):
            specs.append((f"{prefix}.activations1.{d}", "snake", (ch,), False))
        for d in range(3):
            specs.append((f"{prefix}.activations2.{d}", "snake", (ch,), False))

    for i, k in enumerate(SOURCE_RESBLOCK_KERNELS):
        resblock(f"source_resblocks.{i}", BASE_CHANNELS >> (i + 1), k)
    for i in range(3):
        for j, k in enumerate(RESBLOCK_KERNELS):
            resblock(f"resblocks.{3 * i + j}", BASE_CHANNELS >> (i + 1), k)
    specs.append(("conv_post", "conv", (N_FFT + 2, BASE_CHANNELS >> 3, 7), True))
    for i in range(5):
        specs.append((f"f0_predictor.condnet.{2 * i}", "conv",
                      (BASE_CHANNELS, MEL_CHANNELS if i == 0 else BASE_CHANNELS, 3), True))
    specs.append(("f0_predictor.classifier", "linear", (1, BASE_CHANNELS), False))
    return specs


def _norm_except_dim0(v: torch.Tensor) -> torch.Tensor:
    return v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, *([1] * (v.dim() - 1)))


def random_state_dict(seed: int = 0, corners: bool = False) -> Dict[str, torch.Tensor]:
    """Seeded synthetic weights in the upstream state-dict layout (torch default initialisers:
    U(+-1/sqrt(fan_in)) for weights and biases, Snake alpha = 1, weight-norm g = ||v||).

    corners=True moves the parameters so that the non-linear corners of the decoder are reached:
    Snake alpha ~ U(0.5, 2), an f0 head that spans 0..~300 Hz (voiced and unvoiced frames), and a
    x20 output gain so that min(exp(.), 100) and clamp(+-0.99) both fire."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for path, kind, shape, wn in layer_specs():
        if kind == "snake":
            a = torch.ones(shape)
            if corners:
                a = 0.5 + 1.5 * torch.rand(shape, generator=g)
            sd[f"{path}.alpha"] = a
            continue
        if kind == "convT":
            fan_in = shape[1] * shape[2]          # torch computes fan_in from dim 1 for every conv
        elif kind == "conv":
            fan_in = shape[1] * shape[2]
        else:
            fan_in = shape[1]
        bound = 1.0 / math.sqrt(fan_in)
        w = (torch.rand(shape, generator=g) * 2 - 1) * bound
        n_bias = shape[1] if kind == "convT" else shape[0]
        b = (torch.rand(n_bias, generator=g) * 2 - 1) * bound
        if corners and path == "f0_predictor.classifier":
            w = w * 3000.0
            b = torch.full_like(b, -100.0)
        if wn:
            gnorm = _norm_except_dim0(w)
            if corners and path == "conv_post":
                gnorm = gnorm * 20.0
            sd[f"{path}.parametrizations.weight.original0"] = gnorm
            sd[f"{path}.parametrizations.weight.original1"] = w
        else:
            sd[f"{path}.weight"] = w
        sd[f"{path}.bias"] = b
    return sd


def fold_state_dict(sd: Dict[str, torch.Tensor], prefix: str = "") -> Dict[str, torch.Tensor]:
    """Upstream state dict (optionally still carrying a `mel2wav.` style prefix) -> effective fp32
    CPU tensors named `<module>.weight|bias|alpha`.  Weight-norm: w = g * v / ||v||, the norm taken
    over every dim but 0 — for ConvTranspose1d that is per *input* channel (upstream quirk)."""
    if prefix:
        sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    out: Dict[str, torch.Tensor] = {}
    for path, kind, shape, wn in layer_specs():
        if kind == "snake":
            key = f"{path}.alpha"
            if key not in sd:
                raise KeyError(f"state dict has no '{key}'")
            out[key] = sd[key].detach().to("cpu", torch.float32).contiguous()
            continue
        w = None
        if f"{path}.weight" in sd:
            w = sd[f"{path}.weight"].detach().to("cpu", torch.float32)
        else:
            for gk, vk in ((f"{path}.parametrizations.weight.original0", f"{path}.parametrizations.weight.original1"),
                           (f"{path}.weight_g", f"{path}.weight_v")):
                if gk in sd and vk in sd:
                    gt = sd[gk].detach().to("cpu", torch.float32)
                    v = sd[vk].detach().to("cpu", torch.float32)
                    w = v * (gt / _norm_except_dim0(v))
                    break
        if w is None:
            raise KeyError(f"state dict has no weight for '{path}'")
        if tuple(w.shape) != tuple(shape):
            raise ValueError(f"'{path}.weight' has shape {tuple(w.shape)}, expected {tuple(shape)}")
        out[f"{path}.weight"] = w.contiguous()
        bk = f"{path}.bias"
        if bk not in sd:
            raise KeyError(f"state dict has no '{bk}'")
        out[bk] = sd[bk].detach().to("cpu", torch.float32).contiguous()
    return out


def write_flat(sd: Dict[str, torch.Tensor], path: str, prefix: str = "") -> int:
    """Folded weights as one flat binary file for hosts that are not Python (examples/c_host.c):
    int32 n; per tensor: int32 name_len, name, int32 ndim, int64 shape[ndim], float32 data (little endian)."""
    import struct

    folded = fold_state_dict(sd, prefix=prefix)
    with open(path, "wb") as f:
        f.write(struct.pack("<i", len(folded)))
        for name in sorted(folded):
            t = folded[name].detach().to(torch.float32).contiguous().cpu()
            nb = name.encode()
            f.write(struct.pack("<i", len(nb)) + nb + struct.pack("<i", t.dim()))
            f.write(struct.pack("<%dq" % t.dim(), *t.shape))
            f.write(t.numpy().tobytes())
    return len(folded)
