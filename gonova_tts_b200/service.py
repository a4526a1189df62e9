"""Service-side glue: how the reference's StreamingSynthesizer gets the B200 decoder.

install() is the whole integration: call it in StreamingSynthesizer.load() right after
`self.model = ChatterboxTTS.from_pretrained(device=self.device)` (reference
services/tts/core/synthesizer.py:185) and before the warm-up loop (:199-207).  Nothing in server.py,
queue_manager.py or voice_manager.py changes; the wire format stays float32 LE PCM
(synthesizer.py:352-357, server.py:152) unless the caller asks for int16 via PcmSink."""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch

from .decoder import B200HiFT, mulaw_encode, pcm_tail, trim_fade_window


def install(model, dtype: Optional[str] = None, device=None, bucket_frames: int = 8,
            max_frames: int = 3000) -> B200HiFT:
    """Swap `model.s3gen.mel2wav` for a B200HiFT built from its weights.  Returns the new decoder.
    Raises (never falls back) if the CUDA library is missing or the device is not a B200.

    The service decodes one sentence per call, each with its own length: `bucket_frames` rounds T up to a multiple
    (masked, results unchanged) so the launch-plan cache keeps hitting, and the workspace for `max_frames` (60 s) is
    reserved once so plans never move."""
    dtype = dtype or os.environ.get("GONOVA_DECODER_DTYPE", "bf16")
    old = model.s3gen.mel2wav
    new = B200HiFT.from_module(old, device=device, dtype=dtype, bucket_frames=bucket_frames)
    if max_frames > 0:
        new.reserve(1, max_frames)
    model.s3gen.mel2wav = new
    return new


class PcmSink:
    """Device float32 wav -> host bytes for the WebSocket, through one pinned staging buffer.
    fmt 'f32' reproduces the reference's `audio.astype(np.float32).tobytes()` bytes exactly;
    fmt 'i16' is the opt-in int16 pack (clamp + round-half-even, one kernel); fmt 'mulaw' adds G.711 companding
    (8 bits per sample, the telephony clients of the reference's `phone` extra)."""

    def __init__(self, device, max_samples: int = 24000 * 60, fmt: str = "f32", limit: float = 0.99):
        if fmt not in ("f32", "i16", "mulaw"):
            raise ValueError("fmt must be 'f32', 'i16' or 'mulaw'")
        self.fmt, self.limit = fmt, limit
        self.device = torch.device(device)
        host_dtype = {"f32": torch.float32, "i16": torch.int16, "mulaw": torch.uint8}[fmt]
        self._host = torch.empty(max_samples, dtype=host_dtype).pin_memory()

    @torch.no_grad()
    def to_bytes(self, wav: torch.Tensor, trim_fade: bool = False) -> bytes:
        wav = wav.reshape(1, -1)
        n = wav.shape[1]
        if n > self._host.numel():
            self._host = torch.empty(n, dtype=self._host.dtype).pin_memory()
        fw = trim_fade_window(self.device) if trim_fade else None
        i16, f32 = pcm_tail(wav, None, fw, self.limit, want_i16=self.fmt != "f32", want_f32=self.fmt == "f32")
        src = f32 if self.fmt == "f32" else (i16 if self.fmt == "i16" else mulaw_encode(i16))
        self._host[:n].copy_(src.reshape(-1), non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self._host[:n].numpy().tobytes()
