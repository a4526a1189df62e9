"""Chunked decode + crossfade + PCM pack: the streaming tail of the service path.

The reference service ships one float32 chunk per sentence (services/tts/core/synthesizer.py:320-321,
:352-357; services/tts/server.py:150-155) and config.yaml:9 names a 50-token (= 100 mel frame, 2 s)
chunk it never uses.  StreamingDecoder makes that chunking real for the vocoder: the source signal is
generated once for the whole utterance (its phase is a running sum over time), then every chunk of
`chunk` frames is decoded with a `halo`-frame context on both sides (the decoder's receptive field is
+-14.3 frames), its head is crossfaded with the previous chunk's look-ahead and the result is clamped
and packed to int16 by one kernel."""
from __future__ import annotations

from typing import Iterator, Optional, Tuple

import torch

from .decoder import B200HiFT, SAMPLES_PER_FRAME, fade_window, pcm_tail, trim_fade_window


def chunk_plan(T: int, chunk: int = 100, halo: int = 16):
    """Frame windows of a chunked decode.  Chunk c owns frames [c*chunk, min(T, (c+1)*chunk)); it is
    decoded on [lo, hi) = owned frames +- halo plus one look-ahead frame whose 480 samples are held
    back and crossfaded into the next chunk's head.  Yields (own_lo, own_hi, lo, hi, last)."""
    c = 0
    while c * chunk < T:
        own_lo = c * chunk
        own_hi = min(T, own_lo + chunk)
        last = own_hi >= T
        lo = max(0, own_lo - halo)
        hi = T if last else min(T, own_hi + 1 + halo)
        yield own_lo, own_hi, lo, hi, last
        c += 1


class StreamingDecoder:
    def __init__(self, hift: B200HiFT, chunk: int = 100, halo: int = 16, fade: int = 480, limit: float = 0.99,
                 trim_fade: bool = False):
        self.hift = hift
        self.chunk, self.halo, self.fade, self.limit = chunk, halo, fade, limit
        self.trim_fade = trim_fade
        self._w = fade_window(fade, hift.device)
        self._tw = trim_fade_window(hift.device)

    @torch.no_grad()
    def stream(self, mel: torch.Tensor, s: Optional[torch.Tensor] = None, want_i16: bool = True,
               want_f32: bool = False, seed: Optional[int] = None) -> Iterator[Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]]:
        """mel [B,80,T] on the decoder's device (+ optional source s [B,1,480T]).  Yields one
        (int16 [B, n] | None, fp32 [B, n] | None) pair per chunk, in order."""
        hift = self.hift
        B, _, T = mel.shape
        spf = SAMPLES_PER_FRAME
        if s is None:
            f0 = hift.predict_f0(mel)
            s = hift.source_from_f0(f0, seed=seed)
        s = s.reshape(B, 1, T * spf)
        prev_tail = None
        first = True
        for own_lo, own_hi, lo, hi, last in chunk_plan(T, self.chunk, self.halo):
            wav = hift.decode(mel[:, :, lo:hi], s[:, :, lo * spf:hi * spf])
            a = (own_lo - lo) * spf
            n_emit = (own_hi - own_lo) * spf
            cur = wav[:, a:a + n_emit]
            if prev_tail is not None:
                out = pcm_tail(cur, prev_tail, self._w, self.limit, want_i16, want_f32)
            elif first and self.trim_fade:
                out = pcm_tail(cur, None, self._tw, self.limit, want_i16, want_f32)
            else:
                out = pcm_tail(cur, None, None, self.limit, want_i16, want_f32)
            first = False
            if not last:
                prev_tail = wav[:, a + n_emit:a + n_emit + self.fade]
            yield out

    @torch.no_grad()
    def decode_all(self, mel: torch.Tensor, s: Optional[torch.Tensor] = None, want_i16=True, want_f32=True,
                   seed: Optional[int] = None):
        i16s, f32s = [], []
        for i16, f32 in self.stream(mel, s, want_i16, want_f32, seed):
            if i16 is not None:
                i16s.append(i16)
            if f32 is not None:
                f32s.append(f32)
        return (torch.cat(i16s, 1) if i16s else None), (torch.cat(f32s, 1) if f32s else None)


class IncrementalDecoder:
    """Intra-sentence streaming: mel frames arrive in pieces of any size (`push`), PCM leaves chunk by
    chunk as soon as a chunk's look-ahead is there, `finish()` flushes the rest.

    The stream is EXACTLY `StreamingDecoder.stream` of the whole utterance, whatever the piece sizes: same chunk
    windows (chunk_plan), same crossfade, and the same NSF source — f0 of a frame is final once `f0_ctx` = 5 more
    frames have arrived (the predictor's receptive field: five k = 3 convs), the source of a piece continues the
    running phase of the pieces before it (fp64 sum of f0) and draws its noise by absolute sample index
    (`gnv_source_stream`).  A chunk is emitted when frames up to own_hi + 1 + halo + f0_ctx exist: 22 frames = 0.44 s
    of look-ahead at the defaults (upstream's token-chunk streaming re-decodes a 50-token context instead and accepts
    the seam; `cache_source` of HiFTGenerator.inference is the same idea)."""

    def __init__(self, hift: B200HiFT, B: int = 1, chunk: int = 100, halo: int = 16, fade: int = 480, limit: float = 0.99,
                 seed: int = 1, f0_ctx: int = 5, want_i16: bool = True, want_f32: bool = False, trim_fade: bool = False):
        if f0_ctx < 5:
            raise ValueError("f0_ctx must cover the f0 predictor's receptive field (5 frames)")
        self.hift, self.B = hift, B
        self.chunk, self.halo, self.fade, self.limit, self.seed, self.f0_ctx = chunk, halo, fade, limit, seed, f0_ctx
        self.want_i16, self.want_f32, self.trim_fade = want_i16, want_f32, trim_fade
        self._w = fade_window(fade, hift.device)
        self._tw = trim_fade_window(hift.device)
        dev = hift.device
        self._base = 0                                     # absolute index of the first frame still held
        self._mel = torch.empty(B, 80, 0, dtype=torch.float32, device=dev)
        self._s = torch.empty(B, 1, 0, dtype=torch.float32, device=dev)   # source of frames [_base, _f0_done)
        self._f0_done = 0                                  # frames whose source exists
        self._f0_sum = None
        self._next = 0                                     # next chunk to emit
        self._prev_tail = None
        self._finished = False

    @property
    def frames_in(self) -> int:
        return self._base + self._mel.shape[2]

    @torch.no_grad()
    def push(self, mel_piece: torch.Tensor):
        """mel_piece [B, 80, n >= 0] on the decoder's device.  Returns the list of (int16 | None, fp32 | None) chunks
        that became decodable (often empty)."""
        if self._finished:
            raise RuntimeError("the stream was finished")
        if mel_piece.dim() != 3 or mel_piece.shape[0] != self.B or mel_piece.shape[1] != 80:
            raise ValueError("mel_piece must be [B, 80, n]")
        self._mel = torch.cat([self._mel, mel_piece.to(self.hift.device, torch.float32)], dim=2)
        return self._advance(False)

    @torch.no_grad()
    def finish(self):
        """End of the utterance: emits the remaining chunks (the last one without look-ahead, as the offline plan)."""
        if self._finished:
            return []
        self._finished = True
        return self._advance(True)

    def _advance(self, final: bool):
        hift, spf = self.hift, SAMPLES_PER_FRAME
        T_av = self.frames_in
        # nothing on the GPU until the next chunk's look-ahead is complete (a push of one token costs a torch.cat)
        if not final and T_av - self.f0_ctx < self._next * self.chunk + self.chunk + 1 + self.halo:
            return []
        # 1. source for the frames whose f0 is final now
        f_final = T_av if final else max(self._f0_done, T_av - self.f0_ctx)
        if f_final > self._f0_done:
            w0 = max(self._base, self._f0_done - self.f0_ctx, 0)
            w1 = min(T_av, f_final + self.f0_ctx)
            f0w = hift.predict_f0(self._mel[:, :, w0 - self._base:w1 - self._base].contiguous())
            f0 = f0w[:, self._f0_done - w0:f_final - w0].contiguous()
            s_new, self._f0_sum = hift.source_stream(f0, self.seed, self._f0_done, self._f0_sum)
            self._s = torch.cat([self._s, s_new], dim=2)
            self._f0_done = f_final
        # 2. every chunk whose decode window is covered
        out = []
        while True:
            own_lo = self._next * self.chunk
            if own_lo >= T_av:
                break
            own_hi = min(T_av, own_lo + self.chunk)
            last = final and own_hi >= T_av
            lo = max(0, own_lo - self.halo)
            if last:
                hi = T_av
            else:
                hi = own_lo + self.chunk + 1 + self.halo
                if own_lo + self.chunk > T_av or hi > self._f0_done:
                    if not final:
                        break
                    hi = min(T_av, hi)                     # flushing: a full chunk near the end keeps what look-ahead exists
            wav = hift.decode(self._mel[:, :, lo - self._base:hi - self._base].contiguous(),
                              self._s[:, :, (lo - self._base) * spf:(hi - self._base) * spf].contiguous())
            a = (own_lo - lo) * spf
            n_emit = (own_hi - own_lo) * spf
            cur = wav[:, a:a + n_emit]
            if self._prev_tail is not None:
                out.append(pcm_tail(cur, self._prev_tail, self._w, self.limit, self.want_i16, self.want_f32))
            elif self._next == 0 and self.trim_fade:
                out.append(pcm_tail(cur, None, self._tw, self.limit, self.want_i16, self.want_f32))
            else:
                out.append(pcm_tail(cur, None, None, self.limit, self.want_i16, self.want_f32))
            self._prev_tail = None if last else wav[:, a + n_emit:a + n_emit + self.fade].clone()
            self._next += 1
            # frames no later chunk or f0 window needs: drop them
            keep = max(0, min(self._next * self.chunk - self.halo, self._f0_done - self.f0_ctx))
            if keep > self._base:
                d = keep - self._base
                self._mel = self._mel[:, :, d:].contiguous()
                self._s = self._s[:, :, d * spf:].contiguous()
                self._base = keep
        return out


class GraphedInference:
    """inference() + PCM tail for one fixed (B, T), captured once in a CUDA graph and replayed.

    The per-launch host cost of the ~70 kernels of a decode dominates the latency of a small chunk
    (BASELINE config 2: batch 1, first chunk); a graph removes it.  All C-ABI decode calls are
    capture-safe: they launch on the caller's current stream and neither allocate nor synchronise once
    the (B, T) plan exists, which the warm-up below guarantees.

    Two things a captured graph bakes in are kept out of harm's way:
      * the workspace — the graph's kernel arguments and tensor maps point into it for good, so the graph owns a
        PRIVATE workspace tensor (never the decoder's shared one, which later, larger eager calls regrow and free);
      * the NSF noise seed — it lives in a device word (`self.seed_dev`) that the source kernel reads at run time and a
        one-thread kernel in the graph bumps, so every replay draws fresh noise like upstream's `torch.randn_like`
        (`reseed()` sets it; replay i after reseed(s) equals an eager `inference(seed=s + i)`)."""

    def __init__(self, hift: B200HiFT, B: int, T: int, emit_frames: Optional[int] = None, seed: int = 1,
                 limit: float = 0.99, want_i16: bool = True, trim_fade: bool = False):
        dev = hift.device
        self.hift, self.B, self.T = hift, B, T
        n_emit = (emit_frames if emit_frames is not None else T) * SAMPLES_PER_FRAME
        self.mel = torch.zeros(B, 80, T, dtype=torch.float32, device=dev)
        self.wav = torch.empty(B, T * SAMPLES_PER_FRAME, dtype=torch.float32, device=dev)
        self.src = torch.empty(B, 1, T * SAMPLES_PER_FRAME, dtype=torch.float32, device=dev)
        self.pcm = torch.empty(B, n_emit, dtype=torch.int16 if want_i16 else torch.float32, device=dev)
        self.seed_dev = torch.full((1,), int(seed), dtype=torch.int64, device=dev)
        self._ws = hift.new_workspace(B, T)
        fw = trim_fade_window(dev) if trim_fade else None

        def body():
            hift.inference(self.mel, out=self.wav, source_out=self.src, workspace=self._ws, seed_dev=self.seed_dev)
            pcm_tail(self.wav[:, :n_emit], None, fw, limit, want_i16=want_i16, want_f32=not want_i16,
                     out_i16=self.pcm if want_i16 else None, out_f32=None if want_i16 else self.pcm)

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):                      # builds the plan, warms the kernels
                body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            body()
        self.reseed(seed)                           # the warm-up calls advanced the device seed

    def reseed(self, seed: int) -> None:
        self.seed_dev.fill_(int(seed))

    @torch.no_grad()
    def __call__(self, mel: torch.Tensor) -> torch.Tensor:
        """mel [B,80,T] (device or pinned host) -> the graph's PCM buffer [B, n_emit] (overwritten by the next call)."""
        self.mel.copy_(mel, non_blocking=True)
        self.graph.replay()
        return self.pcm
