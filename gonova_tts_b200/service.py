"""Service-side glue: how the reference's StreamingSynthesizer gets the B200 decoder.

install() is the whole integration: call it in StreamingSynthesizer.load() right after
`self.model = ChatterboxTTS.from_pretrained(device=self.device)` (reference
services/tts/core/synthesizer.py:185) and before the warm-up loop (:199-207).  Nothing in server.py,
queue_manager.py or voice_manager.py changes; the wire format stays float32 LE PCM
(synthesizer.py:352-357, server.py:152) unless the caller asks for int16 via PcmSink.

Beyond the plain swap (all optional):
  * chunk_tap(decoder, sink): while active on the calling thread, the engine's ONE `mel2wav.inference(...)` call per
    sentence (synthesizer.py:344-350) hands its PCM to `sink` chunk by chunk, 2 s at a time, as soon as each chunk is
    decoded — the first after ~0.6 ms instead of after the whole sentence — and still returns the whole waveform.
    That turns the one-chunk-per-sentence loop (synthesizer.py:296-321) into intra-sentence streaming without touching
    the engine; examples/patched_synthesizer.py is the patched `_generate_sentence`.
  * PcmSink.to_memoryview(): the chunk as a view of pinned host memory for `websocket.send_bytes` (server.py:280):
    one D2H copy, no `.numpy().tobytes()` copy behind it.
  * decoder_stats() / patch_get_stats(): the decoder's counters under `get_stats()["decoder"]`
    (synthesizer.py:411-420)."""
from __future__ import annotations

import contextlib
import os
import threading
from typing import Callable, Optional

import torch

from .decoder import B200HiFT, SAMPLES_PER_FRAME, fade_window, mulaw_encode, pcm_tail, trim_fade_window
from .streaming import chunk_plan


def install(model, dtype: Optional[str] = None, device=None, bucket_frames: int = 8,
            max_frames: int = 3000) -> B200HiFT:
    """Swap `model.s3gen.mel2wav` for a B200HiFT built from its weights.  Returns the new decoder.
    Raises (never falls back) if the CUDA library is missing or the device is not a B200.

    `s3gen` is an nn.Module and `mel2wav` a registered child, so the replacement has to be an nn.Module too
    (B200HiFT is one, with no parameters); the old module is dropped.

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


def install_flow(model, dtype: Optional[str] = None, device=None, front: bool = True):
    """Swap `model.s3gen.flow.decoder` (upstream CausalConditionalCFM: the estimator + the Euler loop that turn the encoder's
    `mu` into mel frames) for a B200Flow built from its estimator's weights and its fixed noise buffer.  The
    engine keeps calling `self.decoder(mu=..., mask=..., spks=..., cond=..., n_timesteps=10)` and gets `(mel, None)` back.
    With `front` (and a flow module that holds `input_embedding` / `encoder` / `encoder_proj` / `spk_embed_affine_layer`),
    `model.s3gen.flow.inference(token=..., prompt_token=..., prompt_feat=..., embedding=..., finalize=...)` is rebound too:
    token embedding, Conformer encoder and projections run on B200FlowFront, so tokens -> mel never leaves this library.
    Same placement as install(): right after ChatterboxTTS.from_pretrained (synthesizer.py:185)."""
    from .flow import B200Flow

    dtype = dtype or os.environ.get("GONOVA_FLOW_DTYPE", "bf16")
    old = model.s3gen.flow.decoder
    if device is None:
        try:
            device = next(old.parameters()).device
        except StopIteration:
            device = "cuda:0"
        if torch.device(device).type != "cuda":
            device = "cuda:0"
    new = B200Flow(old.estimator.state_dict(), device=device, dtype=dtype)
    noise = getattr(old, "rand_noise", None)
    if isinstance(noise, torch.Tensor):                 # the same noise the engine would have used: same mel
        new.rand_noise = noise.detach().to(new.device, torch.float32).contiguous()
    model.s3gen.flow.decoder = new
    flow = model.s3gen.flow
    if front and all(hasattr(flow, n) for n in ("input_embedding", "encoder", "encoder_proj", "spk_embed_affine_layer")):
        from .flow_front import B200FlowFront, B200FlowInference

        sd = {k: v for k, v in flow.state_dict().items() if not k.startswith("decoder.")}
        whole = B200FlowInference(front=B200FlowFront(sd, device=new.device, dtype=dtype), decoder=new)
        flow.inference = whole.inference                 # an instance attribute: shadows the class's method
        # (a plain attribute, not a registered child: `whole.decoder` is `new`, and a module cycle would make
        # s3gen.eval() / .to() / .state_dict() recurse forever)
        object.__setattr__(new, "flow_inference", whole)
    return new


# -------------------------------------------------------------------------------------------------
# intra-sentence streaming through the engine's own call
# -------------------------------------------------------------------------------------------------
_tap = threading.local()


@contextlib.contextmanager
def chunk_tap(decoder: B200HiFT, sink: Callable[[object, int, bool], None], fmt: str = "f32", chunk: int = 100,
              halo: int = 16, fade: int = 480, trim_fade: bool = True, as_memoryview: bool = False):
    """While active ON THIS THREAD, `decoder.inference(speech_feat [1,80,T], cache_source)` — the call the engine makes
    once per sentence — decodes in `chunk`-frame pieces and calls `sink(pcm, chunk_id, is_last)` for each as soon as it
    is on the host (`pcm`: bytes, or a memoryview of pinned memory valid until the next-but-one chunk).  The call
    still returns (wav, source) for the whole sentence: wav is the concatenation of the streamed float chunks, i.e. the
    chunked decode (16-frame halos, 480-sample raised-cosine crossfades), which differs from the single-shot decode
    only inside the fp32 tolerance (tests/test_gpu_decode.py, 60 s case).

    trim_fade: upstream S3Token2Wav multiplies the first 960 samples by `trim_fade` AFTER mel2wav.inference returns;
    a streamed first chunk has left by then, so the tap applies it to the first chunk itself (and un-applies nothing:
    the returned wav carries it too; multiplying again upstream only squares the 20 ms fade-in)."""
    if getattr(_tap, "cfg", None) is not None:
        raise RuntimeError("chunk_tap is already active on this thread")
    sinkbuf = PcmSink(decoder.device, max_samples=(chunk + 1) * SAMPLES_PER_FRAME, fmt=fmt, buffers=3)
    _tap.cfg = dict(decoder=decoder, sink=sink, chunk=chunk, halo=halo, fade=fade, trim_fade=trim_fade,
                    pcm=sinkbuf, mv=as_memoryview)
    decoder._tap_hook = _tapped_inference
    try:
        yield
    finally:
        _tap.cfg = None


def _tapped_inference(decoder: B200HiFT, mel: torch.Tensor, cache_source, seed):
    """Called by B200HiFT.inference when a tap is active on the calling thread for this decoder; returns None to fall
    through to the plain path (another thread, another decoder, a batch)."""
    cfg = getattr(_tap, "cfg", None)
    if cfg is None or cfg["decoder"] is not decoder or mel.shape[0] != 1:
        return None
    T = mel.shape[2]
    spf = SAMPLES_PER_FRAME
    f0 = decoder.predict_f0(mel)
    s = decoder.source_from_f0(f0, seed=seed)
    if cache_source is not None and cache_source.numel():
        n = min(cache_source.shape[-1], T * spf)
        s[:, :, :n] = cache_source.reshape(1, 1, -1)[:, :, :n]
    w = fade_window(cfg["fade"], decoder.device)
    tw = trim_fade_window(decoder.device)
    pcm: PcmSink = cfg["pcm"]
    pieces = []
    prev_tail = None
    for cid, (own_lo, own_hi, lo, hi, last) in enumerate(chunk_plan(T, cfg["chunk"], cfg["halo"])):
        wav = decoder.decode(mel[:, :, lo:hi].contiguous(), s[:, :, lo * spf:hi * spf].contiguous())
        a, n_emit = (own_lo - lo) * spf, (own_hi - own_lo) * spf
        cur = wav[:, a:a + n_emit]
        if prev_tail is not None:
            _, f32 = pcm_tail(cur, prev_tail, w, decoder.audio_limit, want_i16=False, want_f32=True)
        elif cfg["trim_fade"]:
            _, f32 = pcm_tail(cur, None, tw, decoder.audio_limit, want_i16=False, want_f32=True)
        else:
            _, f32 = pcm_tail(cur, None, None, decoder.audio_limit, want_i16=False, want_f32=True)
        prev_tail = None if last else wav[:, a + n_emit:a + n_emit + cfg["fade"]].clone()
        pieces.append(f32)
        out = pcm.to_memoryview(f32, clamp=False) if cfg["mv"] else pcm.to_bytes(f32, clamp=False)
        cfg["sink"](out, cid, last)
    return torch.cat(pieces, dim=1), s


# -------------------------------------------------------------------------------------------------
# device float32 wav -> host bytes
# -------------------------------------------------------------------------------------------------
class PcmSink:
    """Device float32 wav -> host bytes for the WebSocket, through pinned staging buffers.
    fmt 'f32' reproduces the reference's `audio.astype(np.float32).tobytes()` bytes exactly (the clamp to +-limit is
    the decoder's own `audio_limit`, a no-op on its output);
    fmt 'i16' is the opt-in int16 pack (clamp + round-half-even, one kernel); fmt 'mulaw' adds G.711 companding
    (8 bits per sample, the telephony clients of the reference's `phone` extra).

    to_bytes() returns an owned `bytes` (one host copy behind the D2H copy, like the reference's `.tobytes()`);
    to_memoryview() returns a read-only view of the pinned buffer itself — what `websocket.send_bytes`
    (server.py:280) can take without another copy.  The sink rotates over `buffers` staging buffers, so a view stays
    valid until `buffers - 1` further calls have been made."""

    def __init__(self, device, max_samples: int = 24000 * 60, fmt: str = "f32", limit: float = 0.99, buffers: int = 2):
        if fmt not in ("f32", "i16", "mulaw"):
            raise ValueError("fmt must be 'f32', 'i16' or 'mulaw'")
        if buffers < 1:
            raise ValueError("buffers must be positive")
        self.fmt, self.limit = fmt, limit
        self.device = torch.device(device)
        self._dtype = {"f32": torch.float32, "i16": torch.int16, "mulaw": torch.uint8}[fmt]
        self._hosts = [torch.empty(max_samples, dtype=self._dtype).pin_memory() for _ in range(buffers)]
        self._k = 0
        self.bytes_out = 0

    def _stage(self, wav: torch.Tensor, trim_fade: bool, clamp: bool) -> torch.Tensor:
        wav = wav.reshape(1, -1)
        n = wav.shape[1]
        self._k = (self._k + 1) % len(self._hosts)
        if n > self._hosts[self._k].numel():
            self._hosts[self._k] = torch.empty(n, dtype=self._dtype).pin_memory()
        host = self._hosts[self._k][:n]
        if self.fmt == "f32" and not trim_fade and not clamp:
            src = wav.to(torch.float32)                     # already the decoder's clamped output: copy as is
        else:
            fw = trim_fade_window(self.device) if trim_fade else None
            i16, f32 = pcm_tail(wav, None, fw, self.limit, want_i16=self.fmt != "f32", want_f32=self.fmt == "f32")
            src = f32 if self.fmt == "f32" else (i16 if self.fmt == "i16" else mulaw_encode(i16))
        host.copy_(src.reshape(-1), non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        self.bytes_out += host.numel() * host.element_size()
        return host

    @torch.no_grad()
    def to_bytes(self, wav: torch.Tensor, trim_fade: bool = False, clamp: bool = True) -> bytes:
        return self._stage(wav, trim_fade, clamp).numpy().tobytes()

    @torch.no_grad()
    def to_memoryview(self, wav: torch.Tensor, trim_fade: bool = False, clamp: bool = True) -> memoryview:
        host = self._stage(wav, trim_fade, clamp)
        return memoryview(host.numpy()).cast("B").toreadonly()


# -------------------------------------------------------------------------------------------------
# counters for StreamingSynthesizer.get_stats()
# -------------------------------------------------------------------------------------------------
def decoder_stats(decoder: B200HiFT) -> dict:
    """What `get_stats()["decoder"]` carries: calls / frames / audio seconds decoded since start-up and the state of
    the launch-plan cache (gnv_plan_stats)."""
    st = dict(decoder.counters)
    st["audio_seconds"] = st["frames"] / 50.0
    st["dtype"] = decoder.dtype
    st["plans"] = decoder.plan_stats()
    return st


def patch_get_stats(synthesizer, decoder: B200HiFT) -> None:
    """`synthesizer.get_stats()` (services/tts/core/synthesizer.py:411-420) gains a "decoder" entry; the service's
    /metrics and /health handlers (server.py) pick it up unchanged."""
    orig = synthesizer.get_stats

    def get_stats():
        stats = orig()
        stats["decoder"] = decoder_stats(decoder)
        return stats

    synthesizer.get_stats = get_stats
