"""numpy oracle for the streaming tail (TEST INFRASTRUCTURE ONLY; see oracle/hift_ref.py header).

The reference service never crossfades, clamps or packs int16: it ships one float32 chunk per
sentence (/root/reference/services/tts/core/synthesizer.py:352-357,
/root/reference/services/tts/server.py:150-155).  north_star item (3) asks for those three steps,
so their semantics are DEFINED HERE (SURVEY.md §8c "Oracles this repo must define"):

* int16 pack : i16 = clip(rint(float32(x) * float32(32767)), -32768, 32767); round-half-to-even,
               one fp32 multiply, no FMA.  This is libsndfile's float->PCM_16 rule, i.e. what
               ``soundfile.write(subtype='PCM_16')`` (the reference's audio library,
               /root/reference/pyproject.toml:26) would emit.
* crossfade  : raised cosine w = (cos(linspace(pi, 0, n)) + 1) / 2 (the curve of upstream
               S3Token2Wav.trim_fade); out = prev_tail * (1 - w) + cur_head * w in fp32, each
               operation individually rounded.  With no prev_tail the head is cur * w
               (= trim_fade when w = [0]*480 || cosine).
* clamp      : clamp(+-limit), limit = 0.99 (upstream HiFTGenerator.audio_limit).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def fade_window(n: int = 480) -> np.ndarray:
    """(cos(linspace(pi, 0, n)) + 1) / 2 in float32, computed exactly as torch does on CPU."""
    import torch

    return ((torch.cos(torch.linspace(torch.pi, 0, n, dtype=torch.float32)) + 1) / 2).numpy().copy()


def trim_fade_window() -> np.ndarray:
    w = np.zeros(960, dtype=F32)
    w[480:] = fade_window(480)
    return w


def pack_i16(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=F32)
    y = np.rint(x * F32(32767.0))            # fp32 multiply, round half to even
    return np.clip(y, -32768.0, 32767.0).astype(np.int16)


def pcm_tail(cur: np.ndarray, prev_tail, fade_w, limit: float = 0.99):
    """cur: [rows, n] float32.  prev_tail: [rows, fade] or None.  fade_w: [fade] or None.
    Returns (float32 [rows, n], int16 [rows, n])."""
    cur = np.asarray(cur, dtype=F32)
    out = cur.copy()
    if fade_w is not None:
        w = np.asarray(fade_w, dtype=F32)
        f = min(w.shape[0], cur.shape[-1])
        w = w[:f]
        if prev_tail is None:
            out[..., :f] = cur[..., :f] * w
        else:
            p = np.asarray(prev_tail, dtype=F32)[..., :f]
            a = p * (F32(1.0) - w)
            b = cur[..., :f] * w
            out[..., :f] = a + b
    lim = F32(limit)
    out = np.minimum(np.maximum(out, -lim), lim)
    return out, pack_i16(out)


def chunk_plan(T: int, chunk: int = 100, halo: int = 16):
    """Frame windows for chunked decode.  Chunk c owns frames [c*chunk, min(T,(c+1)*chunk)); it is
    decoded on frames [lo, hi) = owned +- halo plus ONE extra look-ahead frame whose 480 samples are
    held back and crossfaded into the next chunk's head.  Yields (own_lo, own_hi, lo, hi, last)."""
    c = 0
    while c * chunk < T:
        own_lo = c * chunk
        own_hi = min(T, own_lo + chunk)
        last = own_hi >= T
        lo = max(0, own_lo - halo)
        hi = T if last else min(T, own_hi + 1 + halo)
        yield own_lo, own_hi, lo, hi, last
        c += 1


def stream_decode_ref(decode_fn, mel, s, chunk: int = 100, halo: int = 16, fade: int = 480,
                      limit: float = 0.99):
    """Chunked decode + crossfade with `decode_fn(mel[B,80,t], s[B,1,480t]) -> wav[B,480t]` (numpy
    in/out).  Returns (float32 [B, 480T], int16 [B, 480T]).  This is the definition the CUDA
    streaming path is checked against with the SAME decode_fn outputs (bit-exact), and against
    the full-utterance oracle decode (tolerance)."""
    B, _, T = mel.shape
    spf = 480
    w = fade_window(fade)
    out_f = np.zeros((B, T * spf), dtype=F32)
    out_i = np.zeros((B, T * spf), dtype=np.int16)
    prev_tail = None
    for own_lo, own_hi, lo, hi, last in chunk_plan(T, chunk, halo):
        wav = np.asarray(decode_fn(mel[:, :, lo:hi], s[:, :, lo * spf:hi * spf]), dtype=F32)
        a = (own_lo - lo) * spf
        n_emit = (own_hi - own_lo) * spf
        cur = wav[:, a:a + n_emit]
        f, i16 = pcm_tail(cur, prev_tail, w if prev_tail is not None else None, limit)
        out_f[:, own_lo * spf:own_hi * spf] = f
        out_i[:, own_lo * spf:own_hi * spf] = i16
        if not last:
            prev_tail = wav[:, a + n_emit:a + n_emit + fade].copy()
    return out_f, out_i


def mulaw_encode(pcm_i16: np.ndarray) -> np.ndarray:
    """G.711 mu-law companding of int16 PCM -> uint8 (the wire format of the reference's `phone`
    extra, /root/reference/pyproject.toml:55-57).  Restates CPython's audioop.lin2ulaw(width=2)
    (Modules/audioop.c st_14linear2ulaw: 14-bit magnitude, bias 33, clip 8159, 8 segments); tests/test_tail_oracle.py
    pins it against audioop itself on all 65536 inputs."""
    x = np.asarray(pcm_i16, dtype=np.int16).astype(np.int32) >> 2          # arithmetic shift: 14-bit sample
    neg = x < 0
    mag = np.where(neg, -x, x)
    mag = np.minimum(mag, 8159) + 33
    ends = np.array([0x3F, 0x7F, 0xFF, 0x1FF, 0x3FF, 0x7FF, 0xFFF, 0x1FFF], dtype=np.int32)
    seg = np.searchsorted(ends, mag, side="left").astype(np.int32)        # first segment whose end >= mag
    uval = (seg << 4) | ((mag >> (seg + 1)) & 0xF)
    uval = np.where(seg >= 8, 0x7F, uval)
    mask = np.where(neg, 0x7F, 0xFF)
    return (uval ^ mask).astype(np.uint8)
