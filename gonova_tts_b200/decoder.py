"""B200HiFT — the drop-in for the object the service's engine keeps at `model.s3gen.mel2wav`.

Reference boundary (reference paths relative to the gonova-tts checkout):
  services/tts/core/synthesizer.py:185      ChatterboxTTS.from_pretrained(...)   <- install after this line
  services/tts/core/synthesizer.py:344-350  model.generate(...) -> S3Gen.inference -> mel2wav.inference(...)
This class mirrors the upstream HiFTGenerator surface the engine uses on that path:
  inference(speech_feat[B,80,T], cache_source[B,1,S]) -> (wav[B,480T], source[B,1,480T])
  decode(x, s) -> wav;   f0_predictor(mel) -> f0[B,T];   sampling_rate / istft_params / audio_limit.
All arithmetic runs in libgonova_hift.so (hand-written sm_100a kernels) through the C ABI in
include/gonova_hift.h; torch only owns device memory and streams.  There is no fallback path."""
from __future__ import annotations

import ctypes as C
import threading
from typing import Dict, Optional, Tuple

import torch

from . import _cabi
from .weights import fold_state_dict, random_state_dict

SAMPLES_PER_FRAME = 480


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream_ptr(device: torch.device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _F0Predictor:
    """Callable mirror of upstream ConvRNNF0Predictor: mel [B,80,T] -> f0 [B,T]."""

    def __init__(self, owner: "B200HiFT"):
        self._owner = owner

    def __call__(self, mel: torch.Tensor, lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self._owner.predict_f0(mel, lengths)


class _SourceModule:
    """Callable mirror of upstream SourceModuleHnNSF on the x480-upsampled f0 path: f0 [B,T] -> s [B,1,480T]."""

    def __init__(self, owner: "B200HiFT"):
        self._owner = owner

    def __call__(self, f0: torch.Tensor, seed: Optional[int] = None, phase_vec=None, noise=None) -> torch.Tensor:
        return self._owner.source_from_f0(f0, seed=seed, phase_vec=phase_vec, noise=noise)


class B200HiFT(torch.nn.Module):
    """An nn.Module only so that it can sit where the engine keeps its vocoder: `s3gen.mel2wav` is a registered child
    module, and nn.Module.__setattr__ refuses anything else there.  It has no parameters or buffers (the weights live
    re-packed inside the C handle); `state_dict()` is empty and `S3Gen.load_state_dict` reaches `_load_from_state_dict`
    below, which re-creates the handle from the `mel2wav.*` entries."""
    sampling_rate = 24000
    istft_params = {"n_fft": 16, "hop_len": 4}
    audio_limit = 0.99

    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda:0", dtype: str = "bf16",
                 simt_conv: bool = False, precise_act: bool = False, tc_v1: bool = False, prefix: str = "",
                 bucket_frames: int = 0):
        """bucket_frames > 1: `inference()` rounds T up to a multiple of it and masks the tail through `lengths`, so a
        service that sees a new sentence length on every call (services/tts/server.py:118-182) keeps hitting the
        handle's launch-plan cache (gnv_plan_stats).  Results for the first T frames do not change."""
        super().__init__()
        if dtype not in _cabi.DTYPE:
            raise ValueError(f"dtype must be one of {sorted(_cabi.DTYPE)}")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("B200HiFT runs on a CUDA device only (no CPU fallback)")
        if not torch.cuda.is_available():
            raise RuntimeError("B200HiFT needs a CUDA device (B200, sm_100a); none is visible")
        self.dtype = dtype
        self._lib = _cabi.load()
        self._flags = ((_cabi.FLAG_SIMT_CONV if simt_conv else 0) | (_cabi.FLAG_PRECISE_ACT if precise_act else 0) |
                       (_cabi.FLAG_TC_V1 if tc_v1 else 0))
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", dev_index)
        self._h = None
        self._h = self._create_handle(state_dict, prefix)
        self._ws: Optional[torch.Tensor] = None
        self._ws_pinned = False        # a CUDA graph was captured over self._ws: it may never be freed or moved
        self._ws_retired = []          # ... and if a larger one is needed anyway, the captured one is kept alive here
        self.bucket_frames = int(bucket_frames)
        self._lock = threading.Lock()
        self._seed = 0
        self._tap_hook = None          # service.chunk_tap: streams the PCM of an inference() call chunk by chunk
        self.counters = {"calls": 0, "frames": 0}      # surfaced by service.decoder_stats (get_stats()["decoder"])
        self.f0_predictor = _F0Predictor(self)
        self.m_source = _SourceModule(self)

    def _create_handle(self, state_dict: Dict[str, torch.Tensor], prefix: str = ""):
        folded = fold_state_dict(state_dict, prefix=prefix)
        names = sorted(folded)
        arr = (_cabi.GnvWeight * len(names))()
        keep = []
        for i, n in enumerate(names):
            t = folded[n].contiguous()
            keep.append(t)
            arr[i].name = n.encode()
            arr[i].data = C.cast(C.c_void_p(t.data_ptr()), C.POINTER(C.c_float))
            arr[i].ndim = t.dim()
            for d in range(t.dim()):
                arr[i].shape[d] = t.shape[d]
        h = C.c_void_p()
        rc = self._lib.gnv_create(arr, len(names), self.device.index, _cabi.DTYPE[self.dtype], self._flags, C.byref(h))
        _cabi.check(rc, None, "gnv_create")
        del keep
        return h

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                              error_msgs):
        """`S3Gen.load_state_dict(...)` with `mel2wav.*` entries (upstream key names, weight-norm folded or not):
        the decoder handle is rebuilt from them.  No entries under the prefix: nothing happens (strict: reported as
        missing, like a module whose parameters are absent)."""
        sub = {k[len(prefix):]: v for k, v in state_dict.items() if k.startswith(prefix)}
        if not sub:
            if strict:
                missing_keys.append(prefix + "*")
            return
        try:
            with self._lock:
                new = self._create_handle(sub)
                old, self._h = self._h, new
                if old is not None and old.value:
                    torch.cuda.synchronize(self.device)
                    self._lib.gnv_destroy(old)
        except Exception as e:   # load_state_dict reports, it does not raise from inside the walk
            error_msgs.append(f"{prefix}: {e}")

    # -- construction helpers ---------------------------------------------------------------------
    @classmethod
    def from_module(cls, mel2wav, device=None, dtype: str = "bf16", **kw) -> "B200HiFT":
        """Build from the engine's existing `s3gen.mel2wav` nn.Module (weights are copied, folded)."""
        sd = mel2wav.state_dict()
        if device is None:
            try:
                device = next(mel2wav.parameters()).device
            except StopIteration:
                device = "cuda:0"
            if torch.device(device).type != "cuda":
                device = "cuda:0"
        return cls(sd, device=device, dtype=dtype, **kw)

    @classmethod
    def random_init(cls, seed: int = 0, corners: bool = False, **kw) -> "B200HiFT":
        return cls(random_state_dict(seed, corners), **kw)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                self._lib.gnv_destroy(h)
            except Exception:
                pass
            self.__dict__["_h"] = None          # (not nn.Module.__setattr__: it may be half torn down at interpreter exit)

    # nn.Module-ish no-ops so the engine's `.to(device).eval()` chains keep working
    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    # -- internals --------------------------------------------------------------------------------
    def workspace_bytes(self, B: int, T: int) -> int:
        n = C.c_size_t()
        _cabi.check(self._lib.gnv_workspace_bytes(self._h, B, T, C.byref(n)), self._h, "gnv_workspace_bytes")
        return n.value

    def _workspace(self, B: int, T: int, workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The scratch tensor a (B, T) call runs in: the caller's `workspace` if given (checked), else the decoder's own,
        grown on demand.  A workspace a CUDA graph was captured over (`pin_workspace`) is never freed: the graph's
        kernel arguments and tensor maps point into it for good."""
        need = self.workspace_bytes(B, T)
        if workspace is not None:
            if (workspace.dtype != torch.uint8 or workspace.device != self.device or not workspace.is_contiguous()
                    or workspace.numel() < need + 1024):
                raise ValueError(f"workspace must be a contiguous uint8 tensor on {self.device} of at least "
                                 f"{need + 1024} bytes (workspace_bytes(B, T) + 1024)")
            return workspace
        if self._ws is None or self._ws.numel() < need + 1024:
            # every launch plan is tied to the workspace address: grow by at least half so that a run of slightly
            # longer sentences does not rebuild the plans each time
            grow = 0 if self._ws is None else self._ws.numel() + self._ws.numel() // 2
            if self._ws_pinned:
                self._ws_retired.append(self._ws)
                self._ws_pinned = False
            self._ws = None
            self._ws = torch.empty(max(need + 1024, grow), dtype=torch.uint8, device=self.device)
        return self._ws

    def pin_workspace(self) -> None:
        """Called by whoever captures a CUDA graph over calls that used the decoder's own workspace."""
        with self._lock:
            if self._ws is not None:
                self._ws_pinned = True

    def new_workspace(self, B: int, T: int) -> torch.Tensor:
        """A private workspace for (B, T) to pass as `workspace=` (e.g. one per captured CUDA graph)."""
        return torch.empty(self.workspace_bytes(B, T) + 1024, dtype=torch.uint8, device=self.device)

    def reserve(self, B: int, T: int) -> None:
        """Allocate the workspace for the largest (B, T) the caller will send, once (e.g. at service start-up)."""
        with self._lock:
            self._workspace(B, T)

    @staticmethod
    def _aligned(ws: torch.Tensor) -> Tuple[int, int]:
        base = ws.data_ptr()
        off = (-base) % 1024
        return base + off, ws.numel() - off

    def _check_in(self, t: torch.Tensor, name: str) -> torch.Tensor:
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{name} must be a torch.Tensor")
        if t.device != self.device:
            raise RuntimeError(f"{name} is on {t.device}, the decoder is on {self.device}")
        return t.to(torch.float32).contiguous()

    def _check_out(self, t: Optional[torch.Tensor], numel: int, name: str) -> Optional[torch.Tensor]:
        """Caller-supplied output buffers go to the C ABI by raw pointer: anything but a contiguous fp32 tensor of
        exactly the right size on the decoder's device would be written out of bounds or read back as garbage."""
        if t is None:
            return None
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{name} must be a torch.Tensor")
        if t.dtype != torch.float32 or t.device != self.device or not t.is_contiguous() or t.numel() != numel:
            raise ValueError(f"{name} must be a contiguous float32 tensor on {self.device} with {numel} elements "
                             f"(got {t.dtype}, {t.device}, contiguous={t.is_contiguous()}, {t.numel()} elements)")
        return t

    def _next_seed(self, seed: Optional[int]) -> int:
        if seed is not None:
            return int(seed)
        with self._lock:
            self._seed += 1
            return self._seed

    def _lengths(self, lengths, B):
        if lengths is None:
            return None
        lengths = torch.as_tensor(lengths, dtype=torch.int32, device=self.device).contiguous()
        if lengths.shape != (B,):
            raise ValueError("lengths must have shape [B]")
        return lengths

    # -- the upstream surface ---------------------------------------------------------------------
    @torch.no_grad()
    def predict_f0(self, mel: torch.Tensor, lengths=None) -> torch.Tensor:
        mel = self._check_in(mel, "mel")
        B, Cm, T = mel.shape
        if Cm != 80:
            raise ValueError("mel must be [B, 80, T]")
        lengths = self._lengths(lengths, B)
        with self._lock:
            ws = self._workspace(B, T)
            p, n = self._aligned(ws)
            f0 = torch.empty(B, T, dtype=torch.float32, device=self.device)
            rc = self._lib.gnv_f0(self._h, _ptr(mel), _ptr(lengths), B, T, _ptr(f0), C.c_void_p(p), n,
                                  _stream_ptr(self.device))
            _cabi.check(rc, self._h, "gnv_f0")
        return f0

    @torch.no_grad()
    def source_from_f0(self, f0: torch.Tensor, seed: Optional[int] = None, phase_vec=None, noise=None) -> torch.Tensor:
        f0 = self._check_in(f0, "f0")
        B, T = f0.shape
        if phase_vec is not None:
            phase_vec = self._check_in(phase_vec, "phase_vec").reshape(B, 9)
        if noise is not None:
            noise = self._check_in(noise, "noise").reshape(B, 9, T * SAMPLES_PER_FRAME)
        seed = self._next_seed(seed)
        s = torch.empty(B, 1, T * SAMPLES_PER_FRAME, dtype=torch.float32, device=self.device)
        rc = self._lib.gnv_source(self._h, _ptr(f0), B, T, C.c_uint64(seed), _ptr(phase_vec), _ptr(noise), _ptr(s),
                                  _stream_ptr(self.device))
        _cabi.check(rc, self._h, "gnv_source")
        return s

    @torch.no_grad()
    def source_stream(self, f0: torch.Tensor, seed: int, frame0: int,
                      f0_sum: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """The source for frames [frame0, frame0 + T) of a longer utterance (gnv_source_stream): f0 [B, T] of those
        frames, `f0_sum` [B] float64 = the running sum of f0 over the frames before them (None at frame 0).  Returns
        (s [B, 1, 480 T], the updated running sum).  Pieces of any size concatenate to `source_from_f0` of the whole."""
        f0 = self._check_in(f0, "f0")
        B, T = f0.shape
        if f0_sum is not None:
            f0_sum = f0_sum.to(device=self.device, dtype=torch.float64).contiguous()
            if f0_sum.shape != (B,):
                raise ValueError("f0_sum must have shape [B]")
        s = torch.empty(B, 1, T * SAMPLES_PER_FRAME, dtype=torch.float32, device=self.device)
        out = torch.empty(B, dtype=torch.float64, device=self.device)
        rc = self._lib.gnv_source_stream(self._h, _ptr(f0), B, T, C.c_uint64(seed), C.c_int64(frame0 * SAMPLES_PER_FRAME),
                                         _ptr(f0_sum), _ptr(s), _ptr(out), _stream_ptr(self.device))
        _cabi.check(rc, self._h, "gnv_source_stream")
        return s, out

    @torch.no_grad()
    def decode(self, x: torch.Tensor, s: torch.Tensor, lengths=None, out: Optional[torch.Tensor] = None,
               workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
        """HiFTGenerator.decode(x=mel [B,80,T], s=source [B,1,480T]) -> wav [B,480T]."""
        x = self._check_in(x, "x")
        s = self._check_in(s, "s")
        B, Cm, T = x.shape
        if Cm != 80:
            raise ValueError("x must be [B, 80, T]")
        if s.numel() != B * T * SAMPLES_PER_FRAME:
            raise ValueError(f"s must hold B*480*T = {B * T * SAMPLES_PER_FRAME} samples, got {s.numel()}")
        lengths = self._lengths(lengths, B)
        out = self._check_out(out, B * T * SAMPLES_PER_FRAME, "out")
        wav = out if out is not None else torch.empty(B, T * SAMPLES_PER_FRAME, dtype=torch.float32, device=self.device)
        self.counters["calls"] += 1
        self.counters["frames"] += B * T
        with self._lock:
            ws = self._workspace(B, T, workspace)
            p, n = self._aligned(ws)
            rc = self._lib.gnv_decode(self._h, _ptr(x), _ptr(s), _ptr(lengths), B, T, _ptr(wav), C.c_void_p(p), n,
                                      _stream_ptr(self.device))
            _cabi.check(rc, self._h, "gnv_decode")
        return wav

    @torch.no_grad()
    def inference(self, speech_feat: torch.Tensor, cache_source: Optional[torch.Tensor] = None, lengths=None,
                  seed: Optional[int] = None, out: Optional[torch.Tensor] = None,
                  source_out: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None,
                  seed_dev: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """HiFTGenerator.inference(speech_feat, cache_source) -> (wav [B,480T], source [B,1,480T]).

        seed_dev: int64 CUDA tensor holding the NSF seed(s) (gnv_inference_dseed); `seed` is ignored then.
          one element: the source kernel reads it at RUN time and a one-thread kernel adds 1 to it afterwards, so a
            CUDA graph captured over this call draws fresh noise at every replay;
          B elements (B > 1): one seed per row; row b gets exactly the source `inference(mel[b:b+1], seed=seed_dev[b])`
            gets, whatever batch it rides in (the micro-batcher's per-request seeds)."""
        mel = self._check_in(speech_feat, "speech_feat")
        B, Cm, T = mel.shape
        if Cm != 80:
            raise ValueError("speech_feat must be [B, 80, T]")
        hook = self._tap_hook
        if hook is not None and out is None and source_out is None and seed_dev is None and lengths is None:
            tapped = hook(self, mel, cache_source, seed)
            if tapped is not None:
                return tapped
        self.counters["calls"] += 1
        self.counters["frames"] += B * T
        cache_len = 0
        if cache_source is not None and cache_source.numel() != 0:
            cache_source = self._check_in(cache_source, "cache_source").reshape(B, -1)
            cache_len = cache_source.shape[1]
        else:
            cache_source = None
        if seed_dev is not None:
            if (seed_dev.dtype != torch.int64 or seed_dev.device != self.device or not seed_dev.is_contiguous()
                    or seed_dev.numel() not in (1, B)):
                raise ValueError(f"seed_dev must be a contiguous int64 tensor on {self.device} with 1 or B elements")
        else:
            seed = self._next_seed(seed)
        T_true = T
        if self.bucket_frames > 1 and out is None and source_out is None and T % self.bucket_frames:
            T = -(-T // self.bucket_frames) * self.bucket_frames
            mel = torch.nn.functional.pad(mel, (0, T - T_true))
            if lengths is None:
                lengths = [T_true] * B
        lengths = self._lengths(lengths, B)
        L = T * SAMPLES_PER_FRAME
        out = self._check_out(out, B * L, "out")
        source_out = self._check_out(source_out, B * L, "source_out")
        wav = out if out is not None else torch.empty(B, L, dtype=torch.float32, device=self.device)
        src = source_out if source_out is not None else torch.empty(B, 1, L, dtype=torch.float32, device=self.device)
        with self._lock:
            ws = self._workspace(B, T, workspace)
            p, n = self._aligned(ws)
            if seed_dev is not None:
                per_row = 1 if (seed_dev.numel() == B and B > 1) else 0
                rc = self._lib.gnv_inference_dseed(self._h, _ptr(mel), _ptr(cache_source), cache_len, _ptr(lengths), B,
                                                   T, _ptr(seed_dev), per_row, _ptr(wav), _ptr(src), C.c_void_p(p), n,
                                                   _stream_ptr(self.device))
            else:
                rc = self._lib.gnv_inference(self._h, _ptr(mel), _ptr(cache_source), cache_len, _ptr(lengths), B, T,
                                             C.c_uint64(seed), _ptr(wav), _ptr(src), C.c_void_p(p), n,
                                             _stream_ptr(self.device))
            _cabi.check(rc, self._h, "gnv_inference")
        if T != T_true:                                   # bucketed: hand back exactly the caller's frames
            Lt = T_true * SAMPLES_PER_FRAME
            wav, src = wav[:, :Lt], src[:, :, :Lt]
            if B > 1:
                wav, src = wav.contiguous(), src.contiguous()
        return wav, src

    # -- extras -----------------------------------------------------------------------------------
    def debug_tap(self, name: str, B: int, T: int) -> torch.Tensor:
        """A named intermediate of the last decode of a (B, T) batch, as fp32 [B, C, L]."""
        ws = self._workspace(B, T)
        p, _ = self._aligned(ws)
        cap = B * (120 * T + 1) * 256
        out = torch.empty(cap, dtype=torch.float32, device=self.device)
        shape = (C.c_int64 * 3)()
        rc = self._lib.gnv_debug_tap(self._h, name.encode(), B, T, C.c_void_p(p), _ptr(out), cap, shape,
                                     _stream_ptr(self.device))
        _cabi.check(rc, self._h, "gnv_debug_tap")
        b, c, l = int(shape[0]), int(shape[1]), int(shape[2])
        return out[: b * c * l].view(b, c, l).clone()

    @torch.no_grad()
    def profile_inference(self, speech_feat: torch.Tensor, lengths=None, seed: int = 1):
        """One inference() with per-launch device times (gnv_inference_profile).  Returns
        (wav, [(name, kind, ms, algorithmic_flops), ...]); kind is _cabi.LAUNCH_*."""
        mel = self._check_in(speech_feat, "speech_feat")
        B, _, T = mel.shape
        lengths = self._lengths(lengths, B)
        L = T * SAMPLES_PER_FRAME
        wav = torch.empty(B, L, dtype=torch.float32, device=self.device)
        src = torch.empty(B, 1, L, dtype=torch.float32, device=self.device)
        cap = 256
        ms = (C.c_float * cap)()
        kinds = (C.c_int32 * cap)()
        flops = (C.c_double * cap)()
        names = C.create_string_buffer(cap * _cabi.LAUNCH_NAME_LEN)
        n = C.c_int()
        with self._lock:
            ws = self._workspace(B, T)
            p, nbytes = self._aligned(ws)
            rc = self._lib.gnv_inference_profile(self._h, _ptr(mel), _ptr(lengths), B, T, C.c_uint64(seed), _ptr(wav),
                                                 _ptr(src), C.c_void_p(p), nbytes, _stream_ptr(self.device), cap, ms,
                                                 kinds, flops, names, C.byref(n))
            _cabi.check(rc, self._h, "gnv_inference_profile")
        rows = []
        for i in range(n.value):
            raw = names.raw[i * _cabi.LAUNCH_NAME_LEN:(i + 1) * _cabi.LAUNCH_NAME_LEN]
            rows.append((raw.split(b"\0", 1)[0].decode(), int(kinds[i]), float(ms[i]), float(flops[i])))
        return wav, rows

    def launches(self, B: int, T: int, inference: bool = True) -> int:
        n = C.c_int()
        fn = self._lib.gnv_inference_launches if inference else self._lib.gnv_decode_launches
        _cabi.check(fn(self._h, B, T, C.byref(n)), self._h, "launch count")
        return n.value

    def plan_stats(self) -> dict:
        """Launch-plan cache of the handle (gnv_plan_stats): plans cached / built so far, device slots, pinned plans."""
        out = (C.c_uint64 * 4)()
        _cabi.check(self._lib.gnv_plan_stats(self._h, out), self._h, "gnv_plan_stats")
        return {"cached": int(out[0]), "built": int(out[1]), "slots": int(out[2]), "pinned": int(out[3])}


# -------------------------------------------------------------------------------------------------
# the streaming tail as a free function (no decoder handle needed)
# -------------------------------------------------------------------------------------------------
def fade_window(n: int = 480, device="cpu") -> torch.Tensor:
    """(cos(linspace(pi, 0, n)) + 1) / 2 — the raised cosine of upstream S3Token2Wav.trim_fade."""
    return ((torch.cos(torch.linspace(torch.pi, 0, n, dtype=torch.float32)) + 1) / 2).to(device)


def trim_fade_window(device="cpu") -> torch.Tensor:
    w = torch.zeros(960, dtype=torch.float32)
    w[480:] = fade_window(480)
    return w.to(device)


@torch.no_grad()
def pcm_tail(cur: torch.Tensor, prev_tail: Optional[torch.Tensor] = None, fade_w: Optional[torch.Tensor] = None,
             limit: float = 0.99, want_i16: bool = True, want_f32: bool = False,
             out_i16: Optional[torch.Tensor] = None, out_f32: Optional[torch.Tensor] = None):
    """cur [rows, n] fp32 (row stride free) -> (int16 [rows, n] | None, fp32 [rows, n] | None).
    Head fade / crossfade with prev_tail [rows, fade], clamp(+-limit), round-half-even int16 pack."""
    if cur.dim() != 2 or cur.dtype != torch.float32 or cur.device.type != "cuda":
        raise ValueError("cur must be a 2-D fp32 CUDA tensor")
    if cur.stride(1) != 1:
        cur = cur.contiguous()
    rows, n = cur.shape
    fade = 0
    if fade_w is not None:
        fade = min(int(fade_w.numel()), n)
        fade_w = fade_w.to(cur.device, torch.float32).reshape(-1)[:fade].contiguous()
    if prev_tail is not None:
        if fade_w is None:
            raise ValueError("prev_tail needs fade_w")
        prev_tail = prev_tail.to(torch.float32).reshape(rows, -1)[:, :fade].contiguous()
    if want_i16 and out_i16 is None:
        out_i16 = torch.empty(rows, n, dtype=torch.int16, device=cur.device)
    if want_f32 and out_f32 is None:
        out_f32 = torch.empty(rows, n, dtype=torch.float32, device=cur.device)
    if rows == 0 or n == 0:
        return out_i16, out_f32
    lib = _cabi.load()
    with torch.cuda.device(cur.device):
        rc = lib.gnv_pcm_tail(_ptr(cur), cur.stride(0) if rows > 1 else n, _ptr(prev_tail), _ptr(fade_w), rows, n,
                              fade, C.c_float(limit), _ptr(out_i16), _ptr(out_f32), n, _stream_ptr(cur.device))
    _cabi.check(rc, None, "gnv_pcm_tail")
    return out_i16, out_f32


@torch.no_grad()
def mulaw_encode(pcm_i16: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """int16 PCM (CUDA, contiguous) -> G.711 mu-law bytes (uint8), bit-identical to audioop.lin2ulaw."""
    if pcm_i16.dtype != torch.int16 or pcm_i16.device.type != "cuda":
        raise ValueError("pcm_i16 must be an int16 CUDA tensor")
    pcm_i16 = pcm_i16.contiguous()
    if out is None:
        out = torch.empty(pcm_i16.shape, dtype=torch.uint8, device=pcm_i16.device)
    if pcm_i16.numel() == 0:
        return out
    if pcm_i16.data_ptr() % 16 or out.data_ptr() % 8:
        raise ValueError("mulaw_encode needs a 16-byte aligned input and an 8-byte aligned output")
    lib = _cabi.load()
    with torch.cuda.device(pcm_i16.device):
        rc = lib.gnv_pcm_mulaw(_ptr(pcm_i16), pcm_i16.numel(), _ptr(out), _stream_ptr(pcm_i16.device))
    _cabi.check(rc, None, "gnv_pcm_mulaw")
    return out
