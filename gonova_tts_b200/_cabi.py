"""ctypes binding of libgonova_hift.so (include/gonova_hift.h).  Loading fails loudly: there is no
CPU or PyTorch fallback behind these entry points."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_LIB = None

DTYPE = {"tf32": 0, "bf16": 1, "fp32": 2}
FLAG_SIMT_CONV = 1
FLAG_PRECISE_ACT = 2
FLAG_TC_V1 = 4
LAUNCH_AUX, LAUNCH_CONV_TC, LAUNCH_CONV_SIMT, LAUNCH_NAME_LEN = 0, 1, 2, 48
ACT = {"none": 0, "snake": 1, "lrelu": 2, "elu": 3, "snake_fast": 4}

# every symbol include/gonova_hift.h declares
SYMBOLS = [
    "gnv_create", "gnv_destroy", "gnv_last_error", "gnv_abi_version", "gnv_workspace_bytes", "gnv_f0",
    "gnv_source", "gnv_decode", "gnv_inference", "gnv_inference_profile", "gnv_pcm_tail", "gnv_pcm_mulaw", "gnv_stft", "gnv_istft", "gnv_conv1d",
    "gnv_debug_tap", "gnv_debug_cluster_probe", "gnv_decode_launches", "gnv_inference_launches",
    "gnv_plan_stats", "gnv_source_stream", "gnv_inference_dseed",
    "gnv_debug_chain_trace",
    "gnv_flow_create", "gnv_flow_destroy", "gnv_flow_workspace_bytes", "gnv_flow_decode", "gnv_flow_launches",
    "gnv_flow_profile", "gnv_debug_flow_trace",
    "gnv_flow_enc_create", "gnv_flow_enc_destroy", "gnv_flow_enc_workspace_bytes", "gnv_flow_encode",
    "gnv_flow_encode_profile", "gnv_flow_enc_launches",
]


class GnvWeight(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.POINTER(C.c_float)), ("ndim", C.c_int32),
                ("shape", C.c_int64 * 4)]


def lib_path() -> Path:
    env = os.environ.get("GONOVA_HIFT_LIB")
    return Path(env) if env else Path(__file__).resolve().parent / "lib" / "libgonova_hift.so"


def load():
    """Returns the loaded library with argtypes set; raises if it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not path.exists():
        raise ImportError(
            f"{path} is missing: build it with `python -m gonova_tts_b200.build` (needs nvcc, sm_100a). "
            "gonova_tts_b200 has no CPU fallback.")
    lib = C.CDLL(str(path))
    vp, f32p, i32p, i16p = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p   # raw device/host addresses
    lib.gnv_abi_version.restype = C.c_int
    lib.gnv_last_error.restype = C.c_char_p
    lib.gnv_last_error.argtypes = [vp]
    lib.gnv_create.argtypes = [C.POINTER(GnvWeight), C.c_int, C.c_int, C.c_int, C.c_uint, C.POINTER(vp)]
    lib.gnv_destroy.argtypes = [vp]
    lib.gnv_destroy.restype = None
    lib.gnv_workspace_bytes.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_size_t)]
    lib.gnv_f0.argtypes = [vp, f32p, i32p, C.c_int, C.c_int, f32p, vp, C.c_size_t, vp]
    lib.gnv_source.argtypes = [vp, f32p, C.c_int, C.c_int, C.c_uint64, f32p, f32p, f32p, vp]
    lib.gnv_decode.argtypes = [vp, f32p, f32p, i32p, C.c_int, C.c_int, f32p, vp, C.c_size_t, vp]
    lib.gnv_inference.argtypes = [vp, f32p, f32p, C.c_int, i32p, C.c_int, C.c_int, C.c_uint64, f32p, f32p, vp,
                                  C.c_size_t, vp]
    lib.gnv_inference_dseed.argtypes = [vp, f32p, f32p, C.c_int, i32p, C.c_int, C.c_int, vp, C.c_int, f32p, f32p, vp,
                                        C.c_size_t, vp]
    lib.gnv_inference_profile.argtypes = [vp, f32p, i32p, C.c_int, C.c_int, C.c_uint64, f32p, f32p, vp, C.c_size_t, vp,
                                          C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_double),
                                          C.c_char_p, C.POINTER(C.c_int)]
    lib.gnv_pcm_tail.argtypes = [f32p, C.c_int64, f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_float, i16p, f32p,
                                 C.c_int64, vp]
    lib.gnv_pcm_mulaw.argtypes = [i16p, C.c_int64, vp, vp]
    lib.gnv_stft.argtypes = [f32p, C.c_int, C.c_int, f32p, vp]
    lib.gnv_istft.argtypes = [f32p, C.c_int, C.c_int, C.c_float, f32p, vp]
    lib.gnv_conv1d.argtypes = [C.c_int, C.c_int, C.c_uint, C.c_int, f32p, C.c_int, C.c_int, C.c_int, f32p, f32p,
                               C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, f32p, C.c_float, f32p, f32p,
                               C.c_int, vp]
    lib.gnv_debug_tap.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, vp, f32p, C.c_size_t, C.POINTER(C.c_int64), vp]
    lib.gnv_flow_create.argtypes = [C.POINTER(GnvWeight), C.c_int, C.c_int, C.c_int, C.c_uint, C.POINTER(vp)]
    lib.gnv_flow_destroy.argtypes = [vp]
    lib.gnv_flow_destroy.restype = None
    lib.gnv_flow_workspace_bytes.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_size_t)]
    lib.gnv_flow_decode.argtypes = [vp, f32p, f32p, f32p, f32p, i32p, C.c_int, C.c_int, C.c_int, C.c_float, f32p, vp,
                                    C.c_size_t, vp]
    lib.gnv_flow_profile.argtypes = [vp, f32p, f32p, f32p, f32p, i32p, C.c_int, C.c_int, C.c_int, C.c_float, f32p, vp,
                                     C.c_size_t, vp, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int32),
                                     C.POINTER(C.c_double), C.c_char_p, C.POINTER(C.c_int)]
    lib.gnv_flow_launches.argtypes = [vp, C.c_int, C.POINTER(C.c_int)]
    lib.gnv_flow_enc_create.argtypes = [C.POINTER(GnvWeight), C.c_int, C.c_int, C.c_int, C.c_uint, C.POINTER(vp)]
    lib.gnv_flow_enc_destroy.argtypes = [vp]
    lib.gnv_flow_enc_destroy.restype = None
    lib.gnv_flow_enc_workspace_bytes.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_size_t)]
    lib.gnv_flow_encode.argtypes = [vp, i32p, i32p, f32p, C.c_int, C.c_int, f32p, f32p, vp, C.c_size_t, vp]
    lib.gnv_flow_encode_profile.argtypes = [vp, i32p, i32p, f32p, C.c_int, C.c_int, f32p, f32p, vp, C.c_size_t, vp, C.c_int,
                                            C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_double), C.c_char_p,
                                            C.POINTER(C.c_int)]
    lib.gnv_flow_enc_launches.argtypes = [vp, C.POINTER(C.c_int)]
    lib.gnv_debug_chain_trace.argtypes = [C.POINTER(C.c_uint64), C.c_int, C.POINTER(C.c_int)]
    lib.gnv_debug_cluster_probe.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int)]
    lib.gnv_decode_launches.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_int)]
    lib.gnv_inference_launches.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_int)]
    lib.gnv_plan_stats.argtypes = [vp, C.POINTER(C.c_uint64)]
    lib.gnv_source_stream.argtypes = [vp, f32p, C.c_int, C.c_int, C.c_uint64, C.c_int64, vp, f32p, vp, vp]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if name not in ("gnv_destroy", "gnv_last_error", "gnv_abi_version", "gnv_flow_destroy", "gnv_flow_enc_destroy"):
            fn.restype = C.c_int
    _LIB = lib
    return lib


def last_error(handle=None) -> str:
    msg = load().gnv_last_error(handle)
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, handle=None, what: str = "libgonova_hift"):
    if rc != 0:
        raise RuntimeError(f"{what} failed: {last_error(handle) or last_error(None)}")
