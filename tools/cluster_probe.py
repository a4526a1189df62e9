import ctypes as C, sys
sys.path.insert(0, ".")
import torch
from gonova_tts_b200 import _cabi, build
build.build(); lib = _cabi.load(); torch.zeros(1, device="cuda")
for smem in (0, 49152, 100000, 150000, 200000, 220000, 232448):
    for grid in (2, 148):
        n = C.c_int(-2); rc = lib.gnv_debug_cluster_probe(smem, grid, C.byref(n))
        print(smem, grid, rc, n.value, _cabi.last_error(None) if rc else "")
