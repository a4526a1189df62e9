"""One conv layer through gnv_conv1d (the same kernel gnv_decode launches), for ncu.
usage: python tools/profile_conv.py C L k dil B dtype act residual [transposed stride pad]"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from gonova_tts_b200 import _cabi, build  # noqa: E402

build.build()
lib = _cabi.load()
a = sys.argv[1:]
Cc, L, k, dil, B = (int(x) for x in a[:5])
dtype, act, residual = a[5], a[6], int(a[7])
transposed = int(a[8]) if len(a) > 8 else 0
stride = int(a[9]) if len(a) > 9 else 1
Cin = Cc * 2 if transposed else Cc
pad = int(a[10]) if len(a) > 10 else (k * dil - dil) // 2
import os
FLAGS = int(os.environ.get("HOOK_FLAGS", "0"))
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
x = torch.randn(B, Cin, L, generator=g).to(dev)
w = torch.randn(*((Cin, Cc, k) if transposed else (Cc, Cin, k)), generator=g) * 0.05
b = torch.zeros(Cc)
alpha = torch.ones(Cc)
Lout = L * stride if transposed else L
res = torch.randn(B, Cc, Lout, generator=g).to(dev) if residual else None
out = torch.empty(B, Cc, Lout, device=dev)
for _ in range(3):
    rc = lib.gnv_conv1d(0, _cabi.DTYPE[dtype], FLAGS, transposed, C.c_void_p(x.data_ptr()), B, Cin, L,
                        C.c_void_p(w.data_ptr()), C.c_void_p(b.data_ptr()), Cc, k, stride, pad, dil, _cabi.ACT[act],
                        C.c_void_p(alpha.data_ptr()), C.c_float(0.1), None if res is None else C.c_void_p(res.data_ptr()),
                        C.c_void_p(out.data_ptr()), Lout, None)
    assert rc == 0, _cabi.last_error(None)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
