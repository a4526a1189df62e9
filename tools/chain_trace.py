"""Timeline of CTA 0 of conv_chain_kernel (GONOVA_CHAIN_DBG=8): prints per-role events with clock deltas."""
import ctypes as C
import os
import sys

os.environ["GONOVA_CHAIN_DBG"] = str(8 | int(os.environ.get("DBG_EXTRA", "0")))
import torch

sys.path.insert(0, ".")
from gonova_tts_b200 import B200HiFT, _cabi, random_state_dict  # noqa: E402

B, T = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda:0")
dec = B200HiFT(random_state_dict(0, False), device=dev, dtype="bf16")
g = torch.Generator().manual_seed(0)
mel = (-5 + 2 * torch.randn(B, 80, T, generator=g)).clamp(-11.5, 2.5).to(dev)
s = torch.zeros(B, 1, T * 480, device=dev)
dec.decode(mel, s)
lib = _cabi.load()
buf = (C.c_uint64 * 8192)()
n = C.c_int()
lib.gnv_debug_chain_trace(buf, 8192, C.byref(n))      # warm-up run discarded
dec.decode(mel, s)
lib.gnv_debug_chain_trace(buf, 8192, C.byref(n))
ev = []
for i in range(n.value):
    w = buf[i]
    ev.append((w & 0xFFFFFFFF, ((w >> 56) & 0xFF) - 1, (w >> 48) & 0xFF, (w >> 40) & 0xFF, (w >> 32) & 0xFF))
ev.sort()
t0 = ev[0][0] if ev else 0
names = {0: "EPI", 1: "MMA", 2: "PRD"}
limit = int(os.environ.get("TRACE_LINES", "400"))
for c, role, lane, ph, e in ev[:limit]:
    print(f"{c - t0:9d} {names.get(role, role)} lane{lane} phase{ph} ev{e}")
print("events", n.value)
