"""Timeline of CTA 0 of flow_blk_kernel: python tools/flow_blk_trace.py <mode 0 ff | 1 out | 2 qkv | 3 attention> [B] [T]"""
import ctypes as C
import os
import sys

mode = sys.argv[1] if len(sys.argv) > 1 else "0"
os.environ.setdefault("GONOVA_FB_DBG", "8")
os.environ["GONOVA_FB_TRACE_MODE"] = mode
import torch

sys.path.insert(0, ".")
from gonova_tts_b200 import B200Flow, _cabi  # noqa: E402
from gonova_tts_b200.flow import random_flow_state_dict  # noqa: E402

B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
T = int(sys.argv[3]) if len(sys.argv) > 3 else 500
dev = torch.device("cuda:0")
flow = B200Flow(random_flow_state_dict(0), device=dev, dtype="bf16")
g = torch.Generator().manual_seed(1)
z, mu, cond = (torch.randn(B, 80, T, generator=g).to(dev) for _ in range(3))
spks = torch.randn(B, 80, generator=g).to(dev)
lib = _cabi.load()
buf = (C.c_uint64 * 8192)()
n = C.c_int()
flow.decode(z, mu, spks, cond, n_timesteps=1)
lib.gnv_debug_flow_trace(buf, 8192, C.byref(n))      # warm-up run discarded
flow.decode(z, mu, spks, cond, n_timesteps=1)
lib.gnv_debug_flow_trace(buf, 8192, C.byref(n))
ev = []
for i in range(n.value):
    w = buf[i]
    ev.append((w & 0xFFFFFFFF, ((w >> 56) & 0xFF) - 1, (w >> 48) & 0xFF, (w >> 40) & 0xFF, (w >> 32) & 0xFF))
ev.sort()
t0 = ev[0][0] if ev else 0
names = {0: "COL", 1: "MMA", 2: "PRD", 3: "ROW"} if mode != "3" else {0: "SM0", 1: "SM1", 2: "MMA"}
for c, role, a, b, e in ev[: int(os.environ.get("TRACE_LINES", "400"))]:
    print(f"{c - t0:9d} {names.get(role, role)} item{a} sub{b} ev{e}")
print("events", n.value)
