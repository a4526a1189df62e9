"""Host cost of the first call at a new (B, T): plan build (tensor maps, tile lists) vs a cached call."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gonova_tts_b200 import B200HiFT, random_state_dict
from bench import synthetic_mel

dev = torch.device("cuda:0")
dec = B200HiFT(random_state_dict(0, False), device=dev, dtype="bf16")
dec.inference(synthetic_mel(1, 50, 1).to(dev)); torch.cuda.synchronize()
first, again = [], []
for T in (117, 233, 351, 487, 500, 612, 745):
    mel = synthetic_mel(1, T, T).to(dev)
    torch.cuda.synchronize(); t0 = time.perf_counter(); dec.inference(mel); torch.cuda.synchronize(); t1 = time.perf_counter()
    dec.inference(mel); torch.cuda.synchronize(); t2 = time.perf_counter()
    first.append((t1 - t0) * 1e3); again.append((t2 - t1) * 1e3)
    print(f"T={T}: first call {first[-1]:.2f} ms, cached {again[-1]:.2f} ms")
