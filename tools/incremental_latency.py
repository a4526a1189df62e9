"""Latency of intra-sentence streaming: mel frames arrive 2 at a time (one 25 Hz speech token = 2 mel frames); for every
chunk, the time from the push that completes its look-ahead to its int16 PCM sitting in pinned host memory."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gonova_tts_b200 import B200HiFT, IncrementalDecoder, random_state_dict
from bench import synthetic_mel

dev = torch.device("cuda:0")
hift = B200HiFT(random_state_dict(0, False), device=dev, dtype="bf16")
hift.reserve(1, 160)
T = 1000
mel = synthetic_mel(1, T, 3).to(dev)
host = torch.empty(1, 100 * 480, dtype=torch.int16).pin_memory()
lat, idle = [], []
for rep in range(6):
    inc = IncrementalDecoder(hift, B=1, seed=rep + 1)
    for t in range(0, T, 2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = inc.push(mel[:, :, t:t + 2])
        for i16, _ in out:
            host[:, : i16.shape[1]].copy_(i16, non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        if rep:                                            # first pass warms plans and kernels
            (lat if out else idle).append(dt)
    inc.finish()
lat.sort(); idle.sort()
print(f"pushes that release a 2 s chunk: n={len(lat)}  p50 {lat[len(lat) // 2]:.2f} ms  p99 {lat[int(len(lat) * 0.99)]:.2f} ms")
print(f"pushes that only extend f0 / source: n={len(idle)}  p50 {idle[len(idle) // 2]:.3f} ms  p99 {idle[int(len(idle) * 0.99)]:.3f} ms")
