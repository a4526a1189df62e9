"""Throughput against batch size (T = 500 and T = 116), to see that the tile / CTA-pair / narrow-tile heuristics have no cliff."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gonova_tts_b200 import B200HiFT, random_state_dict
from bench import synthetic_mel

dev = torch.device("cuda:0")
dec = B200HiFT(random_state_dict(0, False), device=dev, dtype="bf16")
dec.reserve(64, 500)
for T in (500, 116):
    ref = None
    for B in (1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64):
        mel = synthetic_mel(B, T, 9).to(dev)
        wav = torch.empty(B, T * 480, device=dev)
        src = torch.empty(B, 1, T * 480, device=dev)
        for _ in range(3):
            dec.inference(mel, seed=1, out=wav, source_out=src)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20 if B <= 8 else 8
        e0.record()
        for _ in range(n):
            dec.inference(mel, seed=1, out=wav, source_out=src)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        if ref is None:
            ref = wav[0].clone()
        same = bool(torch.equal(wav[0], ref))              # row 0 is the same utterance at every B
        print(f"T={T:4d} B={B:3d}: {ms:7.3f} ms  {B * T / 50 / ms * 1e3:8.0f} audio-s/s  {ms / B:6.3f} ms per utterance  row0 identical: {same}")
