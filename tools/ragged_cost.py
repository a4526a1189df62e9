"""Cost of a ragged batch (lengths uniform in [Tmin, T]) against the full padded batch: dead tiles are skipped."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gonova_tts_b200 import B200HiFT, random_state_dict
from bench import synthetic_mel

dev = torch.device("cuda:0")
B, T = 64, 500
dec = B200HiFT(random_state_dict(0, False), device=dev, dtype="bf16")
mel = synthetic_mel(B, T, 1).to(dev)
wav = torch.empty(B, T * 480, device=dev)
src = torch.empty(B, 1, T * 480, device=dev)


def time_ms(lengths, n=5):
    for _ in range(3):
        dec.inference(mel, seed=1, lengths=lengths, out=wav, source_out=src)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        dec.inference(mel, seed=1, lengths=lengths, out=wav, source_out=src)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


full = time_ms(None)
print(f"full batch 64 x 500 frames: {full:.2f} ms")
g = torch.Generator().manual_seed(0)
for tmin in (400, 250, 100, 25):
    lens = torch.randint(tmin, T + 1, (B,), generator=g).tolist()
    ms = time_ms(lens)
    frac = sum(lens) / (B * T)
    print(f"lengths U[{tmin},{T}]: {ms:.2f} ms  = {ms / full:.2f} of full; live frames {frac:.2f} of padded")
