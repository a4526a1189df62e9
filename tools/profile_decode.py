"""One decode of a B x T batch (bf16), for ncu.  usage: python tools/profile_decode.py B T"""
import sys

import torch

sys.path.insert(0, ".")
from gonova_tts_b200 import B200HiFT, random_state_dict  # noqa: E402

B, T = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda:0")
dec = B200HiFT(random_state_dict(0, False), device=dev, dtype="bf16")
g = torch.Generator().manual_seed(0)
mel = (-5 + 2 * torch.randn(B, 80, T, generator=g)).clamp(-11.5, 2.5).to(dev)
wav, src = dec.inference(mel, seed=1)
torch.cuda.synchronize()
print("ok", float(wav.abs().mean()))
