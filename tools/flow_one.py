"""One gnv_flow_decode of a given shape (the command ncu wraps): python tools/flow_one.py bf16 32 500 1"""
import sys
import torch

sys.path.insert(0, ".")
from gonova_tts_b200 import B200Flow  # noqa: E402
from gonova_tts_b200.flow import random_flow_state_dict  # noqa: E402

dev = torch.device("cuda:0")
dtype = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B, T, steps = (int(sys.argv[i]) if len(sys.argv) > i else d for i, d in ((2, 32), (3, 500), (4, 1)))
flow = B200Flow(random_flow_state_dict(0), device=dev, dtype=dtype)
g = torch.Generator().manual_seed(1)
z, mu, cond = (torch.randn(B, 80, T, generator=g).to(dev) for _ in range(3))
spks = torch.randn(B, 80, generator=g).to(dev)
mel = flow.decode(z, mu, spks, cond, n_timesteps=steps)
torch.cuda.synchronize()
print("ok", float(mel.abs().max()))
