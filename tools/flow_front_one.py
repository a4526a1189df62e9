"""One gnv_flow_encode of a given shape (the command ncu wraps): python tools/flow_front_one.py bf16 32 250"""
import sys
import torch

sys.path.insert(0, ".")
from gonova_tts_b200.flow_front import B200FlowFront, random_front_state_dict  # noqa: E402

dev = torch.device("cuda:0")
dtype = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B, L = (int(sys.argv[i]) if len(sys.argv) > i else d for i, d in ((2, 32), (3, 250)))
front = B200FlowFront(random_front_state_dict(0), device=dev, dtype=dtype)
g = torch.Generator().manual_seed(1)
tokens = torch.randint(0, 6561, (B, L), generator=g, dtype=torch.int32).to(dev)
emb = torch.randn(B, 192, generator=g).to(dev)
mu, spks = front.encode(tokens, None, emb)
torch.cuda.synchronize()
print("ok", float(mu.abs().max()))
