import sys, torch
sys.path.insert(0, ".")
from gonova_tts_b200 import B200Flow
from oracle import flow_ref as FR
dev = torch.device("cuda:0")
flow = B200Flow(FR.random_state_dict(0), device=dev, dtype="bf16")
B, T = 8, 500
z, mu, mask, spks, cond = [t.to(dev) for t in FR.synthetic_inputs(B, T, seed=1)]
flow.decode(z, mu, spks, cond, n_timesteps=1)
torch.cuda.synchronize()
flow.decode(z, mu, spks, cond, n_timesteps=1)
torch.cuda.synchronize()
