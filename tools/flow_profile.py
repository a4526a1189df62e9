"""Per-launch device times of ONE estimator evaluation of the flow decoder (gnv_flow_profile), grouped by layer class."""
import collections
import re
import sys
import torch

sys.path.insert(0, ".")
from gonova_tts_b200 import B200Flow  # noqa: E402
from gonova_tts_b200.flow import random_flow_state_dict  # noqa: E402

dev = torch.device("cuda:0")
dtype = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
T = int(sys.argv[3]) if len(sys.argv) > 3 else 500
flow = B200Flow(random_flow_state_dict(0), device=dev, dtype=dtype)
g = torch.Generator().manual_seed(1)
z, mu, cond = (torch.randn(B, 80, T, generator=g).to(dev) for _ in range(3))
spks = torch.randn(B, 80, generator=g).to(dev)
for _ in range(2):
    flow.profile(z, mu, spks, cond, n_timesteps=1)
acc = collections.OrderedDict()
runs = 3
for _ in range(runs):
    rows = flow.profile(z, mu, spks, cond, n_timesteps=1)
    for name, kind, ms, fl in rows:
        cls = re.sub(r"^(down_blocks\.0|mid_blocks\.\d+|up_blocks\.0)\.", "L.", name)
        cls = re.sub(r"^L\.1\.\d+\.", "L.1.j.", cls)
        a = acc.setdefault(cls, [0, 0.0, 0.0])
        a[0] += 1; a[1] += ms; a[2] += fl
tot = sum(a[1] for a in acc.values()) / runs
print(f"# one estimator evaluation, {dtype}, B={B} (x2 CFG rows), T={T}: {tot:.3f} ms over {len(rows)} launches")
print("class,launches,ms_total,ms_each,tflops,share")
for cls, (n, ms, fl) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    n //= runs; ms /= runs; fl /= runs
    print(f"{cls},{n},{ms:.4f},{ms / n:.4f},{fl / ms / 1e9 if ms else 0:.0f},{ms / tot:.3f}")
