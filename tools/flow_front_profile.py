"""Per-launch device times of ONE encode of the flow front (gnv_flow_encode_profile), grouped by layer class and stage."""
import collections
import sys
import torch

sys.path.insert(0, ".")
from gonova_tts_b200.flow_front import B200FlowFront, random_front_state_dict  # noqa: E402

dev = torch.device("cuda:0")
dtype = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
L = int(sys.argv[3]) if len(sys.argv) > 3 else 250
front = B200FlowFront(random_front_state_dict(0), device=dev, dtype=dtype)
g = torch.Generator().manual_seed(1)
tokens = torch.randint(0, 6561, (B, L), generator=g, dtype=torch.int32).to(dev)
emb = torch.randn(B, 192, generator=g).to(dev)
for _ in range(2):
    front.profile(tokens, None, emb)
acc = collections.OrderedDict()
runs = 3
for _ in range(runs):
    rows = front.profile(tokens, None, emb)
    stage = 1
    for name, kind, ms, fl in rows:
        if name == "up_layer.conv":
            stage = 2
        cls = f"{name}@{stage}" if name.startswith(("self_attn", "feed_forward", "norm_")) else name
        a = acc.setdefault(cls, [0, 0.0, 0.0])
        a[0] += 1; a[1] += ms; a[2] += fl
tot = sum(a[1] for a in acc.values()) / runs
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
front.encode(tokens, None, emb)
torch.cuda.synchronize()
ev0.record()
for _ in range(10):
    front.encode(tokens, None, emb)
ev1.record()
torch.cuda.synchronize()
print(f"# one encode, {dtype}, B={B}, L={L} tokens (-> {2 * L} frames): {tot:.3f} ms summed over {len(rows)} launches; "
      f"{ev0.elapsed_time(ev1) / 10:.3f} ms per call back to back")
print("class,launches,ms_total,ms_each,tflops,share")
for cls, (n, ms, fl) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    n //= runs; ms /= runs; fl /= runs
    print(f"{cls},{n},{ms:.4f},{ms / n:.4f},{fl / ms / 1e9 if ms else 0:.0f},{ms / tot:.3f}")
