"""Host cost of the first call at a new (B, L) shape for each stage of tokens -> PCM (plan build: tensor maps, position tables)."""
import sys
import time
import torch

sys.path.insert(0, ".")
from gonova_tts_b200 import B200Flow, B200FlowFront, B200HiFT, random_state_dict  # noqa: E402
from gonova_tts_b200.flow import random_flow_state_dict  # noqa: E402
from gonova_tts_b200.flow_front import random_front_state_dict  # noqa: E402

dev = torch.device("cuda:0")
front = B200FlowFront(random_front_state_dict(0), device=dev, dtype="bf16")
flow = B200Flow(random_flow_state_dict(0), device=dev, dtype="bf16")
hift = B200HiFT(random_state_dict(0, False), device=dev, dtype="bf16")
g = torch.Generator().manual_seed(0)


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3


for B, L in ((4, 200), (4, 201), (4, 233), (7, 233), (16, 240), (16, 250), (4, 200)):
    T = 2 * L
    tok = torch.randint(0, 6561, (B, L), generator=g, dtype=torch.int32).to(dev)
    emb = torch.randn(B, 192, generator=g).to(dev)
    z, mu, cond = (torch.randn(B, 80, T, generator=g).to(dev) for _ in range(3))
    spks = torch.randn(B, 80, generator=g).to(dev)
    a1, a2 = timed(lambda: front.encode(tok, None, emb)), timed(lambda: front.encode(tok, None, emb))
    b1, b2 = timed(lambda: flow.decode(z, mu, spks, cond, n_timesteps=1)), timed(lambda: flow.decode(z, mu, spks, cond, n_timesteps=1))
    c1, c2 = timed(lambda: hift.inference(mu)), timed(lambda: hift.inference(mu))
    print(f"B={B:3d} L={L:3d}: front first {a1:7.2f} ms, again {a2:6.2f} | flow (1 step) first {b1:7.2f}, again {b2:6.2f} | "
          f"vocoder first {c1:7.2f}, again {c2:6.2f}")
