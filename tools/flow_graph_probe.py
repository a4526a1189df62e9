"""One utterance through gnv_flow_decode: eager launches against a CUDA-graph replay of the same decode."""
import sys
import torch

sys.path.insert(0, ".")
from gonova_tts_b200 import B200Flow  # noqa: E402
from gonova_tts_b200.flow import random_flow_state_dict  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
T = int(sys.argv[2]) if len(sys.argv) > 2 else 500
flow = B200Flow(random_flow_state_dict(0), device=dev, dtype="bf16")
g = torch.Generator().manual_seed(1)
z, mu, cond = (torch.randn(B, 80, T, generator=g).to(dev) for _ in range(3))
spks = torch.randn(B, 80, generator=g).to(dev)


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


eager = timed(lambda: flow.decode(z, mu, spks, cond))
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    flow.decode(z, mu, spks, cond)
torch.cuda.current_stream().wait_stream(s)
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    out = flow.decode(z, mu, spks, cond)
ref = flow.decode(z, mu, spks, cond)
graph.replay()
torch.cuda.synchronize()
print("graph == eager:", bool(torch.equal(out, ref)))
replay = timed(graph.replay)
print(f"B={B} T={T}: eager {eager:.2f} ms, graph replay {replay:.2f} ms per decode ({flow.launches(10)} launches)")
