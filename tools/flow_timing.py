"""Device time of one gnv_flow_decode (ten Euler steps, classifier-free guidance) for a few batch shapes."""
import sys
import torch

sys.path.insert(0, ".")
from gonova_tts_b200 import B200Flow  # noqa: E402
from gonova_tts_b200.flow import random_flow_state_dict  # noqa: E402

dev = torch.device("cuda:0")
dtype = sys.argv[1] if len(sys.argv) > 1 else "bf16"
flow = B200Flow(random_flow_state_dict(0), device=dev, dtype=dtype)
for B, T in ((1, 100), (1, 500), (8, 500), (32, 500)):
    g = torch.Generator().manual_seed(1)
    z, mu, cond = (torch.randn(B, 80, T, generator=g).to(dev) for _ in range(3))
    spks = torch.randn(B, 80, generator=g).to(dev)
    for _ in range(2):
        flow.decode(z, mu, spks, cond)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 3
    e0.record()
    for _ in range(n):
        flow.decode(z, mu, spks, cond)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    gflop = 2 * 65e6 * 2 * B * T * 10 / 1e9          # ~65 MMAC per frame per estimator pass (GEMMs), 2 B rows, 10 steps
    print(f"{dtype} B={B} T={T}: {ms:8.2f} ms per decode = {B * T / 50 / (ms / 1e3):9.0f} audio-s/s, launches {flow.launches(10)}, "
          f"~{gflop / ms:.0f} GFLOP/ms GEMM-only")
