"""Closed-loop service simulation one level up (SURVEY §8f-4 over the whole tokens -> PCM path).

C client threads stand for C open connections: each submits one sentence (50-250 speech tokens = 2-10 s, its own length, one of
four voices), waits for its waveform, submits the next.  The reference runs one `generate` at a time
(services/tts/server.py:110-186); here `batching.for_token2wav` lets whatever queued up while a batch ran ride together through
the flow front, the ten Euler steps and the vocoder.  Prints aggregate audio-seconds per second and the request latency.
usage: python tools/service_sim_t2w.py [clients] [requests] [max_batch]"""
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gonova_tts_b200 import B200Token2Wav, random_state_dict
from gonova_tts_b200.batching import for_token2wav
from gonova_tts_b200.flow import random_flow_state_dict
from gonova_tts_b200.flow_front import random_front_state_dict

clients = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n_req = int(sys.argv[2]) if len(sys.argv) > 2 else 256
max_batch = int(sys.argv[3]) if len(sys.argv) > 3 else 32

dev = torch.device("cuda:0")
sd = {"flow." + k: v for k, v in random_front_state_dict(0).items()}
sd.update({"flow.decoder.estimator." + k: v for k, v in random_flow_state_dict(0).items()})
sd.update({"mel2wav." + k: v for k, v in random_state_dict(0, False).items()})
t2w = B200Token2Wav.from_state_dict(sd, device=dev, dtype="bf16")
g = torch.Generator().manual_seed(0)
voices = [{"prompt_token": torch.randint(0, 6561, (1, 25), generator=g, dtype=torch.int32).to(dev), "prompt_token_len": None,
           "prompt_feat": (torch.randn(1, 50, 80, generator=g) * 0.5).to(dev), "prompt_feat_len": None,
           "embedding": torch.randn(1, 192, generator=g).to(dev)} for _ in range(4)]
lengths = torch.randint(50, 251, (n_req,), generator=g).tolist()
tokens = [torch.randint(0, 6561, (n,), generator=g, dtype=torch.int32).to(dev) for n in lengths]
rb = for_token2wav(t2w, max_batch=max_batch, max_queue=max(4 * clients, 256))
for rep in range(3):                                      # warm-up: kernel images, launch plans of the common shapes
    for f in [rb.submit((tokens[(7 * rep + i) % n_req], voices[i % 4])) for i in range(min(2 * max_batch, n_req))]:
        f.result(timeout=300)
for k in list(rb.metrics):
    rb.metrics[k] = 0

lat, lock, nxt = [], threading.Lock(), [0]


def client():
    while True:
        with lock:
            i = nxt[0]
            nxt[0] += 1
        if i >= n_req:
            return
        t0 = time.perf_counter()
        wav = rb.submit((tokens[i], voices[i % 4])).result(timeout=300).cpu()
        dt = time.perf_counter() - t0
        assert wav.shape == (1, 2 * lengths[i] * 480)
        with lock:
            lat.append(dt * 1e3)


t0 = time.perf_counter()
th = [threading.Thread(target=client) for _ in range(clients)]
for t in th:
    t.start()
for t in th:
    t.join()
wall = time.perf_counter() - t0
rb.close()
lat.sort()
audio = sum(lengths) / 25.0
m = rb.metrics
print(f"clients {clients}, requests {n_req}, sentence length U[2,10] s (50-250 tokens + a 25-token prompt), max_batch {max_batch}, bf16")
print(f"  {audio / wall:9.0f} audio-s/s   ({audio:.0f} audio-s in {wall:.2f} s; {n_req / wall:.1f} sentences/s)")
print(f"  batches {m['batches']}, mean batch {m['requests'] / max(1, m['batches']):.1f}, largest {m['largest_batch']}")
print(f"  request latency ms: p50 {lat[len(lat) // 2]:.1f}  p90 {lat[int(len(lat) * 0.9)]:.1f}  p99 {lat[int(len(lat) * 0.99)]:.1f}")
