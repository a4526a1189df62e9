"""Closed-loop service simulation for the micro-batching front end.

C client threads stand for C open WebSocket connections: each submits one sentence (a mel of 2-10 s, its own length),
waits for its waveform on the host, submits the next.  The reference serves these one at a time
(services/tts/server.py:110-186); here `batching.for_decoder` gathers what arrives within max_wait_ms into one ragged
batch per decoder call.  Prints aggregate audio-seconds per second and the request latency distribution.
usage: python tools/service_sim.py [clients] [requests] [max_wait_ms]"""
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gonova_tts_b200 import B200HiFT, random_state_dict
from gonova_tts_b200.batching import for_decoder
from bench import synthetic_mel

clients = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n_req = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
wait_ms = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0

dev = torch.device("cuda:0")
dec = B200HiFT(random_state_dict(0, False), device=dev, dtype="bf16")
mb = for_decoder(dec, max_batch=64, max_wait_ms=wait_ms, max_queue=max(4 * clients, 256), pad_frames=16, max_frames=512)
g = torch.Generator().manual_seed(0)
lengths = torch.randint(100, 501, (n_req,), generator=g).tolist()
pool = synthetic_mel(1, 512, 7)[0]                       # every request is a window of this mel

# warm the shapes the run will see (plan builds, kernel images)
for f in [mb.submit(pool[:, : lengths[i]].clone()) for i in range(min(64, n_req))]:
    f.result(timeout=60)
for k in list(mb.metrics):
    mb.metrics[k] = 0

lat, lock, nxt = [], threading.Lock(), [0]


def client():
    while True:
        with lock:
            i = nxt[0]
            nxt[0] += 1
        if i >= n_req:
            return
        mel = pool[:, : lengths[i]].clone()
        t0 = time.perf_counter()
        wav = mb.submit(mel).result(timeout=120)
        dt = time.perf_counter() - t0
        assert wav.shape[0] == lengths[i] * 480
        with lock:
            lat.append(dt * 1e3)


t0 = time.perf_counter()
th = [threading.Thread(target=client) for _ in range(clients)]
for t in th:
    t.start()
for t in th:
    t.join()
wall = time.perf_counter() - t0
mb.close()
lat.sort()
audio = sum(lengths) / 50.0
m = mb.metrics
print(f"clients {clients}, requests {n_req}, sentence length U[2,10] s, max_wait {wait_ms} ms")
print(f"  {audio / wall:9.0f} audio-s/s   ({audio:.0f} audio-s in {wall:.2f} s; {n_req / wall:.0f} sentences/s)")
print(f"  batches {m['batches']}, mean batch {m['requests'] / max(1, m['batches']):.1f}, padding {m['padded_frames'] / max(1, m['frames']) - 1:.0%}")
print(f"  request latency ms: p50 {lat[len(lat) // 2]:.1f}  p90 {lat[int(len(lat) * 0.9)]:.1f}  p99 {lat[int(len(lat) * 0.99)]:.1f}")
print(f"  plan cache: {dec.plan_stats()}")
