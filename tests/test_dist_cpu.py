"""The N>1 harness on CPU: two gloo ranks shard the streams, run the same step loop bench.py runs
(barrier, timed region, max over ranks) and agree on the aggregate."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_streams, out_q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gonova_tts_b200.dispatch import shard_range

        lo, hi = shard_range(n_streams, world, rank)
        mine = torch.zeros(n_streams, dtype=torch.int64)
        mine[lo:hi] = 1
        dist.barrier()
        fake_ms = torch.tensor([10.0 * (rank + 1)], dtype=torch.float64)   # this rank's device time
        dist.all_reduce(fake_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(mine, op=dist.ReduceOp.SUM)
        units = torch.tensor([hi - lo], dtype=torch.int64)
        dist.all_reduce(units, op=dist.ReduceOp.SUM)
        out_q.put((rank, float(fake_ms), mine.tolist(), int(units)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_sharding_and_max_timing():
    world, n = 2, 13
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ms, owned, units in res:
        assert ms == 20.0                       # max over ranks, not this rank's own time
        assert owned == [1] * n                 # every stream owned by exactly one rank
        assert units == n


@pytest.mark.timeout(300)
def test_reference_arm_prints_one_line_alone_and_under_torchrun():
    """`bench.py --impl reference` (the oracle port timed on the host cores) needs no GPU: one JSON line with the
    contract's keys; under torchrun (N = 2) rank 0 alone prints it and the other rank exits 0 without work."""
    import json
    import subprocess

    base = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
            "--frames", "40"]
    r = subprocess.run(base, capture_output=True, text=True, timeout=200, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0

    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "0", "--frames", "40"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=280, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1 and json.loads(lines[0])["n_gpus"] == 2
