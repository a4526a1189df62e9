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
