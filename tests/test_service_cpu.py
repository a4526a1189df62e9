"""Host logic of the service glue without a GPU: install() on a real nn.Module parent, and the asyncio worker that
replaces the reference's single-request `_tts_worker` (services/tts/server.py:110-186)."""
import asyncio
import time
from concurrent.futures import Future
from types import SimpleNamespace

import numpy as np
import pytest
import torch
from torch import nn

from fake_engine import FakeEngine
from gonova_tts_b200 import random_state_dict
from gonova_tts_b200.aio import AsyncBatchedDecoder, batched_tts_worker
from gonova_tts_b200.batching import MicroBatcher


def test_install_assigns_a_module_child_on_a_real_parent(monkeypatch):
    """`s3gen.mel2wav` is a registered nn.Module child: nn.Module.__setattr__ only takes another nn.Module there
    (round-1 bug: B200HiFT was a plain object and install() raised TypeError on the real engine)."""
    from gonova_tts_b200 import B200HiFT, service

    assert issubclass(B200HiFT, nn.Module)
    seen = {}

    def fake_from_module(cls, old, device=None, dtype="bf16", **kw):
        obj = cls.__new__(cls)                       # a B200HiFT without a CUDA handle (no GPU here)
        nn.Module.__init__(obj)
        obj._h = None
        seen.update(old=old, dtype=dtype, kw=kw, keys=len(old.state_dict()))
        return obj

    monkeypatch.setattr(B200HiFT, "from_module", classmethod(fake_from_module))
    eng = FakeEngine(random_state_dict(0, False), device="cpu")
    old = eng.s3gen.mel2wav
    new = service.install(eng, dtype="tf32", max_frames=0)
    assert eng.s3gen.mel2wav is new and isinstance(new, B200HiFT) and seen["old"] is old
    assert seen["dtype"] == "tf32" and seen["kw"] == {"bucket_frames": 8} and seen["keys"] > 300
    assert dict(eng.s3gen.named_children())["mel2wav"] is new
    assert new.eval() is new and new.to("cpu") is new
    assert eng.s3gen.state_dict() == {} or all(not k.startswith("mel2wav.") for k in eng.s3gen.state_dict())


def test_b200hift_refuses_to_run_without_cuda():
    from gonova_tts_b200 import B200HiFT

    if torch.cuda.is_available():
        pytest.skip("GPU box: covered by the gpu tests")
    with pytest.raises(RuntimeError, match="CUDA"):
        B200HiFT(random_state_dict(0, False), device="cuda:0")
    with pytest.raises(RuntimeError, match="CUDA"):
        B200HiFT(random_state_dict(0, False), device="cpu")


def test_flow_classes_refuse_to_run_without_cuda():
    """No CPU fallback anywhere on the tokens -> mel path either: the constructors raise before touching the weights."""
    from gonova_tts_b200 import B200Flow, B200FlowFront, B200FlowInference, B200Token2Wav

    if torch.cuda.is_available():
        pytest.skip("GPU box: covered by the gpu tests")
    for cls in (B200Flow, B200FlowFront, B200FlowInference):
        for dev in ("cuda:0", "cpu"):
            with pytest.raises(RuntimeError, match="CUDA"):
                cls({}, device=dev)
    with pytest.raises(RuntimeError, match="CUDA"):
        B200Token2Wav.from_state_dict({}, device="cuda:0")


class FakeQueueManager:
    """The three calls of services/tts/core/queue_manager.py the worker uses, same names and argument order."""

    def __init__(self, requests):
        self.input = asyncio.Queue()
        for r in requests:
            self.input.put_nowait(r)
        self.sent = []
        self.done = 0

    async def get_next_request(self, timeout: float = 0.05):
        try:
            return await asyncio.wait_for(self.input.get(), timeout=timeout)
        except asyncio.TimeoutError:
            return None

    async def enqueue_audio_chunk(self, connection_id, audio_data, chunk_id, is_final=False):
        self.sent.append((connection_id, audio_data, chunk_id, is_final))

    async def mark_request_done(self):
        self.done += 1


def test_batched_worker_keeps_the_reference_protocol_and_batches_requests():
    seen_batches = []

    def decode_fn(x, lengths):                       # the "GPU": sample value = its frame's first mel bin
        seen_batches.append(len(lengths))
        time.sleep(0.02)
        return x[:, 0, :].repeat_interleave(480, dim=1)

    mb = MicroBatcher(decode_fn, max_batch=8, max_wait_ms=20.0)
    dec = AsyncBatchedDecoder(mb)
    reqs = [SimpleNamespace(connection_id=f"c{i}", text="t" * (3 + i)) for i in range(12)]
    reqs.append(SimpleNamespace(connection_id="bad", text=""))
    errors = []

    def mel_fn(request):
        if not request.text:
            raise ValueError("empty text")
        return torch.full((80, len(request.text)), float(len(request.text)))

    async def main():
        qm = FakeQueueManager(reqs)
        stop = {"v": False}
        task = asyncio.ensure_future(batched_tts_worker(qm, mel_fn, dec, lambda: stop["v"], max_inflight=16,
                                                        on_error=lambda r, e: errors.append((r.connection_id, str(e)))))
        for _ in range(200):
            await asyncio.sleep(0.02)
            if qm.done == len(reqs):
                break
        stop["v"] = True
        await task
        return qm

    qm = asyncio.run(main())
    dec.close()
    assert qm.done == len(reqs) and errors == [("bad", "empty text")]
    by_conn = {}
    for conn, data, cid, final in qm.sent:
        by_conn.setdefault(conn, []).append((cid, final, data))
    assert "bad" not in by_conn and len(by_conn) == 12
    for i in range(12):
        msgs = by_conn[f"c{i}"]
        assert [(c, f) for c, f, _ in msgs] == [(0, False), (1, True)] and msgs[1][2] == b""
        audio = np.frombuffer(msgs[0][2], dtype=np.float32)
        assert audio.shape == (480 * (3 + i),) and np.all(audio == 3 + i)
    assert sum(seen_batches) >= 12 and max(seen_batches) > 1          # requests really shared batches


def test_per_request_seeds_reach_the_decode_function():
    got = []

    def decode_fn(x, lengths, seeds):
        got.append((list(lengths), list(seeds)))
        return torch.zeros(x.shape[0], x.shape[2] * 480)

    mb = MicroBatcher(decode_fn, max_batch=4, max_wait_ms=1.0, pad_batch=4, per_request_seeds=True)
    f1 = mb.submit(torch.zeros(80, 5), seed=42)
    f1.result(timeout=10)
    f2 = mb.submit(torch.zeros(80, 6))               # no seed given: the batcher's own counter
    f2.result(timeout=10)
    mb.close()
    assert got[0] == ([5], [42]) and got[1][0] == [6] and got[1][1][0] >= 1
    assert isinstance(f1, Future)
