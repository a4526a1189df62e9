"""MicroBatcher (host logic of SURVEY 8f-4) with a fake decoder: no GPU needed."""
import queue
import threading
import time

import pytest
import torch

from gonova_tts_b200.batching import MicroBatcher


class FakeDecoder:
    """wav[b, n] = mean of the utterance's mel (so results identify their request); records every batch."""

    def __init__(self, delay=0.0, fail_on=None):
        self.batches, self.delay, self.fail_on = [], delay, fail_on

    def __call__(self, x, lengths):
        self.batches.append((tuple(x.shape), list(lengths)))
        if self.fail_on is not None and len(self.batches) == self.fail_on:
            raise RuntimeError("decoder exploded")
        time.sleep(self.delay)
        out = torch.zeros(x.shape[0], x.shape[2] * 480)
        for i, n in enumerate(lengths):
            out[i, : n * 480] = x[i, :, :n].mean()
            assert float(x[i, :, n:].abs().sum()) == 0.0          # padding is zeros
        return out


def test_requests_are_batched_and_results_routed():
    dec = FakeDecoder(delay=0.02)
    mb = MicroBatcher(dec, max_batch=8, max_wait_ms=50.0)
    mels = [torch.full((80, 3 + i), float(i)) for i in range(20)]
    futs = [mb.submit(m) for m in mels]
    for i, f in enumerate(futs):
        wav = f.result(timeout=10)
        assert wav.shape == ((3 + i) * 480,) and torch.all(wav == float(i))
    mb.close()
    assert sum(len(l) for _, l in dec.batches) == 20
    assert max(len(l) for _, l in dec.batches) == 8              # never more than max_batch
    assert len(dec.batches) <= 5                                  # ... and it really batched
    for shape, lengths in dec.batches:
        assert shape == (len(lengths), 80, max(lengths))
    m = mb.metrics
    assert m["requests"] == 20 and m["batches"] == len(dec.batches) and m["frames"] == sum(3 + i for i in range(20))
    assert m["padded_frames"] >= m["frames"]


def test_concurrent_producers_and_single_request_latency():
    dec = FakeDecoder()
    mb = MicroBatcher(dec, max_batch=64, max_wait_ms=1.0)
    t0 = time.monotonic()
    assert mb.submit(torch.ones(80, 5)).result(timeout=5).shape == (2400,)   # a lone request waits ~max_wait only
    assert time.monotonic() - t0 < 1.0
    results, errs = {}, []

    def producer(k):
        try:
            results[k] = [mb.submit(torch.full((80, 4), float(k * 10 + j))).result(timeout=10) for j in range(5)]
        except BaseException as e:
            errs.append(e)

    ts = [threading.Thread(target=producer, args=(k,)) for k in range(6)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    mb.close()
    assert not errs
    for k in range(6):
        for j in range(5):
            assert torch.all(results[k][j] == float(k * 10 + j))


def test_full_queue_drops_like_the_reference_and_errors_reach_callers():
    gate = threading.Event()

    def slow(x, lengths):
        gate.wait(5)
        return torch.zeros(x.shape[0], x.shape[2] * 480)

    mb = MicroBatcher(slow, max_batch=1, max_wait_ms=0.0, max_queue=2)
    first = mb.submit(torch.zeros(80, 2))                # taken by the worker, blocks in decode
    time.sleep(0.1)
    mb.submit(torch.zeros(80, 2)); mb.submit(torch.zeros(80, 2))
    with pytest.raises(queue.Full):
        mb.submit(torch.zeros(80, 2))
    assert mb.metrics["dropped"] == 1
    gate.set()
    first.result(timeout=5)
    mb.close()

    dec = FakeDecoder(fail_on=2, delay=0.1)
    mb = MicroBatcher(dec, max_batch=4, max_wait_ms=20.0)
    lone = mb.submit(torch.zeros(80, 2))                  # decoder idle: goes at once, alone (batch 1)
    time.sleep(0.02)
    futs = [mb.submit(torch.zeros(80, 2)) for _ in range(3)]   # arrive while batch 1 is decoding: one batch (2), which fails
    assert lone.result(timeout=5).shape == (960,)
    for f in futs:
        with pytest.raises(RuntimeError, match="exploded"):
            f.result(timeout=5)
    assert [len(l) for _, l in dec.batches] == [1, 3]
    assert mb.submit(torch.ones(80, 2)).result(timeout=5).shape == (960,)   # the worker survived
    mb.close()
    with pytest.raises(RuntimeError):
        mb.submit(torch.zeros(80, 2))
    with pytest.raises(ValueError):
        MicroBatcher(dec).submit(torch.zeros(79, 2))


def test_pad_frames_rounds_the_batch_length_up_and_results_are_unchanged():
    dec = FakeDecoder()
    mb = MicroBatcher(dec, max_batch=4, max_wait_ms=20.0, pad_frames=16)
    futs = [mb.submit(torch.full((80, n), float(n))) for n in (3, 17, 30)]
    for n, f in zip((3, 17, 30), futs):
        wav = f.result(timeout=10)
        assert wav.shape == (n * 480,) and torch.all(wav == float(n))
    mb.close()
    for shape, lengths in dec.batches:
        assert shape[2] % 16 == 0 and shape[2] >= max(lengths) and shape[2] - max(lengths) < 16
    with pytest.raises(ValueError):
        MicroBatcher(dec, pad_frames=0)


def test_a_lone_request_does_not_wait_for_company():
    dec = FakeDecoder()
    mb = MicroBatcher(dec, max_batch=64, max_wait_ms=2000.0)
    t0 = time.monotonic()
    assert mb.submit(torch.ones(80, 5)).result(timeout=5).shape == (2400,)
    assert time.monotonic() - t0 < 1.0                    # max_wait only applies while another batch is decoding
    mb.close()


def test_pad_batch_adds_silent_rows_only():
    dec = FakeDecoder(delay=0.05)
    mb = MicroBatcher(dec, max_batch=8, max_wait_ms=50.0, pad_batch=4)
    first = mb.submit(torch.full((80, 2), 9.0))
    time.sleep(0.01)
    futs = [mb.submit(torch.full((80, 3), float(i))) for i in range(5)]      # gathered while the first one decodes
    assert torch.all(first.result(timeout=5) == 9.0)
    for i, f in enumerate(futs):
        assert torch.all(f.result(timeout=5) == float(i))
    mb.close()
    assert [s[0] for s, _ in dec.batches] == [1, 8]                          # a lone request stays alone, 5 -> 8 rows
    assert dec.batches[0][1] == [2] and dec.batches[1][1] == [3] * 5 + [0] * 3


def test_request_batcher_batches_what_queues_up_and_keeps_order():
    """RequestBatcher (the gathering policy for whole tokens -> PCM requests): while one batch runs, arrivals pile up and ride
    together in the next one; results go back to the right caller; a failure reaches every caller of that batch; a full queue
    drops the request like the reference's."""
    import queue as _q
    import threading
    import time

    from gonova_tts_b200.batching import RequestBatcher

    seen = []
    gate = threading.Event()

    def fn(items):
        seen.append(list(items))
        if len(seen) == 1:
            gate.wait(5)                                # the first batch holds the "GPU" while the others arrive
        if any(i < 0 for i in items):
            raise ValueError("bad request")
        return [i * 10 for i in items]

    rb = RequestBatcher(fn, max_batch=4, max_queue=6)
    first = rb.submit(1)
    time.sleep(0.05)
    rest = [rb.submit(i) for i in (2, 3, 4, 5, 6, 7)]    # 6 queued behind the running batch: the queue is now full
    try:
        rb.submit(8)
        raise AssertionError("a full queue must drop")
    except _q.Full:
        pass
    gate.set()
    assert first.result(5) == 10
    assert [f.result(5) for f in rest] == [20, 30, 40, 50, 60, 70]
    assert seen[0] == [1] and seen[1] == [2, 3, 4, 5] and seen[2] == [6, 7]
    bad = [rb.submit(-1), rb.submit(9)]
    for f in bad:
        try:
            f.result(5)
        except ValueError:
            continue
        # (the two may have landed in different batches: then the second one succeeds)
        assert f.result(5) == 90
    rb.close()
    assert rb.metrics["dropped"] == 1 and rb.metrics["largest_batch"] == 4 and rb.metrics["requests"] == 9
    try:
        rb.submit(1)
        raise AssertionError("closed")
    except RuntimeError:
        pass
