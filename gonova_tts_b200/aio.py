"""asyncio side of the micro-batcher: the replacement for the service's single-request worker.

The reference runs ONE `_tts_worker` coroutine: `request = await queue_manager.get_next_request()`, synthesise it to
the end, `enqueue_audio_chunk(...)` per chunk, a final empty chunk, `mark_request_done()`
(services/tts/server.py:110-186, services/tts/core/queue_manager.py:173-193).  One sentence at a time leaves the
decoder at batch 1 (about 9 k x real time instead of 23 k x at batch 64).  `batched_tts_worker` keeps the same
queue-manager calls and message order per connection but lets up to `max_inflight` requests be in flight: each runs its
front end (text -> mel: the engine's T3 + flow, `mel_fn`, in the executor) and then awaits its waveform from ONE
MicroBatcher, where the requests that arrive within the window share a ragged batch on the GPU.

Nothing here touches CUDA; the only GPU work is inside the batcher's decode function."""
from __future__ import annotations

import asyncio
from typing import Awaitable, Callable, Optional

import numpy as np
import torch

from .batching import MicroBatcher


class AsyncBatchedDecoder:
    """`await decode(mel)` for coroutines: `MicroBatcher.submit` futures as asyncio awaitables."""

    def __init__(self, batcher: MicroBatcher):
        self.batcher = batcher

    async def decode(self, mel: torch.Tensor, seed: Optional[int] = None) -> np.ndarray:
        """mel [80, T] host tensor -> float32 numpy [480 T] (what `_synthesize_sync` returns,
        services/tts/core/synthesizer.py:352-357).  queue.Full propagates when the batcher's bounded queue is full —
        the caller drops the request like queue_manager.enqueue_request does (queue_manager.py:157-171)."""
        fut = self.batcher.submit(mel, seed)
        wav = await asyncio.wrap_future(fut)
        return wav.numpy()

    def close(self):
        self.batcher.close()


async def batched_tts_worker(queue_manager, mel_fn: Callable[[object], torch.Tensor], decoder: AsyncBatchedDecoder,
                             is_shutting_down: Callable[[], bool], max_inflight: int = 64,
                             on_error: Optional[Callable[[object, BaseException], None]] = None) -> None:
    """Drop-in body for `TTSService._tts_worker` (services/tts/server.py:110-186).

    queue_manager: the reference's QueueManager (uses get_next_request / enqueue_audio_chunk / mark_request_done only).
    mel_fn(request) -> mel [80, T] host tensor: the engine's front end for one request (blocking; run in the default
    executor like `_synthesize_sync`, synthesizer.py:312-318).
    Per connection the protocol is the reference's: audio chunks with increasing chunk_id, then an empty final chunk."""
    loop = asyncio.get_running_loop()
    sem = asyncio.Semaphore(max_inflight)
    tasks = set()

    async def one(request):
        try:
            mel = await loop.run_in_executor(None, mel_fn, request)
            audio = await decoder.decode(mel)
            await queue_manager.enqueue_audio_chunk(request.connection_id, audio.astype(np.float32, copy=False).tobytes(),
                                                    0, is_final=False)
            await queue_manager.enqueue_audio_chunk(request.connection_id, b"", 1, is_final=True)
        except BaseException as e:             # one failed request must not take the worker down (server.py:173-179)
            if on_error is not None:
                on_error(request, e)
            if isinstance(e, asyncio.CancelledError):
                raise
        finally:
            await queue_manager.mark_request_done()
            sem.release()

    while not is_shutting_down():
        request = await queue_manager.get_next_request()
        if request is None:
            continue
        await sem.acquire()
        t = asyncio.ensure_future(one(request))
        tasks.add(t)
        t.add_done_callback(tasks.discard)
    if tasks:
        await asyncio.gather(*tasks, return_exceptions=True)
