"""What a gonova-tts maintainer changes in services/tts/core/synthesizer.py to stream INSIDE a sentence.

The shipped `StreamingSynthesizer._generate_sentence` (synthesizer.py:296-321) awaits the whole sentence
(`_synthesize_sync` -> `model.generate`) and yields it as one chunk.  With the B200 decoder installed
(`gonova_tts_b200.service.install`, right after synthesizer.py:185) the vocoder inside that same `generate` call can
hand out its PCM 2 s at a time: `chunk_tap` makes the engine's single `mel2wav.inference(...)` call decode chunk by
chunk and pass every chunk to a sink while `generate` is still running.  The method below is the patched
`_generate_sentence`: same signature, same executor thread, same float32 numpy chunks on the way out
(`server.py:150-155` keeps calling `.tobytes()` on them) — only more of them, and the first one after the front end
plus ~0.6 ms instead of after the whole sentence.

This file is an EXAMPLE of reference-side code (it imports nothing from the reference; tests/test_gpu_service.py drives
it with a stand-in engine that has the reference's call shape).  `StreamingSynthesizerPatch` is a mix-in: the real class
is `class StreamingSynthesizer(StreamingSynthesizerPatch, <original>)` or the two methods pasted in."""
from __future__ import annotations

import asyncio
from typing import AsyncGenerator, Optional

import numpy as np

from gonova_tts_b200.service import chunk_tap, install, install_flow, patch_get_stats


class StreamingSynthesizerPatch:
    """Expects the attributes the reference class has: `self.model` (the engine), `self.device`."""

    decoder = None

    def install_b200_decoder(self, dtype: str = "bf16", flow: bool = False):
        """Call in load(), after `self.model = ChatterboxTTS.from_pretrained(device=self.device)` (synthesizer.py:185)
        and before the warm-up loop (:199-207).  `flow=True` also moves the step in front of the vocoder onto the B200
        kernels: `s3gen.flow.decoder` (ten Euler steps of the CFM estimator) and, when the flow module holds them, the token
        embedding + Conformer encoder behind `s3gen.flow.inference` — tokens -> PCM without a stock-PyTorch op."""
        self.decoder = install(self.model, dtype=dtype)
        if flow:
            self.flow_decoder = install_flow(self.model, dtype=dtype)
        patch_get_stats(self, self.decoder)          # get_stats()["decoder"] (synthesizer.py:411-420)
        return self.decoder

    async def _generate_sentence(self, sentence: str, voice_embedding: Optional[str],
                                 exaggeration: float) -> AsyncGenerator[np.ndarray, None]:
        """Replaces synthesizer.py:296-321.  Yields float32 numpy chunks of <= 2 s while `generate` is still running."""
        loop = asyncio.get_running_loop()
        q: asyncio.Queue = asyncio.Queue()

        def sink(pcm: bytes, chunk_id: int, is_last: bool):          # runs on the executor thread, inside generate()
            loop.call_soon_threadsafe(q.put_nowait, np.frombuffer(pcm, dtype=np.float32))

        def run():
            with chunk_tap(self.decoder, sink, fmt="f32"):
                return self._synthesize_sync(sentence, voice_embedding, exaggeration)

        task = loop.run_in_executor(None, run)
        task.add_done_callback(lambda _f: loop.call_soon_threadsafe(q.put_nowait, None))
        while True:
            chunk = await q.get()
            if chunk is None:
                break
            yield chunk
        await task                                                    # re-raises a synthesis failure (server.py:173-179)
