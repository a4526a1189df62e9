"""Request-level data parallelism: independent utterances / streams are partitioned across GPUs.

The reference deploys one service process per GPU and load-balances connections between them
(services/tts/server.py:397-400, :485-488); nothing is exchanged between GPUs, so there is no
collective on the data path.  `shard_range` is the partition both bench.py (one rank per GPU under
torchrun) and ShardedDecoder (one thread + handle per GPU in one process) use."""
from __future__ import annotations

import threading
from typing import Dict, List, Optional, Sequence, Tuple

import torch


def shard_range(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of n_items for `rank` of `world` (sizes differ by <= 1)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def round_robin(n_items: int, world: int) -> List[List[int]]:
    """Stream i goes to GPU i mod world (the per-connection placement a load balancer gives)."""
    return [list(range(r, n_items, world)) for r in range(world)]


class ShardedDecoder:
    """One B200HiFT per visible GPU, each driven by its own host thread.  decode_many() splits a list
    of utterances across them and returns results in input order."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], devices: Optional[Sequence[int]] = None,
                 dtype: str = "bf16"):
        from .decoder import B200HiFT

        if devices is None:
            devices = list(range(torch.cuda.device_count()))
        if not devices:
            raise RuntimeError("ShardedDecoder needs at least one CUDA device")
        self.devices = list(devices)
        self.decoders = [B200HiFT(state_dict, device=f"cuda:{d}", dtype=dtype) for d in self.devices]

    def decode_many(self, mels: Sequence[torch.Tensor], batch: int = 64) -> List[torch.Tensor]:
        """mels: host tensors [80, T_i].  Returns host fp32 wavs [480*T_i] in the same order."""
        G = len(self.decoders)
        plan = round_robin(len(mels), G)
        out: List[Optional[torch.Tensor]] = [None] * len(mels)
        errs: List[BaseException] = []

        def work(g: int):
            try:
                dec = self.decoders[g]
                dev = dec.device
                with torch.cuda.device(dev):
                    idx = plan[g]
                    for b0 in range(0, len(idx), batch):
                        ids = idx[b0:b0 + batch]
                        Tm = max(mels[i].shape[-1] for i in ids)
                        x = torch.zeros(len(ids), 80, Tm, dtype=torch.float32)
                        for r, i in enumerate(ids):
                            x[r, :, : mels[i].shape[-1]] = mels[i]
                        lens = torch.tensor([mels[i].shape[-1] for i in ids], dtype=torch.int32)
                        wav, _ = dec.inference(x.to(dev, non_blocking=True), lengths=lens.to(dev))
                        wav = wav.cpu()
                        for r, i in enumerate(ids):
                            out[i] = wav[r, : 480 * mels[i].shape[-1]].clone()
            except BaseException as e:  # surfaced to the caller below
                errs.append(e)

        threads = [threading.Thread(target=work, args=(g,)) for g in range(G)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errs:
            raise errs[0]
        return out  # type: ignore[return-value]
