"""Frame-sharded scoring over the GPUs of one box (BASELINE config 5).

Every image and every video window is independent (eval-mode BatchNorm, ConvLSTM state re-zeroed per window —
reference models/video_autoencoder.py:144-145), so the path shards with NO data-path collective: rank r scores the
contiguous block [r*N/P, (r+1)*N/P) of clips and the only exchange is one gather of per-frame fp32 scores to rank 0
(NCCL over NVLink on GPUs, gloo in the CPU tests).  Because the kernels reduce in a fixed order and never mix
frames, the gathered result is bit-identical to scoring everything on one GPU.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block partition; the first (n_items % world) ranks get one extra item."""
    if world <= 0 or not (0 <= rank < world) or n_items < 0:
        raise ValueError(f"bad shard request n={n_items} rank={rank} world={world}")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


@dataclass
class ShardPlan:
    n_items: int
    world: int

    def range(self, rank: int) -> Tuple[int, int]:
        return shard_range(self.n_items, rank, self.world)

    def counts(self) -> List[int]:
        return [hi - lo for lo, hi in (self.range(r) for r in range(self.world))]

    def max_count(self) -> int:
        return max(self.counts()) if self.world else 0


def gather_scores(local: torch.Tensor, plan: ShardPlan, rank: int, dst: int = 0,
                  group=None) -> Optional[torch.Tensor]:
    """Gather per-item score rows [n_local, ...] from every rank to `dst`, in global item order.

    Ranks may hold different item counts (ragged tail): rows are padded to the largest shard for the collective and
    trimmed on `dst`.  Returns the [n_items, ...] tensor on `dst`, None elsewhere.  With world == 1 it is a no-op.
    """
    import torch.distributed as dist
    if plan.world == 1:
        return local
    lo, hi = plan.range(rank)
    if local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} holds {local.shape[0]} rows, plan says {hi - lo}")
    pad = plan.max_count()
    send = local
    if local.shape[0] < pad:
        send = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        send[: local.shape[0]] = local
    send = send.contiguous()
    bufs = [torch.empty_like(send) for _ in range(plan.world)] if rank == dst else None
    dist.gather(send, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([b[:c] for b, c in zip(bufs, plan.counts())], dim=0)
