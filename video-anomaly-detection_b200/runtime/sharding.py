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


def score_clips_sharded(n_clips: int, make_clip, score_clip, rank: int, world: int, dst: int = 0, group=None
                        ) -> Optional[torch.Tensor]:
    """The whole config-5 job: rank r scores clips [lo, hi) of `n_clips`, one after another, and the per-frame scores
    are gathered once at the end.

    `make_clip(clip_id)` returns the clip (for synthetic runs: generated on the rank's device from a seed derived from
    `clip_id`, so any sharding sees identical data — 10 k 720p clips are 7 TB of fp32 and are never materialised at
    once); `score_clip(x)` returns its per-frame scores `[T]` (e.g. `model.get_reconstruction_error(x[None],
    per_frame=True)[0]`).  Returns `[n_clips, T]` on rank `dst` (None elsewhere)."""
    plan = ShardPlan(n_clips, world)
    lo, hi = plan.range(rank)
    rows = [score_clip(make_clip(i)).reshape(1, -1) for i in range(lo, hi)]
    if rows:
        local = torch.cat(rows, 0)
    else:  # an empty shard still takes part in the gather; it needs the row width
        local = score_clip(make_clip(0)).reshape(1, -1)[:0]
    return gather_scores(local, plan, rank, dst=dst, group=group)


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Pin this process (and so its future pinned-memory allocations and H2D staging) to the CPUs of the NUMA node the
    GPU hangs off.  One process per GPU uploads its own shard; without the binding half of the ranks of an 8-GPU box
    stage their uploads through the other socket.  Best effort: returns the node, or None when the topology cannot be
    read (then nothing is changed)."""
    import os
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None
