"""CPU: the N>1 host logic — block partition and the score gather — on world_size-2/3 gloo process groups."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from runtime.sharding import ShardPlan, gather_scores, score_clips_sharded, shard_range


def test_shard_range_covers_everything_once():
    for n in (0, 1, 7, 8, 10000, 10001):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(10000, 3, 8) == (3750, 5000)  # BASELINE cfg5: 1250 clips per GPU
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _fake_frame_scores(lo, hi, t):
    """Deterministic stand-in for model.get_reconstruction_error(clips[lo:hi], per_frame=True)."""
    idx = torch.arange(lo, hi, dtype=torch.float32).view(-1, 1)
    return (idx * 0.001 + torch.arange(t, dtype=torch.float32).view(1, -1) * 1e-5 + 0.25).contiguous()


def _worker(rank, world, n_items, t, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = ShardPlan(n_items, world)
        lo, hi = plan.range(rank)
        local = _fake_frame_scores(lo, hi, t)
        full = gather_scores(local, plan, rank, dst=0)
        if rank == 0:
            torch.save(full, out_path)
        else:
            assert full is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_items", [(2, 10), (2, 7), (3, 8)])
def test_gather_scores_matches_single_process(tmp_path, world, n_items):
    t = 4
    out = str(tmp_path / "gathered.pt")
    port = 29500 + (os.getpid() + world * 7 + n_items) % 2000
    mp.spawn(_worker, args=(world, n_items, t, port, out), nprocs=world, join=True)
    got = torch.load(out)
    assert torch.equal(got, _fake_frame_scores(0, n_items, t)), "sharded + gathered must equal the 1-process result"


def test_single_rank_is_identity():
    x = torch.arange(6.0).view(3, 2)
    assert gather_scores(x, ShardPlan(3, 1), 0) is x


def _make_clip(clip_id, t=3):
    g = torch.Generator().manual_seed(1234 + clip_id)  # seed = 1234 + clip_id, independent of the sharding
    return torch.rand(t, 3, 4, 4, generator=g)


def _score_clip(x):
    return (x * x).mean(dim=(1, 2, 3))


def _job_worker(rank, world, n_clips, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = score_clips_sharded(n_clips, _make_clip, _score_clip, rank, world)
        if rank == 0:
            torch.save(full, out_path)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_clips", [(2, 9), (3, 2)])
def test_score_clips_sharded_is_sharding_invariant(tmp_path, world, n_clips):
    """The config-5 job (per-clip seeds, contiguous shards, one gather) gives bit-identical [n_clips, T] scores for any
    world size — including a rank with an empty shard."""
    out = str(tmp_path / "job.pt")
    port = 31500 + (os.getpid() + world * 11 + n_clips) % 2000
    mp.spawn(_job_worker, args=(world, n_clips, port, out), nprocs=world, join=True)
    got = torch.load(out)
    ref = score_clips_sharded(n_clips, _make_clip, _score_clip, 0, 1)
    assert got.shape == (n_clips, 3) and torch.equal(got, ref)
