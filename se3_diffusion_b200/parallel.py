"""Batch sharding of the sampling path across the GPUs of one node.

Backbones are independent — there is no cross-sample operation anywhere on the path — so rank r simply owns the global
samples [first, first + count) and runs the whole 501-forward loop without communication; the only collective is the final
gather of the coordinates.  Noise is counter-based (Philox keyed by (seed, global sample index, step, residue)), so a
sample's trajectory does not depend on the world size.  One process per GPU, torch.distributed (NCCL on GPUs; gloo in the
CPU tests of this host logic).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(global_batch: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of [0, global_batch): returns (first, count) of this rank."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, rem = divmod(global_batch, world_size)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def gather_samples(local: torch.Tensor, global_batch: int) -> torch.Tensor:
    """All-gather per-rank sample tensors [count_r, ...] into [global_batch, ...] in global-sample order (every rank gets it).
    Ranks may own different counts (ragged), so shards are padded to the largest count for the collective."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    counts = [shard_range(global_batch, world, r)[1] for r in range(world)]
    mx = max(counts)
    pad = local.new_zeros((mx,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)


def sample_sharded(engine, global_batch: int, nres: int, **kw):
    """Run engine.sample_device on this rank's shard and gather the final atom37 / rigids of the whole batch."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    first, count = shard_range(global_batch, world, rank)
    if count == 0:      # more ranks than samples: nothing to run here, but this rank still takes part in the gather
        dev = engine.device
        a37, rig, ms, nl = torch.empty(0, nres, 37, 3, device=dev), torch.empty(0, nres, 7, device=dev), 0.0, 0
    else:
        a37, rig, ms, nl = engine.sample_device(count, nres, first_sample=first, **kw)
    return gather_samples(a37, global_batch), gather_samples(rig, global_batch), ms, nl
