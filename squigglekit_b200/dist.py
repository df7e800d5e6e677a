"""Multi-GPU plumbing: reads are independent, so they shard as contiguous blocks over ranks
(one process per GPU) and the only exchange is ONE all-gather of the fixed-size hit records at
the end.  torch.distributed is the transport (NCCL over NVLink on the GPU box,
gloo in the CPU tests); there is no collective inside the data path.
"""
from __future__ import annotations

import os


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_bounds(n_items: int, rank: int, world: int):
    """Contiguous block [lo, hi) of rank `rank`: sizes differ by at most one, order preserved."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allgather_records(local, counts=None):
    """All-gather per-rank record tensors ([n_local, ...], same trailing shape and dtype) into one
    tensor ordered by rank -- i.e. the original read order.  `counts`: per-rank n_local when the
    shards are uneven (shard_bounds); None = all equal (one all_gather_into_tensor)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    if world == 1:
        return local
    if counts is None or len(set(counts)) == 1:
        out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    biggest = max(counts)
    padded = torch.zeros((biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty((world * biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded)
    parts = [out[r * biggest: r * biggest + counts[r]] for r in range(world)]
    return torch.cat(parts, dim=0)
