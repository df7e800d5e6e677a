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


class GatherPipeline:
    """Per-step all-gather of fixed-size records that does not make the ranks march in lockstep.

    ``submit(local)`` copies nothing: it starts an asynchronous all-gather of ``local`` into one of ``depth`` output
    buffers and returns (out, work); the caller may enqueue the next step's kernels right away.  A buffer (and the
    ``local`` tensor that fed it -- rotate those, too: ``local_buffer(i)``) is only reused after its gather has been
    waited for, so a rank waits for the others once every ``depth`` steps at most instead of every step: the
    step-to-step skew between GPUs is absorbed instead of being paid at every step.  ``drain()`` waits for everything.
    """

    def __init__(self, shape, dtype, device, depth: int = 2):
        import torch
        import torch.distributed as dist
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.depth = depth
        self.local = [torch.empty(shape, dtype=dtype, device=device) for _ in range(depth)]
        self.out = [torch.empty((self.world * shape[0],) + tuple(shape[1:]), dtype=dtype, device=device) for _ in range(depth)]
        self.work = [None] * depth
        self.i = 0

    def local_buffer(self):
        """The record tensor the next step must write (its previous gather has completed)."""
        k = self.i % self.depth
        if self.work[k] is not None:
            self.work[k].wait()
            self.work[k] = None
        return self.local[k]

    def submit(self):
        import torch.distributed as dist
        k = self.i % self.depth
        self.i += 1
        if self.world == 1:
            return self.local[k]
        self.work[k] = dist.all_gather_into_tensor(self.out[k], self.local[k], async_op=True)
        return self.out[k]

    def drain(self):
        for k in range(self.depth):
            if self.work[k] is not None:
                self.work[k].wait()
                self.work[k] = None
