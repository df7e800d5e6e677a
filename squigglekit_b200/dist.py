"""Multi-GPU plumbing: reads are independent, so they shard as contiguous blocks over ranks
(one process per GPU) and the only exchange is the gather of the fixed-size hit records at the end
of a step.  Two transports:

* :class:`PeerGather` -- the kernels that produce a record store it straight into every peer's
  gathered buffer through P2P-mapped pointers (NVLink): a fused compute + all-gather with no
  collective kernel on the SMs.  torch.distributed only carries the 64-byte IPC handles at set-up.
* :func:`allgather_records` / :class:`GatherPipeline` -- ``all_gather_into_tensor`` (NCCL over NVLink
  on the GPU box, gloo in the CPU tests).
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

    def __init__(self, shape, dtype, device, depth: int = 2, enabled: bool = True):
        import torch
        import torch.distributed as dist
        self.world = dist.get_world_size() if (enabled and dist.is_initialized()) else 1
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

    def close(self):
        self.drain()


class _DevView:
    """Expose a raw device allocation to torch through the CUDA array interface."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerGather:
    """Gather of the hit records of a device-mode ``Context.motifseq`` call without a collective: the kernels that
    produce a record store it into the gathered buffer of every rank (their own through ``out=``, the others through
    P2P-mapped pointers obtained once over CUDA IPC), and one flag per step and peer says when a rank is done.

    ``local_buffer()`` -> the uint8 tensor [n_local, n_models, 16] the next step must pass as ``out=`` (this rank's block
    of its own gathered buffer); ``submit()`` signals the step and returns the gathered tensor [world * n_local,
    n_models, 16] of that step (complete after ``drain()``, or after ``wait(step)``).  ``depth`` buffers rotate: a step
    overwrites the buffer of the step ``depth`` before it, so read a result before the ranks get that far ahead.
    """

    def __init__(self, ctx, n_local: int, n_models: int, device, depth: int = 2):
        import torch
        import torch.distributed as dist
        self.ctx, self.depth, self.n_local, self.n_models = ctx, depth, n_local, n_models
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        if self.world > 16:
            raise ValueError("PeerGather supports up to 16 ranks (one node)")
        self.rec_bytes = 16 * n_models
        self.slot_bytes = self.world * n_local * self.rec_bytes
        self.bufs = [ctx.device_alloc(max(self.slot_bytes, 16)) for _ in range(depth)]
        self.flags = ctx.device_alloc(16 * 8)
        mine = [ctx.ipc_export(p) for p in self.bufs] + [ctx.ipc_export(self.flags)]
        everyone = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(everyone, mine)
        else:
            everyone[0] = mine
        self.opened = []
        self.peer_bufs = [[] for _ in range(depth)]       # per buffer: the OTHER ranks' bases, as seen from this GPU
        flag_ptrs = []
        for q in range(self.world):
            if q == self.rank:
                flag_ptrs.append(self.flags)
                continue
            ptrs = [ctx.ipc_open(h) for h in everyone[q]]
            self.opened += ptrs
            for k in range(depth):
                self.peer_bufs[k].append(ptrs[k])
            flag_ptrs.append(ptrs[depth])
        ctx.set_flag_peers(flag_ptrs, self.rank)
        self.views = [torch.as_tensor(_DevView(p, max(self.slot_bytes, 16)), device=device)[: self.slot_bytes]
                      .view(self.world * n_local, n_models, 16) for p in self.bufs]
        self.i = 0
        if self.world > 1:
            dist.barrier()          # every rank has opened every handle before anyone publishes

    def local_buffer(self):
        k = self.i % self.depth
        self.ctx.set_hit_peers(self.peer_bufs[k], first_record=self.rank * self.n_local,
                               capacity_records=self.world * self.n_local)       # arms the next motifseq call only
        return self.views[k][self.rank * self.n_local:(self.rank + 1) * self.n_local]

    def submit(self):
        k = self.i % self.depth
        self.i += 1
        self.ctx.peer_signal(self.i)
        return self.views[k]

    def wait(self, step: int):
        """Stream-ordered: what follows on the stream sees every rank's records of step `step` (1-based)."""
        self.ctx.peer_wait(step)

    def drain(self):
        if self.i:
            self.ctx.peer_wait(self.i)

    def close(self):
        import torch
        import torch.distributed as dist
        self.ctx.set_hit_peers([], 0)
        self.ctx.set_flag_peers([], 0)
        torch.cuda.synchronize()
        if self.world > 1 and dist.is_initialized():
            dist.barrier()          # nobody unmaps a buffer a peer may still be writing
        self.views = []
        for p in self.opened:
            self.ctx.ipc_close(p)
        self.opened = []
        for p in self.bufs + [self.flags]:
            self.ctx.device_free(p)
        self.bufs = []
