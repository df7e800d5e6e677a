"""Two-rank GPU test (NCCL): contiguous read shards per rank + one all-gather of the hit records must equal the
single-GPU result byte for byte.  Skipped when fewer than two GPUs are visible (the driver's -m gpu run uses
one)."""
import os
import socket
import subprocess
import sys
import textwrap
import time

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.environ["SQK_ROOT"])
    import squigglekit_b200 as sqk
    from squigglekit_b200 import synth
    from squigglekit_b200.dist import allgather_records, env_rank_world, shard_bounds

    rank, world, local = env_rank_world()
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    motif = synth.make_motif()
    n_reads = 1001                                   # uneven shards
    sig, off, _ = synth.motifseq_reads_np(n_reads, 2048, motif)
    lo, hi = shard_bounds(n_reads, rank, world)
    counts = [shard_bounds(n_reads, r, world)[1] - shard_bounds(n_reads, r, world)[0] for r in range(world)]
    ctx = sqk.Context(local)
    dsig = torch.from_numpy(sig).cuda()
    doff = torch.from_numpy(off[lo:hi + 1].copy()).cuda()        # absolute offsets of this rank's block
    hits, _ = ctx.motifseq(dsig, doff, motif, scale="zscale", max_read_len=2048)
    full = allgather_records(hits, counts)
    torch.cuda.synchronize()
    if rank == 0:
        one, _ = ctx.motifseq(dsig, torch.from_numpy(off).cuda(), motif, scale="zscale", max_read_len=2048)
        torch.cuda.synchronize()
        assert full.shape == one.shape and torch.equal(full, one), "sharded result differs from single-GPU result"
    dist.barrier()
    print("rank", rank, "all-gather part done", flush=True)

    # ---- the same through PeerGather: the producing kernels store every record into every rank's buffer (P2P) ----
    # (local_buffer() arms ONE motifseq call; the comparison runs below, on the same context, must not publish -- an
    # earlier version of the library left publication on, and rank 1's 1000-record comparison run stored past the end of
    # rank 0's 1000-record buffer)
    from squigglekit_b200.dist import PeerGather
    n_even = 1000
    per = n_even // world
    pg = PeerGather(ctx, per, 1, torch.device("cuda", local), depth=2)
    for step in range(3):                                  # three steps: both buffers get reused
        out = pg.local_buffer()
        blk = torch.from_numpy(off[rank * per:(rank + 1) * per + 1].copy()).cuda()
        ctx.motifseq(dsig, blk, motif, scale="zscale", max_read_len=2048, out=out, want_kept=False)
        gathered = pg.submit()
    pg.drain()
    torch.cuda.synchronize()
    one, _ = ctx.motifseq(dsig, torch.from_numpy(off[:n_even + 1].copy()).cuda(), motif, scale="zscale", max_read_len=2048)
    torch.cuda.synchronize()
    assert torch.equal(gathered, one), f"rank {rank}: records gathered over P2P differ from the single-GPU result"
    print("rank", rank, "P2P part done", flush=True)
    # single-pass plan and a medmad run publish through the generic kernel / status records
    ctx.set_dtw_plan("single_pass")
    out = pg.local_buffer()
    ctx.motifseq(dsig, blk, motif, scale="medmad", max_read_len=2048, out=out, want_kept=False)
    gathered = pg.submit()
    pg.drain()
    torch.cuda.synchronize()
    got = gathered.clone()
    pg.close()                                             # turns publication off again
    one, _ = ctx.motifseq(dsig, torch.from_numpy(off[:n_even + 1].copy()).cuda(), motif, scale="medmad", max_read_len=2048)
    ctx.set_dtw_plan("auto")
    torch.cuda.synchronize()
    assert torch.equal(got, one), f"rank {rank}: single-pass / medmad records gathered over P2P differ"
    ctx.close()                                            # (as bench.py does: library context gone before the process group)
    dist.barrier()
    print("rank", rank, "ok", flush=True)
    dist.destroy_process_group()
""")


def test_two_rank_shards_equal_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), SQK_ROOT=ROOT)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    # one deadline for both ranks; a rank that outlives it is killed and both outputs are shown (a rank that fails early
    # leaves its peer waiting in a collective or on a P2P flag, so the interesting output is usually the OTHER rank's)
    deadline = time.monotonic() + 240
    outs, timed_out = [], []
    for rank, p in enumerate(procs):
        try:
            outs.append(p.communicate(timeout=max(1.0, deadline - time.monotonic()))[0])
        except subprocess.TimeoutExpired:
            p.kill()
            outs.append(p.communicate()[0])
            timed_out.append(rank)
    report = "\n".join(f"---- rank {r} (exit {p.returncode}) ----\n{o[-3000:]}" for r, (p, o) in enumerate(zip(procs, outs)))
    assert not timed_out, f"rank(s) {timed_out} still running after 240 s\n{report}"
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, report
        assert f"rank {rank} ok" in out, report
