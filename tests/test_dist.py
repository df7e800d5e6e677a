"""World-size-2 test (gloo, CPU) of the multi-GPU host logic: contiguous read sharding and the one
all-gather of fixed-size hit records.  The records here are a deterministic function of the read
index, so the gathered tensor must equal the single-process result byte for byte -- the same
"shard determinism" property the GPU test checks with real kernels."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.environ["SQK_ROOT"])
    from squigglekit_b200.dist import GatherPipeline, allgather_records, env_rank_world, shard_bounds

    rank, world, _ = env_rank_world()
    dist.init_process_group("gloo")
    for n_reads in (10, 11, 1, 257):
        full = torch.arange(n_reads * 16, dtype=torch.int64).reshape(n_reads, 1, 16).to(torch.uint8)
        lo, hi = shard_bounds(n_reads, rank, world)
        counts = [shard_bounds(n_reads, r, world)[1] - shard_bounds(n_reads, r, world)[0] for r in range(world)]
        got = allgather_records(full[lo:hi].clone(), counts)
        assert got.shape == full.shape, (got.shape, full.shape)
        assert torch.equal(got, full), n_reads
    even = torch.full((4, 1, 16), rank, dtype=torch.uint8)
    got = allgather_records(even)
    assert got.shape[0] == 4 * world and all(int(got[4 * r, 0, 0]) == r for r in range(world))
    # double-buffered per-step gather (bench.py): results of step i stay intact while step i+1 is written
    pipe = GatherPipeline((3, 1, 16), torch.uint8, torch.device("cpu"), depth=2)
    outs = []
    for step in range(5):
        buf = pipe.local_buffer()
        buf.fill_(10 * step + rank)
        outs.append((step, pipe.submit()))
        if len(outs) == 2:            # the older of the two in-flight results is still valid here
            st, o = outs.pop(0)
            pipe.work[st % 2].wait()
            assert o.shape[0] == 3 * world and all(int(o[3 * r, 0, 0]) == 10 * st + r for r in range(world)), st
    pipe.drain()
    st, o = outs.pop(0)
    assert all(int(o[3 * r, 0, 0]) == 10 * st + r for r in range(world))
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_bounds_cover_everything_in_order():
    sys.path.insert(0, ROOT)
    from squigglekit_b200.dist import shard_bounds
    for n in (0, 1, 7, 8, 100_000, 10_000_001):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_allgather_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), SQK_ROOT=ROOT)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, out
        assert f"rank {rank} ok" in out
