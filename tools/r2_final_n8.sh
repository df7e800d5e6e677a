#!/bin/bash
# Round 2 evidence run (eight GPUs): weak scaling N = 2 / 4 / 8 with the P2P publication, BASELINE configs[4]
# (10 M x 50 000 over 8 GPUs) and the two-rank shard-invariance test.  (N = 8 over NCCL and the reference arm under
# torchrun were measured earlier in the round: profiles/r02_bench_n8_nccl.json, r02_bench_n2_reference_arm.json.)
set -u
O=gpurun_out/r2n8f; mkdir -p $O
run() { # name nproc args...
  local name=$1 np=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $np "$@" > $O/$name.json 2> $O/$name.err; echo "$name rc=$?"; tail -1 $O/$name.err
}
run bench_n8_p2p 8 --steps 20 --warmup 3 --exchange p2p
run bench_n4_p2p 4 --steps 20 --warmup 3 --exchange p2p
run bench_n2_p2p 2 --steps 20 --warmup 3 --exchange p2p
run bench_n8_config4 8 --workload configs4 --steps 2 --warmup 3 --exchange p2p
timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q -rs > $O/pytest_multigpu.log 2>&1; echo "pytest multigpu rc=$?"; tail -3 $O/pytest_multigpu.log
python - <<'PY'
import json
for f in ("bench_n8_p2p", "bench_n8_nccl", "bench_n4_p2p", "bench_n2_p2p", "bench_n8_config4", "bench_n8_reference"):
    try:
        d = json.loads(open(f"gpurun_out/r2n8f/{f}.json").read().strip().split("\n")[-1])
        print(f, "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", d["e2e"] and round(d["e2e"]["value"]), d.get("parity"), d.get("plan"), (d.get("cpu_baseline") or {}).get("cores"))
        print("   ", d.get("per_rank"))
    except Exception as e:
        print(f, "unreadable", e)
PY
