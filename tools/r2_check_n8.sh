#!/bin/bash
# Round 2, eight-GPU check: default workload with both exchanges (per-rank kernel times in the line), N=4, and
# BASELINE configs[4] (10 M x 50 000 over 8 GPUs = 1.25 M reads = 125 GB resident per GPU) with the gather at the end.
set -u
O=gpurun_out/r2n8; mkdir -p $O
run() { # name nproc args...
  local name=$1 np=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $np "$@" > $O/$name.json 2> $O/$name.err; echo "$name rc=$?"; tail -2 $O/$name.err
}
run bench_n8_p2p 8 --steps 20 --warmup 3 --exchange p2p
run bench_n8_nccl 8 --steps 20 --warmup 3 --exchange nccl --no-extras --no-e2e
run bench_n8_none 8 --steps 20 --warmup 3 --exchange none --no-extras --no-e2e
run bench_n4_p2p 4 --steps 20 --warmup 3 --exchange p2p
run bench_n8_config4 8 --workload configs4 --steps 2 --warmup 3 --exchange p2p
python - <<'PY'
import json
for f in ("bench_n8_p2p", "bench_n8_nccl", "bench_n8_none", "bench_n4_p2p", "bench_n8_config4"):
    try:
        d = json.loads(open(f"gpurun_out/r2n8/{f}.json").read().strip().split("\n")[-1])
        print(f, "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", d["e2e"] and round(d["e2e"]["value"]), d.get("parity"), d.get("plan"))
        print("   ", d.get("per_rank"))
    except Exception as e:
        print(f, "unreadable", e)
PY
