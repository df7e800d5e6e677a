#!/bin/bash
# Round 2, two-GPU check: the NCCL / P2P shard-invariance test, bench at N=2 with both exchanges.
set -u
O=gpurun_out/r2n2; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q -rs > $O/pytest_multigpu.log 2>&1; echo "pytest multigpu rc=$?"; tail -15 $O/pytest_multigpu.log
for ex in p2p nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --exchange $ex > $O/bench_n2_$ex.json 2> $O/bench_n2_$ex.err; echo "bench n2 $ex rc=$?"; tail -3 $O/bench_n2_$ex.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 1 --impl reference > $O/bench_n2_ref.json 2> $O/bench_n2_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
for f in ("bench_n2_p2p", "bench_n2_nccl", "bench_n2_ref"):
    try:
        d = json.loads(open(f"gpurun_out/r2n2/{f}.json").read().strip().split("\n")[-1])
        print(f, "value", round(d["value"]), "e2e", d["e2e"] and round(d["e2e"]["value"]), d.get("parity"), d.get("per_rank"), d["cpu_baseline"] and d["cpu_baseline"].get("cores"))
    except Exception as e:
        print(f, "unreadable", e)
PY
