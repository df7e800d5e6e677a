#!/bin/bash
# Round 2, eleventh GPU check (1 GPU): second-attempt windows + W = 1.33 N + 16, suite, bench (default, rank-3 batch, N=163).
set -u
O=gpurun_out/r2k; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extras --seed-offset 3 > $O/bench_seed3.json 2> $O/bench_seed3.err; echo "bench seed3 rc=$?"
timeout 300 python bench.py --motif-len 163 --steps 5 --no-cpu-baseline --no-extras --no-e2e > $O/bench_n163.json 2> /dev/null; echo "n163 rc=$?"
SQK_LB_WINDOW=96 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $O/bench_w96.json 2> $O/bench_w96.err; echo "bench w96 rc=$?"
python - <<'PY'
import json
for f in ("bench", "bench_seed3", "bench_n163", "bench_w96"):
    try:
        d = json.loads(open(f"gpurun_out/r2k/{f}.json").read().strip().split("\n")[-1])
        print(f, "value", round(d["value"]), "e2e", d["e2e"] and round(d["e2e"]["value"]), "lb", round(d["roofline"]["kernel_ms_per_launch"], 3),
              "win", d["roofline"].get("exact_windows_ms_per_step"), "stats", round(d["roofline"]["stats_kernel_ms_per_launch"], 3), d["parity"], "launches", d["gpu_launches"], d["plan"])
    except Exception as e:
        print(f, "unreadable", e)
PY
