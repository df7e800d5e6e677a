#!/bin/bash
set -u
O=gpurun_out/r2q; mkdir -p $O
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $O/bench_$i.json 2> $O/bench_$i.err; echo "bench rc=$?"
done
timeout 300 python -m pytest tests/test_twopass_gpu.py tests/test_fuzz_gpu.py -m gpu -q -x 2>&1 | tail -3
python - <<'PY'
import json
for f in ("bench_1","bench_2"):
    d = json.loads(open(f"gpurun_out/r2q/{f}.json").read().strip().split("\n")[-1])
    print(f, "value", round(d["value"]), "lb", round(d["roofline"]["kernel_ms_per_launch"], 3), "win", round(d["roofline"]["exact_windows_ms_per_step"],3), "stats", round(d["roofline"]["stats_kernel_ms_per_launch"], 3), d["parity"], d["plan"])
PY
