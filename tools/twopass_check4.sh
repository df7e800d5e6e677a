#!/bin/bash
set -u
O=gpurun_out/twopass4; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $O/bench_l4.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
SQK_LB_LANES=8 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $O/bench_l8.json 2> /dev/null; echo "bench l8 rc=$?"
python - <<'PY'
import json
for n in ("l4", "l8"):
    d = json.load(open(f"gpurun_out/twopass4/bench_{n}.json"))
    print(n, "value", round(d["value"]), "e2e", d["e2e"] and round(d["e2e"]["value"]), "kernel", d["roofline"]["kernel"], round(d["roofline"]["kernel_ms_per_launch"], 3),
      "win ms", d["roofline"].get("exact_windows_ms_per_step"), "stats", round(d["roofline"]["stats_kernel_ms_per_launch"], 3), "plan", d.get("plan"), "alu", d["roofline_alu"]["frac"], d["parity"])
PY
