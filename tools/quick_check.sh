#!/bin/bash
# Dev helper: GPU suite + default bench, one summary line.
set -u
O=gpurun_out/quick; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 60 squigglekit_b200/sqk_ubench 2>&1 | grep lb_step | head -4
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/quick/bench.json"))
print("value", round(d["value"]), "e2e", d["e2e"] and round(d["e2e"]["value"]), "kernel", d["roofline"]["kernel"], round(d["roofline"]["kernel_ms_per_launch"], 3),
      "win ms", d["roofline"].get("exact_windows_ms_per_step"), "stats", round(d["roofline"]["stats_kernel_ms_per_launch"], 3), "plan", d.get("plan"), "alu", d["roofline_alu"]["frac"], d["parity"])
PY
