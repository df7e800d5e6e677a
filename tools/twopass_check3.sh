#!/bin/bash
set -u
O=gpurun_out/twopass3; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
for v in "" _mb8 _mb3; do timeout 120 squigglekit_b200/sqk_ubench$v 2>&1 | grep lb_step > $O/ubench$v.jsonl; echo "ubench$v"; cat $O/ubench$v.jsonl; done
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sqk_dtw_lb_kernel -s 3 -c 1 -f -o $O/lb \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1; echo "ncu lb rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/twopass3/bench.json"))
print("value", round(d["value"]), "e2e", d["e2e"] and round(d["e2e"]["value"]), "kernel", d["roofline"]["kernel"], round(d["roofline"]["kernel_ms_per_launch"], 3),
      "win ms", d["roofline"].get("exact_windows_ms_per_step"), "stats", round(d["roofline"]["stats_kernel_ms_per_launch"], 3), "plan", d.get("plan"), "alu", d["roofline_alu"]["frac"], d["parity"])
PY
