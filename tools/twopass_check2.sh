#!/bin/bash
set -u
O=gpurun_out/twopass2; mkdir -p $O
timeout 600 python -m pytest tests/test_twopass_gpu.py -x -q > $O/pytest_twopass.log 2>&1; echo "twopass tests rc=$?"; tail -5 $O/pytest_twopass.log
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
timeout 120 squigglekit_b200/sqk_ubench > $O/ubench.jsonl 2>&1; echo "ubench rc=$?"; head -4 $O/ubench.jsonl
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $O/bench_lb_l4.json 2> $O/bench_lb_l4.err; echo "bench rc=$?"; tail -3 $O/bench_lb_l4.err
SQK_LB_LANES=8 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $O/bench_lb_l8.json 2> /dev/null; echo "bench l8 rc=$?"
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --motif-len 163 > $O/bench_n163.json 2> /dev/null; echo "bench n163 rc=$?"
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --samples 20000 --reads 50000 > $O/bench_20k.json 2> /dev/null; echo "bench 20k rc=$?"
python - <<'PY'
import json
for n in ("lb_l4", "lb_l8", "n163", "20k"):
    try:
        d = json.load(open(f"gpurun_out/twopass2/bench_{n}.json"))
        print(n, "value", round(d["value"]), "e2e", d["e2e"] and round(d["e2e"]["value"]), "kernel", d["roofline"]["kernel"], round(d["roofline"]["kernel_ms_per_launch"], 3),
              "win ms", d["roofline"].get("exact_windows_ms_per_step"), "stats", round(d["roofline"]["stats_kernel_ms_per_launch"], 3),
              "plan", d.get("plan"), "alu", d["roofline_alu"]["frac"], d["parity"])
    except Exception as e:
        print(n, "failed", e)
PY
