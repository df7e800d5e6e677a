#!/bin/bash
# First GPU run of the two-pass DTW plan: new tests, full GPU suite, register-only micro-benchmarks, bench both plans.
set -u
O=gpurun_out/twopass; mkdir -p $O
timeout 600 python -m pytest tests/test_twopass_gpu.py -x -q > $O/pytest_twopass.log 2>&1; echo "twopass tests rc=$?"; tail -25 $O/pytest_twopass.log
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_gpu.log
timeout 120 squigglekit_b200/sqk_ubench > $O/ubench.jsonl 2>&1; echo "ubench rc=$?"; head -8 $O/ubench.jsonl
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $O/bench_two_pass.json 2> $O/bench_two_pass.err; echo "bench two_pass rc=$?"; tail -3 $O/bench_two_pass.err
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --plan single_pass > $O/bench_single_pass.json 2> $O/bench_single_pass.err; echo "bench single rc=$?"
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --lanes 4 > $O/bench_two_pass_l4.json 2> /dev/null; echo "bench l4 rc=$?"
python - <<'PY'
import json
for n in ("two_pass", "single_pass", "two_pass_l4"):
    try:
        d = json.load(open(f"gpurun_out/twopass/bench_{n}.json"))
        print(n, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "kernel", d["roofline"]["kernel"], round(d["roofline"]["kernel_ms_per_launch"], 3),
              "win ms", d["roofline"].get("exact_windows_ms_per_step"), "stats", round(d["roofline"]["stats_kernel_ms_per_launch"], 3),
              "plan", d.get("plan"), "alu", d["roofline_alu"]["frac"], d["parity"])
    except Exception as e:
        print(n, "failed", e)
PY
