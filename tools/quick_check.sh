#!/bin/bash
# Dev helper: GPU suite + default bench + segmenter bench, one summary line each.
set -u
O=gpurun_out/quick; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/quick/bench.json"))
print("value", round(d["value"]), "e2e", d["e2e"] and round(d["e2e"]["value"]), "kernel", d["roofline"]["kernel"], round(d["roofline"]["kernel_ms_per_launch"], 3),
      "win ms", d["roofline"].get("exact_windows_ms_per_step"), "stats", round(d["roofline"]["stats_kernel_ms_per_launch"], 3), "plan", d.get("plan"), "alu", d["roofline_alu"]["frac"], d["parity"])
PY
python tools/bench_segmenter.py --reads 1000000 --steps 5 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('segmenter 1M value', round(d['value']), d['kernels_ms'], 'frac', round(d['roofline']['frac_step'], 4), d.get('parity_subsample_bit_exact'))"
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --samples 20000 --reads 200000 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().split('\n')[-1]); print('200k x 20000 value', round(d['value']), 'lb', round(d['roofline']['kernel_ms_per_launch'], 2), 'stats', round(d['roofline']['stats_kernel_ms_per_launch'], 2), d['parity'])"
