#!/bin/bash
# Round 2, sixth GPU check (1 GPU): lower-bound kernel variants (2 / 4 columns per step, one candidate test per two steps,
# per-step free-start row, cheaper refill), suite.
set -u
O=gpurun_out/r2f; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_gpu.log
for v in 0 2 4; do
  SQK_LB_COLS=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $O/bench_cols$v.json 2> $O/bench_cols$v.err; echo "cols $v rc=$?"
done
SQK_LB_COLS=4 SQK_LB_LANES=4 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $O/bench_cols4_l4.json 2> $O/bench_cols4_l4.err
SQK_LB_COLS=4 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-extras --motif-len 163 > $O/bench_cols4_n163.json 2> $O/bench_cols4_n163.err
SQK_LB_COLS=2 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-extras --motif-len 163 > $O/bench_cols2_n163.json 2> $O/bench_cols2_n163.err
SQK_LB_COLS=4 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extras --seed-offset 3 > $O/bench_cols4_seed3.json 2> $O/bench_cols4_seed3.err
python - <<'PY'
import json
for f in ("bench_cols0", "bench_cols2", "bench_cols4", "bench_cols4_l4", "bench_cols4_n163", "bench_cols2_n163", "bench_cols4_seed3"):
    try:
        d = json.loads(open(f"gpurun_out/r2f/{f}.json").read().strip().split("\n")[-1])
        print(f, "value", round(d["value"]), "lb", round(d["roofline"]["kernel_ms_per_launch"], 3), "win", round(d["roofline"]["exact_windows_ms_per_step"], 3), "stats", round(d["roofline"]["stats_kernel_ms_per_launch"], 3), d["parity"], d["plan"])
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sqk_dtw_lb_kernel -s 3 -c 1 -f -o $O/lb_cols4 \
    env SQK_LB_COLS=4 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1; echo "ncu lb rc=$?"
