#!/bin/bash
# ncu evidence for the two-pass plan: launch list of one bench run, one `--set full` capture of the lower-bound
# kernel and of the window kernel, and the per-pipe micro-benchmarks.
set -u
O=gpurun_out/lbprof; mkdir -p $O
timeout 200 squigglekit_b200/sqk_ubench full > $O/ubench_full.jsonl 2>&1; echo "ubench rc=$?"; grep -E "pipe" $O/ubench_full.jsonl | tail -12
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sqk_ -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sqk_dtw_lb_kernel -s 3 -c 1 -f -o $O/lb \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1; echo "ncu lb rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sqk_dtw_kernel -s 6 -c 1 -f -o $O/win \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1; echo "ncu win rc=$?"
ls -la $O
