#!/bin/bash
# Round-end evidence run (one GPU): full GPU test suite, both bench arms, the ncu launch list and one
# `ncu --set full` capture each of the lower-bound kernel, the exact window kernel and the stats kernel,
# the register-only micro-benchmarks, the single-pass plan on the same box.  Outputs land in gpurun_out/final/.
set -u
O=gpurun_out/final; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" ; tail -2 $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
timeout 300 python bench.py --plan single_pass --no-cpu-baseline > $O/bench_single_pass.json 2> /dev/null; echo "single-pass rc=$?"
timeout 200 squigglekit_b200/sqk_ubench full > $O/ubench_full.jsonl 2>&1; echo "ubench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sqk_ -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sqk_dtw_lb_kernel -s 3 -c 1 -f -o $O/lb \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1; echo "ncu lb rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sqk_dtw_kernel -s 6 -c 1 -f -o $O/win \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1; echo "ncu win rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sqk_stats_kernel -s 3 -c 1 -f -o $O/stats \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1; echo "ncu stats rc=$?"
python tools/bench_segmenter.py --reads 1000000 --steps 5 2>/dev/null | tail -1 > $O/segmenter_1M.json; echo "seg rc=$?"
cat $O/bench_n1.json
