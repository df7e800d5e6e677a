#!/bin/bash
# Round 2, third GPU check (1 GPU): third-generation stats kernel (one warp per read) -- suite, bench, segmenter bench,
# A/B against the second generation (SQK_STATS_GEN=2), ncu captures.
set -u
O=gpurun_out/r2c; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -7 $O/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
SQK_STATS_GEN=2 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $O/bench_gen2.json 2> $O/bench_gen2.err; echo "bench gen2 rc=$?"
timeout 300 python tools/bench_segmenter.py --reads 10000 1000000 --steps 5 > $O/seg.jsonl 2> $O/seg.err; echo "seg rc=$?"; tail -2 $O/seg.err
SQK_STATS_GEN=2 timeout 300 python tools/bench_segmenter.py --reads 1000000 --steps 5 > $O/seg_gen2.jsonl 2> $O/seg_gen2.err; echo "seg gen2 rc=$?"
python - <<'PY'
import json
for f in ("bench", "bench_gen2"):
    try:
        d = json.loads(open(f"gpurun_out/r2c/{f}.json").read().strip().split("\n")[-1])
        print(f, "value", round(d["value"]), "e2e", d["e2e"] and round(d["e2e"]["value"]), "lb", round(d["roofline"]["kernel_ms_per_launch"], 3),
              "win", d["roofline"].get("exact_windows_ms_per_step"), "stats", round(d["roofline"]["stats_kernel_ms_per_launch"], 3), d["parity"], "launches", d["gpu_launches"])
    except Exception as e:
        print(f, "unreadable", e)
for f in ("seg", "seg_gen2"):
    try:
        for ln in open(f"gpurun_out/r2c/{f}.jsonl"):
            d = json.loads(ln); print(f, d["reads"], "value", round(d["value"]), d["kernels_ms"], "frac", round(d["roofline"]["frac_step"], 4), "e2e", round(d["e2e"]["value"]), d.get("parity_subsample_bit_exact"))
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sqk_stats3_kernel -s 3 -c 1 -f -o $O/stats3_zscale \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1; echo "ncu stats3 zscale rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sqk_stats3_kernel -s 3 -c 1 -f -o $O/stats3_seg \
    python tools/bench_segmenter.py --reads 1000000 --steps 1 > /dev/null 2>&1; echo "ncu stats3 seg rc=$?"
