#!/bin/bash
# Round 2, fourth GPU check (1 GPU): stats3 (single staging buffer), long window jobs first (rank-3 batch on one GPU),
# rolling-mean adapter finder.
set -u
O=gpurun_out/r2d; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 $O/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extras --seed-offset 3 > $O/bench_seed3.json 2> $O/bench_seed3.err; echo "bench seed3 rc=$?"
timeout 300 python tools/bench_segmenter.py --reads 10000 1000000 --steps 5 > $O/seg.jsonl 2> $O/seg.err; echo "seg rc=$?"; tail -2 $O/seg.err
python - <<'PY'
import json
for f in ("bench", "bench_seed3"):
    try:
        d = json.loads(open(f"gpurun_out/r2d/{f}.json").read().strip().split("\n")[-1])
        print(f, "value", round(d["value"]), "e2e", d["e2e"] and round(d["e2e"]["value"]), "lb", round(d["roofline"]["kernel_ms_per_launch"], 3),
              "win", d["roofline"].get("exact_windows_ms_per_step"), "stats", round(d["roofline"]["stats_kernel_ms_per_launch"], 3), d["parity"], "launches", d["gpu_launches"], d["plan"])
    except Exception as e:
        print(f, "unreadable", e)
for f in ("seg",):
    try:
        for ln in open(f"gpurun_out/r2d/{f}.jsonl"):
            d = json.loads(ln); print(f, d["reads"], "value", round(d["value"]), d["kernels_ms"], "frac", round(d["roofline"]["frac_step"], 4), "e2e", round(d["e2e"]["value"]), d.get("parity_subsample_bit_exact"))
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sqk_stats3_kernel -s 3 -c 1 -f -o $O/stats3_zscale \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1; echo "ncu stats3 zscale rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sqk_stats3_kernel -s 3 -c 1 -f -o $O/stats3_seg \
    python tools/bench_segmenter.py --reads 1000000 --steps 1 > /dev/null 2>&1; echo "ncu stats3 seg rc=$?"
