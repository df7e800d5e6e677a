#!/bin/bash
# Round 2, eighth GPU check (1 GPU): in-place patching in stats3 (no patch area), long motifs (row blocks), suite.
set -u
O=gpurun_out/r2h; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
timeout 300 python tools/bench_segmenter.py --reads 10000 1000000 --steps 5 > $O/seg.jsonl 2> $O/seg.err; echo "seg rc=$?"; tail -2 $O/seg.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-e2e --reads 2000 --motif-len 2000 --plan single_pass > $O/bench_n2000.json 2> $O/bench_n2000.err; echo "bench n2000 rc=$?"; tail -2 $O/bench_n2000.err
python - <<'PY'
import json
for f in ("bench", "bench_n2000"):
    try:
        d = json.loads(open(f"gpurun_out/r2h/{f}.json").read().strip().split("\n")[-1])
        print(f, "value", round(d["value"]), "e2e", d["e2e"] and round(d["e2e"]["value"]), "lb", round(d["roofline"]["kernel_ms_per_launch"], 3),
              "win", d["roofline"].get("exact_windows_ms_per_step"), "stats", round(d["roofline"]["stats_kernel_ms_per_launch"], 3), d["parity"], "launches", d["gpu_launches"], d["plan"])
    except Exception as e:
        print(f, "unreadable", e)
for f in ("seg",):
    try:
        for ln in open(f"gpurun_out/r2h/{f}.jsonl"):
            d = json.loads(ln); print(f, d["reads"], "value", round(d["value"]), d["kernels_ms"], "frac", round(d["roofline"]["frac_step"], 4), "e2e", round(d["e2e"]["value"]), d.get("parity_subsample_bit_exact"))
    except Exception as e:
        print(f, "unreadable", e)
PY
