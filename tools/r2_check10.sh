#!/bin/bash
# Round 2, tenth GPU check (1 GPU): rare stats paths, suite, segmenter bench.
set -u
O=gpurun_out/r2j; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_gpu.log
timeout 300 python tools/bench_segmenter.py --reads 10000 1000000 --steps 5 > $O/seg.jsonl 2> $O/seg.err; echo "seg rc=$?"; tail -2 $O/seg.err
python - <<'PY'
import json
for f in ("seg",):
    try:
        for ln in open(f"gpurun_out/r2j/{f}.jsonl"):
            d = json.loads(ln); print(f, d["reads"], "value", round(d["value"]), d["kernels_ms"], "frac", round(d["roofline"]["frac_step"], 4), "e2e", round(d["e2e"]["value"]), d.get("parity_subsample_bit_exact"))
    except Exception as e:
        print(f, "unreadable", e)
PY
