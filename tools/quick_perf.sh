#!/bin/bash
# Dev helper: one-line perf summary of both paths (MotifSeq bench + segmenter at 1M reads).
python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null > /tmp/a.json
python - <<'PY'
import json; d=json.load(open("/tmp/a.json")); print("motifseq value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "dtw ms", round(d["roofline"]["kernel_ms_per_launch"],3), "stats ms", round(d["roofline"]["stats_kernel_ms_per_launch"],3), d["parity"]["indices_bit_exact"], d["parity"]["dist_bit_exact"])
PY
python tools/bench_segmenter.py --reads 1000000 --steps 5 2>/dev/null > /tmp/s.json
python - <<'PY'
import json; d=json.loads(open("/tmp/s.json").read().strip().split("\n")[-1]); print("segmenter 1M value", round(d["value"]), d["kernels_ms"], "frac", round(d["roofline"]["frac_step"],4), d["parity_subsample_bit_exact"])
PY
