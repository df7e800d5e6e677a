#!/bin/bash
# Dev helper: stats-kernel variants (threads per read) on the MotifSeq and segmenter workloads.
for nt in 32 128; do
  echo "== SQK_STATS_NT=$nt"
  SQK_STATS_NT=$nt python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null > /tmp/a.json
  python - <<'PY'
import json; d=json.load(open("/tmp/a.json")); print("motifseq value", round(d["value"]), "stats ms", d["roofline"]["stats_kernel_ms_per_launch"], d["parity"]["indices_bit_exact"])
PY
  SQK_STATS_NT=$nt python tools/bench_segmenter.py --reads 1000000 --steps 5 2>/dev/null > /tmp/s.json
  python - <<'PY'
import json; d=json.loads(open("/tmp/s.json").read().strip().split("\n")[-1]); print("segmenter 1M value", round(d["value"]), d["kernels_ms"], "frac", round(d["roofline"]["frac_step"],4), d["parity_subsample_bit_exact"])
PY
done
