#!/bin/bash
# Round 2, twelfth GPU check (1 GPU): parallel staging of pageable buffers, CLI profile, full default bench line.
set -u
O=gpurun_out/r2l; mkdir -p $O
SQK_CLI_PROFILE=1 timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
timeout 900 python -m pytest tests/test_motifseq_gpu.py tests/test_segmenter_gpu.py tests/test_cli_gpu.py -m gpu -q -x 2>&1 | tail -3
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2l/bench.json").read().strip().split("\n")[-1])
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "pageable", d["e2e"]["pageable"], "lb", round(d["roofline"]["kernel_ms_per_launch"], 3), "win", d["roofline"]["exact_windows_ms_per_step"], "stats", d["roofline"]["stats_kernel_ms_per_launch"], d["parity"])
c = d.get("cli_e2e") or {}
print("cli", {k: (v if not isinstance(v, dict) else {kk: vv for kk, vv in v.items() if kk != "sample"}) for k, v in c.items()})
print("sustained", d["sustained"] and round(d["sustained"]["value"]), "alu", d["roofline_alu"]["frac"], "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"]))
sg = d.get("segmenter") or {}
for r in sg.get("runs", []):
    print("  seg", r["reads"], round(r["value"]), r["kernels_ms"], round(r["roofline"]["frac_step"], 4), "e2e", round(r["e2e"]["value"]), r["parity_subsample_bit_exact"])
PY
