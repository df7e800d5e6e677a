#!/bin/bash
# Round 2, first GPU check: GPU suite with the second-generation stats kernel + bit-mask state machine, A/B against the
# first generation (SQK_STATS_GEN=1), ncu captures of the new kernels.
set -u
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.log
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
SQK_STATS_GEN=1 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_gen1.json 2> $O/bench_gen1.err; echo "bench gen1 rc=$?"
timeout 300 python tools/bench_segmenter.py --reads 10000 1000000 --steps 5 > $O/seg.jsonl 2> $O/seg.err; echo "seg rc=$?"; tail -2 $O/seg.err
SQK_STATS_GEN=1 timeout 300 python tools/bench_segmenter.py --reads 1000000 --steps 5 > $O/seg_gen1.jsonl 2> $O/seg_gen1.err; echo "seg gen1 rc=$?"
timeout 300 python tools/bench_segmenter.py --reads 200000 --steps 5 --pa > $O/seg_pa.jsonl 2> $O/seg_pa.err; echo "seg pa rc=$?"
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --scale medmad > $O/bench_medmad.json 2> $O/bench_medmad.err; echo "bench medmad rc=$?"
python - <<'PY'
import json
for f in ("bench", "bench_gen1", "bench_medmad"):
    try:
        d = json.loads(open(f"gpurun_out/r2a/{f}.json").read().strip().split("\n")[-1])
        print(f, "value", round(d["value"]), "e2e", d["e2e"] and round(d["e2e"]["value"]), "lb", round(d["roofline"]["kernel_ms_per_launch"], 3),
              "win", d["roofline"].get("exact_windows_ms_per_step"), "stats", round(d["roofline"]["stats_kernel_ms_per_launch"], 3), d["parity"])
    except Exception as e:
        print(f, "unreadable", e)
for f in ("seg", "seg_gen1", "seg_pa"):
    try:
        for ln in open(f"gpurun_out/r2a/{f}.jsonl"):
            d = json.loads(ln); print(f, d["reads"], "value", round(d["value"]), d["kernels_ms"], "frac", round(d["roofline"]["frac_step"], 4), "e2e", round(d["e2e"]["value"]), d.get("parity_subsample_bit_exact"))
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sqk_stats2_kernel -s 3 -c 1 -f -o $O/stats2_zscale \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1; echo "ncu stats2 zscale rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sqk_stats2_kernel -s 3 -c 1 -f -o $O/stats2_seg \
    python tools/bench_segmenter.py --reads 1000000 --steps 1 > /dev/null 2>&1; echo "ncu stats2 seg rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sqk_fsm_mask_kernel -s 3 -c 1 -f -o $O/fsm_mask \
    python tools/bench_segmenter.py --reads 1000000 --steps 1 > /dev/null 2>&1; echo "ncu fsm mask rc=$?"
ls -la $O
