#!/bin/bash
# Round 2, ninth GPU check (1 GPU): medmad through stats3, suite, launch lists of the MotifSeq and segmenter steps.
set -u
O=gpurun_out/r2i; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline --no-e2e > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --no-e2e --scale medmad > $O/bench_medmad.json 2> $O/bench_medmad.err; echo "bench medmad rc=$?"; tail -3 $O/bench_medmad.err
SQK_STATS_GEN=2 timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline --no-e2e --scale medmad > $O/bench_medmad_gen2.json 2> $O/bench_medmad_gen2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sqk_ -c 400 --csv --log-file $O/launches_segmenter.csv python tools/bench_segmenter.py --reads 1000000 --steps 1 > /dev/null 2>&1; echo "ncu seg launches rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sqk_ -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2>&1; echo "ncu bench launches rc=$?"
python - <<'PY'
import json, csv
for f in ("bench", "bench_medmad", "bench_medmad_gen2"):
    try:
        d = json.loads(open(f"gpurun_out/r2i/{f}.json").read().strip().split("\n")[-1])
        print(f, "value", round(d["value"]), "lb", round(d["roofline"]["kernel_ms_per_launch"], 3),
              "win", d["roofline"].get("exact_windows_ms_per_step"), "stats", round(d["roofline"]["stats_kernel_ms_per_launch"], 3), d["parity"], "launches", d["gpu_launches"])
    except Exception as e:
        print(f, "unreadable", e)
rows = [r for r in csv.reader(open("gpurun_out/r2i/launches_segmenter.csv")) if len(r) > 10 and r[0].isdigit()]
for r in rows[-8:]:
    print(r[4][:60], r[-1], r[-2])
PY
