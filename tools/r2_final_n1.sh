#!/bin/bash
# Round 2 evidence run (one GPU): GPU suite, both bench arms, other shapes, ncu launch lists and `ncu --set full` captures of
# every kernel of the two hot paths, compute-sanitizer over the new kernels.  Outputs land in gpurun_out/r2final/.
set -u
O=gpurun_out/r2final; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/smi.txt
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu.log
SQK_CLI_PROFILE=1 timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -2 $O/bench_n1.err
timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
timeout 300 python bench.py --plan single_pass --no-cpu-baseline --no-extras > $O/bench_single_pass.json 2> /dev/null; echo "single-pass rc=$?"
timeout 300 python bench.py --scale medmad --no-cpu-baseline --no-extras --no-e2e > $O/bench_medmad.json 2> /dev/null; echo "medmad rc=$?"
timeout 300 python bench.py --motif-len 163 --steps 5 --no-cpu-baseline --no-extras --no-e2e > $O/bench_n163.json 2> /dev/null; echo "n163 rc=$?"
timeout 300 python bench.py --reads 2000 --motif-len 2000 --steps 3 --no-cpu-baseline --no-extras --no-e2e > $O/bench_motif2000.json 2> /dev/null; echo "motif2000 rc=$?"
timeout 900 python bench.py --workload configs3 --steps 2 --no-cpu-baseline --no-extras > $O/bench_config3.json 2> $O/bench_config3.err; echo "config3 rc=$?"; tail -2 $O/bench_config3.err
timeout 300 python tools/bench_segmenter.py --reads 10000 1000000 --steps 5 > $O/segmenter.jsonl 2> /dev/null; echo "seg rc=$?"
timeout 300 python tools/bench_segmenter.py --reads 200000 --steps 5 --pa > $O/segmenter_pa.jsonl 2> /dev/null; echo "seg pa rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sqk_ -c 400 --csv --log-file $O/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $O/under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sqk_ -c 40 --csv --log-file $O/launches_segmenter.csv \
    python tools/bench_segmenter.py --reads 1000000 --steps 1 --no-e2e > /dev/null 2>&1; echo "seg launch list rc=$?"
cap() { # name regex skip cmd...   (the .ncu-rep is summarised on the box and removed: gpurun brings back <= 64 MiB)
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o /tmp/$name "$@" > /dev/null 2>&1; echo "ncu $name rc=$?"
  python tools/ncu_summary.py /tmp/$name.ncu-rep > $O/${name}_ncu_full.txt 2>&1
  python tools/ncu_lines.py /tmp/$name.ncu-rep > $O/${name}_ncu_lines.txt 2>&1
}
cap lb sqk_dtw_lb_kernel 3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras
cap win sqk_dtw_kernel 6 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras
cap stats3_zscale sqk_stats3_kernel 3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras
cap stats3_medmad sqk_stats3_kernel 3 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras --scale medmad
cap stats3_seg sqk_stats3_kernel 3 python tools/bench_segmenter.py --reads 1000000 --steps 1 --no-e2e
cap fsm_mask sqk_fsm_mask_kernel 3 python tools/bench_segmenter.py --reads 1000000 --steps 1 --no-e2e
export PYTHONPATH=$PWD
for tool in memcheck racecheck; do
  echo "== $tool" >> $O/compute_sanitizer.log
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 5 \
    python -m pytest tests/test_stats_paths_gpu.py tests/test_rollmean_gpu.py tests/test_long_motif_gpu.py tests/test_twopass_gpu.py tests/test_segmenter_gpu.py -q -x -m gpu \
      -k "outlier_counts or every_length or golden or 1025 or small_window or lane_layouts or unaligned" 2>&1 | tail -6 >> $O/compute_sanitizer.log
  echo "exit: $?" >> $O/compute_sanitizer.log
done
tail -20 $O/compute_sanitizer.log
python - <<'PY'
import json
for f in ("bench_n1", "bench_reference", "bench_single_pass", "bench_medmad", "bench_n163", "bench_motif2000", "bench_config3"):
    try:
        d = json.loads(open(f"gpurun_out/r2final/{f}.json").read().strip().split("\n")[-1])
        r = d.get("roofline") or {}
        print(f, "value", round(d["value"]), "e2e", d.get("e2e") and round(d["e2e"]["value"]), "lb", r.get("kernel_ms_per_launch"), "win", r.get("exact_windows_ms_per_step"), "stats", r.get("stats_kernel_ms_per_launch"), d.get("parity"), d.get("plan"))
    except Exception as e:
        print(f, "unreadable", e)
PY
