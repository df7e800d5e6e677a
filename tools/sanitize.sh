#!/bin/bash
# compute-sanitizer memcheck + racecheck over a small slice of the GPU tests (run under gpurun): both DTW plans
# (lower-bound scan, window jobs, finalize, fallback), the stats / FSM kernels, the adapter finder.
set -o pipefail
export PYTHONPATH=$PWD
for tool in memcheck racecheck; do
  echo "== $tool"
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 5 \
    python -m pytest tests/test_motifseq_gpu.py tests/test_segmenter_gpu.py tests/test_twopass_gpu.py tests/test_adapter_gpu.py -q -x -m gpu \
      -k "ragged or unaligned or golden or pa_mode_golden or example or small_window or lane_layouts or tie_heavy" 2>&1 | tail -15
  echo "exit: $?"
done
