#!/bin/bash
# A/B of the lower-bound kernel: the tree at 43eb63b (before the two-ended job list) vs the working tree, same box.
set -u
O=gpurun_out/r2ab; mkdir -p $O
for i in 1 2; do
  (cd _ab_old && timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > ../$O/old_$i.json 2> ../$O/old_$i.err); echo "old $i rc=$?"
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $O/new_$i.json 2> $O/new_$i.err; echo "new $i rc=$?"
done
SQK_LB_LANES=4 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $O/new_l4.json 2> $O/new_l4.err; echo "new l4 rc=$?"
SQK_LB_LANES=16 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $O/new_l16.json 2> $O/new_l16.err; echo "new l16 rc=$?"
python - <<'PY'
import json
for f in ("old_1", "new_1", "old_2", "new_2", "new_l4", "new_l16"):
    try:
        d = json.loads(open(f"gpurun_out/r2ab/{f}.json").read().strip().split("\n")[-1])
        print(f, "value", round(d["value"]), "lb", round(d["roofline"]["kernel_ms_per_launch"], 3), "win", round(d["roofline"]["exact_windows_ms_per_step"], 3), "stats", round(d["roofline"]["stats_kernel_ms_per_launch"], 3), d["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "unreadable", e)
PY
nvidia-smi --query-gpu=name,serial,uuid,clocks.sm,clocks.mem,power.limit --format=csv
