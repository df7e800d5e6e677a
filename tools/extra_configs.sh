#!/bin/bash
# Dev helper: the BASELINE configs that are not the driver's default bench line.
#   N=163 (the reference's example model expands to 163 points) at 100k x 4096, three lane layouts
#   configs[3]: 1M reads x 20000 samples (40 GB of int16 resident in HBM), N=80
summ() { python - "$1" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
print(d["config"]["n_motif"], "x", d["config"]["n_samples"], "reads/GPU", d["config"]["reads_per_gpu"], "lanes", d["config"]["dtw_lanes_per_read"],
      "| value", round(d["value"]), "reads/s | dtw ms", round(d["roofline"]["kernel_ms_per_launch"], 2),
      "| hbm frac", round(d["roofline"]["frac"], 4), "| cells/s", f'{d["roofline_alu"]["achieved_cells_per_s"]:.3e}', "|", d["parity"])
PY
}
for L in 0 8 32; do
  python bench.py --motif-len 163 --lanes $L --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_n163_l$L.json 2>/dev/null && summ gpurun_out/bench_n163_l$L.json
done
python bench.py --reads 1000000 --samples 20000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_config3_1Mx20k.json 2>gpurun_out/bench_config3.err && summ gpurun_out/bench_config3_1Mx20k.json || tail -3 gpurun_out/bench_config3.err
