#!/bin/bash
# The BASELINE configs that are not the driver's default bench line, under the default (two-pass) plan:
#   e2e check after the chunk ramp-down; N=163 (the reference's example model) at 100k x 4096;
#   configs[3]: 1M reads x 20000 samples (40 GB resident in HBM); 200k x 50000 (read shape of configs[4]).
O=gpurun_out/extra; mkdir -p $O
summ() { python - "$1" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
print(d["config"]["n_motif"], "x", d["config"]["n_samples"], "reads/GPU", d["config"]["reads_per_gpu"], "plan", d["plan"],
      "| value", round(d["value"]), "reads/s | e2e", d["e2e"] and round(d["e2e"]["value"]), "|", d["roofline"]["kernel"], "ms", round(d["roofline"]["kernel_ms_per_launch"], 2),
      "win ms", d["roofline"]["exact_windows_ms_per_step"], "stats ms", round(d["roofline"]["stats_kernel_ms_per_launch"], 2),
      "| hbm frac", round(d["roofline"]["frac"], 4), "| cells/s", f'{d["roofline_alu"]["achieved_cells_per_s"]:.3e}', "|", d["parity"])
PY
}
timeout 300 python -m pytest tests/test_motifseq_gpu.py -q -k "chunk or pipeline" 2>&1 | tail -2
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $O/bench_default.json 2>/dev/null && summ $O/bench_default.json
timeout 300 python bench.py --motif-len 163 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_n163.json 2>/dev/null && summ $O/bench_n163.json
timeout 600 python bench.py --reads 1000000 --samples 20000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_config3_1Mx20k.json 2>$O/bench_config3.err && summ $O/bench_config3_1Mx20k.json || tail -3 $O/bench_config3.err
timeout 600 python bench.py --reads 200000 --samples 50000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_200kx50k.json 2>$O/bench_50k.err && summ $O/bench_200kx50k.json || tail -3 $O/bench_50k.err
