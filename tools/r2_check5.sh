#!/bin/bash
# Round 2, fifth GPU check (1 GPU): wide-lane fallback (rank-3 batch), larger patch area, batched text reader in the CLIs,
# full default bench line (segmenter block + cli_e2e).
set -u
O=gpurun_out/r2e; mkdir -p $O
nproc > $O/nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 $O/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-extras --seed-offset 3 > $O/bench_seed3.json 2> $O/bench_seed3.err; echo "bench seed3 rc=$?"
python - <<'PY'
import json
for f in ("bench", "bench_seed3"):
    try:
        d = json.loads(open(f"gpurun_out/r2e/{f}.json").read().strip().split("\n")[-1])
        print(f, "value", round(d["value"]), "e2e", d["e2e"] and round(d["e2e"]["value"]), "lb", round(d["roofline"]["kernel_ms_per_launch"], 3),
              "win", d["roofline"].get("exact_windows_ms_per_step"), "stats", round(d["roofline"]["stats_kernel_ms_per_launch"], 3), d["parity"], "launches", d["gpu_launches"], d["plan"])
        if f == "bench":
            sg = d.get("segmenter") or {}
            for r in sg.get("runs", []):
                print("  seg", r["reads"], round(r["value"]), r["kernels_ms"], round(r["roofline"]["frac_step"], 4), "e2e", round(r["e2e"]["value"]), r["parity_subsample_bit_exact"])
            print("  seg cpu", sg.get("cpu_baseline"), sg.get("unavailable"))
            print("  cli", json.dumps(d.get("cli_e2e"))[:1500])
            print("  sustained", d["sustained"] and round(d["sustained"]["value"]), "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"]))
    except Exception as e:
        print(f, "unreadable", e)
PY
