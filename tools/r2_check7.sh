#!/bin/bash
# Round 2, seventh GPU check (1 GPU): stats3 with the cold paths out of line, mmap text reader; full default bench line.
set -u
O=gpurun_out/r2g; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
python - <<'PY'
import json
for f in ("bench",):
    try:
        d = json.loads(open(f"gpurun_out/r2g/{f}.json").read().strip().split("\n")[-1])
        print(f, "value", round(d["value"]), "e2e", d["e2e"] and round(d["e2e"]["value"]), "lb", round(d["roofline"]["kernel_ms_per_launch"], 3),
              "win", d["roofline"].get("exact_windows_ms_per_step"), "stats", round(d["roofline"]["stats_kernel_ms_per_launch"], 3), d["parity"], "launches", d["gpu_launches"], d["plan"])
        sg = d.get("segmenter") or {}
        for r in sg.get("runs", []):
            print("  seg", r["reads"], round(r["value"]), r["kernels_ms"], round(r["roofline"]["frac_step"], 4), "e2e", round(r["e2e"]["value"]), r["parity_subsample_bit_exact"])
        c = d.get("cli_e2e") or {}
        print("  cli", {k: (v if not isinstance(v, dict) else {kk: vv for kk, vv in v.items() if kk != "sample"}) for k, v in c.items()})
        print("  sustained", d["sustained"] and round(d["sustained"]["value"]), "alu", d["roofline_alu"]["frac"], "pageable", d["e2e"]["pageable"])
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sqk_stats3_kernel -s 3 -c 1 -f -o $O/stats3_zscale \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > /dev/null 2>&1; echo "ncu stats3 zscale rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sqk_stats3_kernel -s 3 -c 1 -f -o $O/stats3_seg \
    python tools/bench_segmenter.py --reads 1000000 --steps 1 > /dev/null 2>&1; echo "ncu stats3 seg rc=$?"
