#!/bin/bash
# Weak-scaling bench line at N GPUs of one box (the driver's own launch line).  usage: tools/scale_run.sh N
N=${1:-2}
mkdir -p gpurun_out/scale
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 \
    > gpurun_out/scale/bench_n$N.json 2> gpurun_out/scale/bench_n$N.err; echo "rc=$?"
tail -2 gpurun_out/scale/bench_n$N.err
python - "$N" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/scale/bench_n{sys.argv[1]}.json").read().strip().split("\n")[-1])
print("n_gpus", d["n_gpus"], "value", round(d["value"]), "per GPU", round(d["value"] / d["n_gpus"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 3), d["parity"], d["clocks"])
PY
