#!/bin/bash
# Dev helper: e2e throughput of bench.py as a function of the host-mode chunk size.
# (A ramp-down of the last chunk -- 1/2, 1/4, 1/4 -- was measured 3-5 % slower at every size and is not used.)
for c in 16777216 33554432 67108864 134217728; do
  SQK_CHUNK_SAMPLES=$c python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null > /tmp/sweep.json
  python - "$c" <<'PY'
import json, sys
d = json.load(open('/tmp/sweep.json'))
print('chunk_samples', sys.argv[1], 'value', round(d['value']), 'e2e', round(d['e2e']['value']))
PY
done
