#!/bin/bash
for L in 16 32; do
  python bench.py --motif-len 163 --lanes $L --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('N=163 lanes', d['config']['dtw_lanes_per_read'], 'value', round(d['value']), 'dtw ms', round(d['roofline']['kernel_ms_per_launch'],2), d['parity']['indices_bit_exact'])"
done
for N in 40 120 240 400; do
  for L in 0; do
  python bench.py --motif-len $N --lanes $L --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --reads 50000 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('N', d['config']['n_motif'], 'lanes', d['config']['dtw_lanes_per_read'], 'cells/s', '%.3e' % d['roofline_alu']['achieved_cells_per_s'], d['parity']['indices_bit_exact'])"
  done
done
