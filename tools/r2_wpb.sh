#!/bin/bash
# Block-shape experiment for the fused iteration kernel: warps per block (DILQR_WPB) at the headline shape.
O=gpurun_out; P=${1:-r2wpb}
for w in 4 2 1 3; do
  DILQR_WPB=$w python bench.py --steps 10 --warmup 3 --no-cpu-baseline --single-dtype --no-ncu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('WPB=$w', d['ms_per_step'], d['kernel_ms']['dilqr_mpc_iterate'], d['kernel_ms']['dilqr_mpc_commit'], d['kernel_ms']['dilqr_mpc_gains'])" | tee -a $O/${P}.txt
done
