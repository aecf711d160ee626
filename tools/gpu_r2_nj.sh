#!/bin/bash
# Round 2 (1 GPU): stored-type fit, subjects per ring stage: 16 x 3 stages (shipped) against 32 x 2 and 24 x 2.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_nj.log) 2>&1
export RMG_BENCH_SPIN_S=0
run() { timeout -k 10 600 python bench.py --steps 20 --warmup 5 --parts stored 2>gpurun_out/r2nj_err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stored_u16']; print('$1', round(s['ms_per_step'],4), 'ms stored')" || tail -5 gpurun_out/r2nj_err.txt; }
run default
RMG_TMA_CFG=5 run nj32x2
RMG_TMA_CFG=6 run nj24x2
run default_again
echo "=== done"
