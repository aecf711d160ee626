#!/bin/bash
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/vertex_exp.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
export RMG_BENCH_SPIN_S=0
run() { python bench.py --small --steps 20 --warmup 3 --parts vertex | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])['vertexLm']; print('$1', round(d['ms_per_step'],4), 'ms', round(d['achieved_gbs']), 'GB/s', d['kernel'])"; }
run default
RMG_TMA_CFG=4 run cfg4_wide
RMG_TMA_CFG=5 run cfg5_narrow
RMG_TMA_CFG=0 run cfg0
RMG_FIT_VARIANT=ldg run ldg_auto
RMG_FIT_VARIANT=ldg RMG_FIT_SPLIT=4 run ldg_s4
RMG_FIT_VARIANT=ldg RMG_FIT_SPLIT=8 run ldg_s8
RMG_FIT_VARIANT=ldg RMG_FIT_SPLIT=16 run ldg_s16
echo "=== done"
