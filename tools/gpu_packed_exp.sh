#!/bin/bash
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/packed_exp.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
export RMG_BENCH_SPIN_S=0
run() { python bench.py --steps 5 --warmup 3 --parts stored | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])['stored_u16']; print('$1', round(d['ms_per_step'],4), 'ms', d['kernel'])"; }
run default
for c in 9 11; do RMG_TMA_CFG=$c run cfg$c; done
echo "=== done"
