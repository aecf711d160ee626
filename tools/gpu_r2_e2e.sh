#!/bin/bash
# Round 2 (1 GPU): the end-to-end legs alone (rmg_lm_fit / rmg_randomize from pinned host memory, doubles and uint16).
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_e2e.log) 2>&1
export RMG_BENCH_SPIN_S=0
for rep in 1 2; do
timeout -k 10 900 python bench.py --steps 3 --warmup 3 --parts randomize,e2e,e2e_u16 --skip-cpu 2>gpurun_out/r2e2e_err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['randomize']; e=d['e2e']
print('rep $rep: e2e', round(e['value']/1e6,2), 'M voxels/s; u16', round(e['stored_u16']['value']/1e6,2), '; randomize', round(r['ms'],1), 'ms; e2e', round(r['e2e']['seconds'],3), 's; e2e u16', round(r['e2e_stored_u16']['seconds'],3), 's')" || tail -5 gpurun_out/r2e2e_err.txt
done
echo "=== done"
