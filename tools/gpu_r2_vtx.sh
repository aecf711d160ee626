#!/bin/bash
# Round 2 (1 GPU): vertexLm config 3 -- ring depth of the wide-tile TMA kernel when every tile is resident at once
# (cfg 6: 8 subjects x 4 stages x 3 CTAs/SM; cfg 7: 8 x 6 x 2; cfg 8: 4 x 8 x 3) against the shipped cfg 4 (4 x 4 x 4).
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_vtx.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
export RMG_BENCH_SPIN_S=0
echo "=== pytest -m gpu -k 'config3 or vertex or tma or padded'"
timeout -k 10 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "config3 or vertex or tma or padded or fullsize" 2>&1 | tail -4
run() { timeout -k 10 600 python bench.py --steps 20 --warmup 5 --parts vertex 2>gpurun_out/r2vtx_err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); v=d['vertexLm']; print('$1', round(v['ms_per_step'],4), 'ms vertexLm;', round(v['frac_of_hbm_peak'],3), '; headline', round(d['ms_per_step'],4))" || tail -5 gpurun_out/r2vtx_err.txt; }
run default
for c in 4 6 7 8 5; do RMG_TMA_CFG=$c run cfg$c; done
run default_again
echo "=== done"
