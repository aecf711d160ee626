#!/bin/bash
# Round 2, call P (1 GPU): one-division epilogue (lm_common.cuh) -- whole GPU suite; stored-type shapes; headline fit,
# vertexLm, masked legs.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_p.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
export RMG_BENCH_SPIN_S=0
echo "=== pytest -m gpu (all)"
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -8
run() { timeout -k 10 600 python bench.py --steps 10 --warmup 3 --parts stored,vertex,masked 2>gpurun_out/r2p_err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stored_u16']; print('$1', round(s['ms_per_step'],4), 'ms stored;', round(d['ms_per_step'],4), 'ms fit;', round(d['vertexLm']['ms_per_step'],4), 'ms vertexLm;', round(d['masked_variant']['ms_per_step'],4), 'ms masked; sm_mhz', d['clocks']['sm_mhz'])" || tail -5 gpurun_out/r2p_err.txt; }
run default
for c in 1 2 8 9 11; do RMG_TMA_CFG=$c run cfg$c; done
RMG_STORED_EXACT=1 run exact
echo "=== done"
