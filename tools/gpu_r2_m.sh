#!/bin/bash
# Round 2, call M (1 GPU): stored-type fit with the one-FMA expansion (FOLD) -- parity, shapes (library built with
# -DRMG_PACKED_EXPERIMENTS), the three-rounding twin, then the whole GPU suite and smoke on this build.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_m.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
export RMG_BENCH_SPIN_S=0
echo "=== pytest -m gpu -k stored"
timeout -k 10 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "stored" 2>&1 | tail -5
echo "=== the same with RMG_STORED_EXACT=1"
RMG_STORED_EXACT=1 timeout -k 10 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "stored" 2>&1 | tail -3
run() { timeout -k 10 600 python bench.py --steps 10 --warmup 3 --parts stored 2>gpurun_out/r2m_err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stored_u16']; print('$1', round(s['ms_per_step'],4), 'ms', s['kernel'], 'sm_mhz', d['clocks']['sm_mhz'])" || tail -5 gpurun_out/r2m_err.txt; }
run default
RMG_STORED_EXACT=1 run exact
for c in 1 2 3 4 5 6; do RMG_TMA_CFG=$c run cfg$c; done
run default_again
echo "=== pytest -m gpu (all)"
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -8
echo "=== smoke"
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== done"
