#!/bin/bash
# Round 2, call Q (1 GPU): stored-type fit -- staggered start of the CTAs that share an SM (RMG_STAGGER_NS).
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_q.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
export RMG_BENCH_SPIN_S=0
run() { timeout -k 10 600 python bench.py --steps 10 --warmup 3 --parts stored 2>gpurun_out/r2q_err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stored_u16']; print('$1', round(s['ms_per_step'],4), 'ms stored; sm_mhz', d['clocks']['sm_mhz'])" || tail -5 gpurun_out/r2q_err.txt; }
for c in 2 0 9; do
  for sg in 0 8000 16000 24000 32000; do RMG_TMA_CFG=$c RMG_STAGGER_NS=$sg run "cfg$c stagger=$sg"; done
done
echo "=== done"
