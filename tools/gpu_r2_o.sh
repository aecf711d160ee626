#!/bin/bash
# Round 2, call O (1 GPU): mbarrier waits with a suspend-time hint (async_ptx.cuh) -- stored-type fit, permutation GEMM,
# headline fit; parity of the ring kernels.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_o.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
export RMG_BENCH_SPIN_S=0
echo "=== parity (ring kernels)"
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -p no:cacheprovider -x 2>&1 | tail -3
run() { timeout -k 10 600 python bench.py --steps 10 --warmup 3 --parts stored 2>gpurun_out/r2o_err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stored_u16']; print('$1', round(s['ms_per_step'],4), 'ms stored;', round(d['ms_per_step'],4), 'ms fit; sm_mhz', d['clocks']['sm_mhz'])" || tail -5 gpurun_out/r2o_err.txt; }
run default
for c in 1 2 9; do RMG_TMA_CFG=$c run cfg$c; done
for rep in 1 2; do
  timeout -k 10 600 python bench.py --steps 1 --warmup 3 --perms 512 --parts randomize 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['randomize']; print('  ms for 512 perms:', round(r['ms'],2), ' clocks', d['clocks']['sm_mhz'])"
done
timeout -k 10 600 python bench.py --steps 1 --warmup 3 --perms 1000 --parts randomize 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['randomize']; print('  ms for 1000 perms:', round(r['ms'],2), ' clocks', d['clocks']['sm_mhz'])"
echo "=== done"
