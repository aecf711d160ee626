#!/bin/bash
# Round 2, call K (1 GPU): permutation GEMM experiment switches -- early barrier probe (RMG_PERM_VAR bit 1), register
# double-buffered fragments (bit 2), 224-register build (RMG_PERM_WIDE=1); parity of every variant, 512 permutations each.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_k.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
export RMG_BENCH_SPIN_S=0
for cfg in "0 0" "0 1" "0 2" "0 3" "1 0" "1 1" "1 2" "1 3"; do
  set -- $cfg
  export RMG_PERM_WIDE=$1 RMG_PERM_VAR=$2
  echo "=== wide=$1 var=$2"
  timeout -k 10 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -p no:cacheprovider -x -k "randomize" 2>&1 | tail -1
  for rep in 1 2; do
    timeout -k 10 600 python bench.py --steps 1 --warmup 3 --perms 512 --parts randomize 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['randomize']; print('  ms for 512:', round(r['ms'],2), ' clocks', d['clocks']['sm_mhz'])"
  done
done
echo "=== done"
