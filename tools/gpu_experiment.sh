#!/bin/bash
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/experiment.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
export RMG_BENCH_SPIN_S=0
echo "=== perm tests"
timeout -k 10 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "randomize or Randomize" 2>&1 | tail -4
echo "=== GEMM decomposition, R=512"
for c in 1 2; do for v in 0 1 2; do
  RMG_PERM_CTAS=$c RMG_PERM_DEBUG_SKIP=$v RMG_PERM_DELAY_PCT=0 timeout -k 10 600 python bench.py --steps 3 --warmup 3 --perms 512 --parts randomize | python -c "
import json,sys
d=json.loads(sys.stdin.read())['randomize']; print('ctas=$c skip=$v', round(d['value'],1), 'perms/s', round(d['ms'],1), 'ms')"
done; done
echo "=== done"
