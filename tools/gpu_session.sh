#!/bin/bash
# One gpurun call: parity tests, benches.  (ncu captures live in tools/gpu_profile.sh.)
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/session.log) 2>&1
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,power.limit --format=csv
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "=== pytest -m gpu"
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -30
echo "=== smoke"
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()"
export RMG_BENCH_SPIN_S=0
echo "=== fit kernel variants on config 2 (headline) and config 3 (vertexLm)"
for v in "ldg -1" "tma 4" "tma 5" "tma 0"; do
  set -- $v
  RMG_FIT_VARIANT=$1 RMG_TMA_CFG=$2 timeout -k 10 600 python bench.py --steps 5 --warmup 3 --parts vertex | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', 'cfg2 ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), d['roofline']['kernel'], '| cfg3 ms', round(d['vertexLm']['ms_per_step'],3), 'frac', round(d['vertexLm']['frac_of_hbm_peak'],3))"
done
unset RMG_BENCH_SPIN_S
echo "=== bench full (defaults)"
timeout -k 10 1500 python bench.py --steps 10 --warmup 3 | tee gpurun_out/bench_full.json
echo "=== done"
