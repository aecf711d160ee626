#!/bin/bash
# One gpurun call: parity tests, benches.  (ncu captures live in tools/gpu_profile.sh.)
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/session.log) 2>&1
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,power.limit --format=csv
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "=== pytest -m gpu"
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 2>&1 | tail -40
echo "=== smoke"
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()"
echo "=== bench full"
timeout -k 10 1500 python bench.py --steps 10 --warmup 3 | tee gpurun_out/bench_full.json
echo "=== randomize: GEMM with / without the invariant-row skip, R=256"
export RMG_BENCH_SPIN_S=0
timeout -k 10 600 python bench.py --steps 2 --warmup 3 --perms 256 --parts randomize | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('INV  ', json.dumps(d['randomize']))"
RMG_PERM_NO_INV=1 timeout -k 10 600 python bench.py --steps 2 --warmup 3 --perms 256 --parts randomize | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('NOINV', json.dumps(d['randomize']))"
echo "=== vertexLm split sweep"
for s in 0 1 2 4 8; do
  RMG_FIT_SPLIT=$s timeout -k 10 600 python bench.py --steps 5 --warmup 3 --parts vertex | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('split=$s', json.dumps(d['vertexLm']))"
done
RMG_FIT_VARIANT=tma RMG_TMA_CFG=4 timeout -k 10 600 python bench.py --steps 5 --warmup 3 --parts vertex | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('tma4', json.dumps(d['vertexLm']))"
echo "=== done"
