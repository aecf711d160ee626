#!/bin/bash
# One gpurun call: parity tests first (safe paths, then TMA / GEMM paths), then benches.
# Everything is logged under gpurun_out/ ; each stage has its own hard timeout.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/session.log) 2>&1
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,power.limit --format=csv
free -g | head -2; nproc
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "=== pytest safe paths (ldg fit, refit permutations)"
timeout -k 10 900 python -m pytest tests -m gpu -q -x -k "not tma and not gemm" -p no:cacheprovider 2>&1 | tail -25
echo "=== pytest TMA fit + GEMM permutation paths"
timeout -k 10 900 python -m pytest tests -m gpu -q -k "tma or gemm" -p no:cacheprovider 2>&1 | tail -40
echo "=== smoke"
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()"
echo "=== bench headline per variant"
for v in tma ldg; do
  RMG_FIT_VARIANT=$v timeout -k 10 600 python bench.py --steps 5 --warmup 3 --skip-extras | tee gpurun_out/bench_$v.json
done
echo "=== bench full"
timeout -k 10 1200 python bench.py --steps 10 --warmup 3 | tee gpurun_out/bench_full.json
echo "=== randomize refit path"
RMG_PERM_PATH=refit timeout -k 10 600 python bench.py --steps 2 --warmup 3 --perms 32 --skip-cpu | tee gpurun_out/bench_refit.json
echo "=== done"
