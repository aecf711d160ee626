#!/bin/bash
# Round 2 (1 GPU): the stored-type fit on data with tissue / background transitions inside a thread's 4 voxels: the tile / warp
# vote for per-voxel shifts against the unguarded one-shift-per-thread form (26 ms) and the round-2a form (RMG_TMA_CFG=4).
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_edges.log) 2>&1
export RMG_BENCH_SPIN_S=0
echo "=== pytest -m gpu -k 'stored or mask or minc2'"
timeout -k 10 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "stored or mask or minc2 or edges" 2>&1 | tail -4
run() { timeout -k 10 600 python bench.py --steps 10 --warmup 3 --parts stored --skip-cpu 2>gpurun_out/r2e_err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stored_u16']; print('$1', round(s['ms_per_step'],4), 'ms stored;', 'masked', round(s['masked']['ms_per_step'],4))" || tail -5 gpurun_out/r2e_err.txt; }
run smooth
export RMG_BENCH_FLAT_RANGES=1
run flat_smooth
RMG_BENCH_EDGES=64 run flat_edges_every_64
RMG_BENCH_EDGES=16 run flat_edges_every_16
RMG_BENCH_EDGES=16 RMG_TMA_CFG=4 run flat_edges_every_16_round2a_form
RMG_BENCH_BACKGROUND=1 run flat_constant_background
RMG_BENCH_BACKGROUND=1 RMG_TMA_CFG=4 run flat_constant_background_round2a_form
echo "=== done"
