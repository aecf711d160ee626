#!/bin/bash
# ncu --set full of the vertexLm-sized TMA launch (config 3): the lm_fit_tma_kernel launches after the headline's.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_prof_vtx.log) 2>&1
export RMG_BENCH_SPIN_S=0
python bench.py --steps 1 --warmup 3 --parts vertex --skip-cpu > gpurun_out/plain_vertex.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"lm_fit_tma_kernel" -s 5 -c 4 -o /tmp/r02h_vertex -f python bench.py --steps 1 --warmup 3 --parts vertex --skip-cpu > gpurun_out/ncu_vertex.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/r02h_vertex.ncu-rep --page raw --csv > gpurun_out/r02h_prof_vertex_raw.csv 2>/dev/null
python tools/summarize_profiles.py raw gpurun_out/r02h_prof_vertex_raw.csv 2>/dev/null | head -20
echo "=== done"
