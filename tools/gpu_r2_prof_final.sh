#!/bin/bash
# Round 2, final ncu --set full captures of the kernels changed late in the round: the compacted masked fit with the 64-byte L2
# prefetch size, the vertexLm-sized TMA configuration, the quad-compacted masked stored-type fit.  Each after its plain run exited 0.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_prof_final.log) 2>&1
export RMG_BENCH_SPIN_S=0
prof() {   # name, kernel regex, launches to keep, matching launches to skip, command...
  local name=$1 regex=$2 count=$3 skip=$4; shift 4
  "$@" > gpurun_out/plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"$regex" -s "$skip" -c "$count" -o /tmp/r02h_$name -f "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "ncu $name rc=$?"
  ncu -i /tmp/r02h_$name.ncu-rep --page raw --csv > gpurun_out/r02h_prof_${name}_raw.csv 2>/dev/null
  python tools/summarize_profiles.py raw gpurun_out/r02h_prof_${name}_raw.csv "$regex" 2>/dev/null | head -20
}
prof masked "lm_fit_ldg" 1 1 python bench.py --steps 1 --warmup 3 --parts masked --skip-cpu
prof vertex "lm_fit_tma_kernel<\(int\)7|lm_fit_tma_kernel<7" 1 2 python bench.py --steps 1 --warmup 3 --parts vertex --skip-cpu
prof quads "lm_fit_packed_quads" 1 1 python bench.py --steps 1 --warmup 3 --parts stored --skip-cpu
echo "=== done"
