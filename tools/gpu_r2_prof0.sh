#!/bin/bash
# Round 2, first gpurun call: source-level ncu evidence (--set full --import-source on, source page exported as CSV) for
# the kernels VERDICT r1 names as below their roof: perm_gemm_kernel, lm_fit_packed_tma_kernel, the masked lm_fit_ldg_kernel
# and the tfce_* level kernels.  Every profiled command first exits 0 un-profiled.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_prof0.log) 2>&1
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,power.limit --format=csv
nvidia-smi topo -m 2>/dev/null | head -20
nproc; free -g | head -2
python -c "import __graft_entry__ as g; g.build()" || exit 1
export RMG_BENCH_SPIN_S=0

prof() {   # name, kernel regex, launches to keep, matching launches to skip, command...
  local name=$1 regex=$2 count=$3 skip=$4; shift 4
  "$@" > gpurun_out/plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"$regex" -s "$skip" -c "$count" -o /tmp/r02_$name -f "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "ncu $name rc=$?"
  ncu -i /tmp/r02_$name.ncu-rep --page raw --csv > gpurun_out/r02a_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/r02_$name.ncu-rep --page source --csv > gpurun_out/r02a_${name}_source.csv 2>/dev/null
  ls -la /tmp/r02_$name.ncu-rep gpurun_out/r02a_${name}_*.csv
}

prof perm "perm_gemm_kernel" 1 1 python bench.py --steps 1 --warmup 3 --perms 128 --parts randomize
prof stored "lm_fit_packed" 1 3 python bench.py --steps 1 --warmup 3 --parts stored
prof masked "lm_fit_ldg" 1 1 python bench.py --steps 1 --warmup 3 --parts masked
prof tfce "tfce_(link|mass|add)" 30 180 python bench.py --steps 1 --warmup 3 --tfce-perms 4 --parts tfce --small
echo "=== tfce timing (host-side breakdown)"
RMG_TIMING=1 timeout -k 10 600 python bench.py --steps 1 --warmup 3 --parts tfce --skip-cpu 2>&1 | tail -5
echo "=== done"
