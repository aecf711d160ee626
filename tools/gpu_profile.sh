#!/bin/bash
# ncu evidence for profiles/ (one gpurun call; every profiled command first exits 0 un-profiled).
# The .ncu-rep files are converted to CSV on the box and deleted (gpurun_out/ is capped at 64 MiB).
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/profile.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
export RMG_BENCH_SPIN_S=0
CMD="python bench.py --steps 3 --warmup 3 --parts masked,vertex"
$CMD > gpurun_out/plain_fit.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_fit.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
# skip the 25 input-generation launches and the warm-up; keep one launch of each kernel of interest
ncu --set full --clock-control none --import-source on -k regex:"lm_fit|mask_" -c 24 -o /tmp/prof_fit_final -f $CMD > gpurun_out/ncu_fit.log 2>&1
echo "ncu fit rc=$?"
ncu -i /tmp/prof_fit_final.ncu-rep --page raw --csv > gpurun_out/prof_fit_final_raw.csv 2>/dev/null
CMD2="python bench.py --steps 1 --warmup 3 --perms 64 --parts randomize"
$CMD2 > gpurun_out/plain_perm.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_perm.csv $CMD2 > gpurun_out/ncu_launches_perm.log 2>&1
echo "launch list perm rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"perm_" -c 8 -o /tmp/prof_perm_final -f $CMD2 > gpurun_out/ncu_perm.log 2>&1
echo "ncu perm rc=$?"
ncu -i /tmp/prof_perm_final.ncu-rep --page raw --csv > gpurun_out/prof_perm_final_raw.csv 2>/dev/null
ls -la gpurun_out | head -30
