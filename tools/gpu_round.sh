#!/bin/bash
# One gpurun call: full GPU test suite, smoke, the default bench line, the reference arm, then ncu evidence
# (launch list + --set full of the dominant kernels) on commands that have just exited 0 un-profiled.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/round.log) 2>&1
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,power.limit --format=csv
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "=== pytest -m gpu"
timeout -k 10 1800 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15
echo "=== smoke"
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()"
echo "=== bench (defaults)"
timeout -k 10 1500 python bench.py | tee gpurun_out/bench_round.json
echo "=== reference arm"
timeout -k 10 900 python bench.py --impl reference --steps 2 --warmup 1 | tee gpurun_out/bench_reference.json
echo "=== ncu"
export RMG_BENCH_SPIN_S=0
CMD="python bench.py --steps 2 --warmup 3 --parts masked,vertex,cov,stored"
$CMD > gpurun_out/plain_fit.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_fit.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"lm_fit|lm_cov|mask_|pack_tables" -c 40 -o /tmp/prof_fit -f $CMD > gpurun_out/ncu_fit.log 2>&1
echo "ncu fit rc=$?"
ncu -i /tmp/prof_fit.ncu-rep --page raw --csv > gpurun_out/prof_fit_r1b_raw.csv 2>/dev/null
ls -la gpurun_out | tail -12
echo "=== done"
