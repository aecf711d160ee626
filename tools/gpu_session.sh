#!/bin/bash
# One gpurun call: parity tests, then the bench lines.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/session.log) 2>&1
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,power.limit --format=csv
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "=== pytest -m gpu"
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15
echo "=== smoke"
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()"
echo "=== LDG check"
RMG_FIT_VARIANT=ldg RMG_BENCH_SPIN_S=0 timeout -k 10 600 python bench.py --steps 6 --warmup 3 --parts masked | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ldg cfg2', d['roofline']['step_ms'], 'masked', d['masked_variant']['ms_per_step'])"
echo "=== bench full (defaults)"
timeout -k 10 1500 python bench.py --steps 10 --warmup 3 | tee gpurun_out/bench_full.json
echo "=== reference arm"
timeout -k 10 900 python bench.py --impl reference --steps 2 --warmup 1 | tee gpurun_out/bench_reference.json
echo "=== done"
echo "=== ncu: traffic of the fit kernels after the local-memory fix"
export RMG_BENCH_SPIN_S=0
CMD="python bench.py --steps 2 --warmup 3 --parts masked"
$CMD > gpurun_out/plain_fit.log 2>&1 &&
ncu --set full --clock-control none -k regex:"lm_fit" -s 3 -c 4 -o /tmp/prof_fit_fix -f $CMD > gpurun_out/ncu_fit.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/prof_fit_fix.ncu-rep --page raw --csv > gpurun_out/prof_fit_fix_raw.csv 2>/dev/null
