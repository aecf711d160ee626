#!/bin/bash
# One gpurun call: parity tests of the stored-type and covariate paths + their bench legs + ncu of the new kernels.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/stored.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "=== pytest stored / covariate"
timeout -k 10 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "stored or covariate" 2>&1 | tail -40
echo "=== bench legs"
export RMG_BENCH_SPIN_S=0
timeout -k 10 900 python bench.py --steps 5 --warmup 3 --parts cov,stored,e2e_u16 | tee gpurun_out/bench_stored.json
echo "=== MINB+1"
RMG_TMA_CFG=1 timeout -k 10 900 python bench.py --steps 5 --warmup 3 --parts stored | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('minb+1', d['stored_u16']['ms_per_step'], d['stored_u16']['kernel'])"
echo "=== ncu"
CMD="python bench.py --steps 1 --warmup 3 --parts cov,stored"
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"lm_fit_packed|lm_cov" -c 4 -o /tmp/prof_new -f $CMD > gpurun_out/ncu_new.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/prof_new.ncu-rep --page raw --csv > gpurun_out/prof_new_raw.csv 2>/dev/null
ncu -i /tmp/prof_new.ncu-rep --page details --csv > gpurun_out/prof_new_details.csv 2>/dev/null
ls -la gpurun_out | tail -5
echo "=== done"
