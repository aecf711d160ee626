#!/bin/bash
# Round 2, call N (1 GPU): stored-type FOLD kernel -- ring depth experiments and a source-level ncu capture.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_n.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
export RMG_BENCH_SPIN_S=0
run() { timeout -k 10 600 python bench.py --steps 10 --warmup 3 --parts stored 2>gpurun_out/r2n_err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stored_u16']; print('$1', round(s['ms_per_step'],4), 'ms', s['kernel'], 'sm_mhz', d['clocks']['sm_mhz'])" || tail -5 gpurun_out/r2n_err.txt; }
run default
for c in 2 7 8 9 10; do RMG_TMA_CFG=$c run cfg$c; done
name=stored_fold
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"lm_fit_packed" -s 3 -c 1 -o /tmp/r02_$name -f python bench.py --steps 1 --warmup 3 --parts stored > gpurun_out/ncu_$name.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/r02_$name.ncu-rep --page raw --csv > gpurun_out/r02m_${name}_raw.csv 2>/dev/null
ncu -i /tmp/r02_$name.ncu-rep --page source --csv > gpurun_out/r02m_${name}_source.csv 2>/dev/null
ls -la /tmp/r02_$name.ncu-rep gpurun_out/r02m_${name}_*.csv
python tools/ncu_source_summary.py gpurun_out/r02m_${name}_source.csv 30 > gpurun_out/r02m_${name}_source_summary.md
echo "=== done"
