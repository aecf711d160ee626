#!/bin/bash
# Round 2, call S (1 GPU): stored-type fit without epilogue spills (exact recomputation after the regular stores, one
# slice division per tile) -- parity, timing, DRAM bytes; the 4-CTA/SM shapes of the developer build.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_s.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
export RMG_BENCH_SPIN_S=0
echo "=== pytest -m gpu -k stored"
timeout -k 10 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "stored" 2>&1 | tail -3
run() { timeout -k 10 600 python bench.py --steps 10 --warmup 3 --parts stored 2>gpurun_out/r2s_err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stored_u16']; print('$1', round(s['ms_per_step'],4), 'ms stored; sm_mhz', d['clocks']['sm_mhz'])" || tail -5 gpurun_out/r2s_err.txt; }
run default
RMG_TMA_CFG=1 run nj8
RMG_TMA_CFG=2 run minb4_nj16x2
RMG_TMA_CFG=3 run minb4_nj8x4
RMG_STORED_EXACT=1 run exact
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"lm_fit_packed" -s 3 -c 1 --csv --log-file gpurun_out/r02p_stored_traffic.csv python bench.py --steps 1 --warmup 3 --parts stored > /dev/null 2>&1
cat gpurun_out/r02p_stored_traffic.csv | tail -6
echo "=== done"
