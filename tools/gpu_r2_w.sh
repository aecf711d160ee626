#!/bin/bash
# Round 2, call W (1 GPU): stored-type fit with byte-aligned sample placement (w = 1 + u 2^-20: one logic op per sample) and
# ONE shift per thread folded into the offsets (25 FP64 instructions per 4 samples instead of 28) -- parity, timing against
# the round-2a form (RMG_TMA_CFG=4), the 4-CTA/SM shapes, FP64-pipe share.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_w.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
export RMG_BENCH_SPIN_S=0
echo "=== pytest -m gpu -k stored"
timeout -k 10 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "stored or packed or u16" 2>&1 | tail -5
run() { timeout -k 10 600 python bench.py --steps 10 --warmup 3 --parts stored 2>gpurun_out/r2w_err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stored_u16']; print('$1', round(s['ms_per_step'],4), 'ms stored; sm_mhz', d['clocks']['sm_mhz'])" || tail -5 gpurun_out/r2w_err.txt; }
run default
RMG_TMA_CFG=4 run round2a_form
RMG_TMA_CFG=1 run nj8
RMG_TMA_CFG=2 run minb4_nj16x2
RMG_TMA_CFG=3 run minb4_nj8x4
RMG_STORED_EXACT=1 run exact
run default_again
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k regex:"lm_fit_packed" -s 3 -c 1 --csv --log-file gpurun_out/r02w_stored_metrics.csv python bench.py --steps 1 --warmup 3 --parts stored > /dev/null 2>&1
tail -8 gpurun_out/r02w_stored_metrics.csv
echo "=== done"
