#!/bin/bash
# Round 2, call Y (1 GPU): L2 prefetch size of the compacted masked fit's loads (ld.global...L2::64B / 128B / 256B / default):
# time per step and DRAM bytes per variant; Dataset.from_minc2 test.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_y.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
export RMG_BENCH_SPIN_S=0
echo "=== pytest -m gpu -k 'mask or minc2 or dataset'"
timeout -k 10 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "mask or minc2 or dataset" 2>&1 | tail -4
run() { timeout -k 10 600 python bench.py --steps 10 --warmup 3 --parts masked 2>gpurun_out/r2y_err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); m=d['masked_variant']; print('$1', round(m['ms_per_step'],4), 'ms masked step;', round(m.get('achieved_gbs',0),1), 'GB/s; sm_mhz', d['clocks']['sm_mhz'])" || tail -5 gpurun_out/r2y_err.txt; }
for h in -1 0 64 128 256; do
  if [ $h = -1 ]; then run default; else RMG_LDG_L2=$h run l2_$h; fi
done
for h in 0 64; do
RMG_LDG_L2=$h ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"lm_fit_ldg" -s 1 -c 1 --csv --log-file gpurun_out/r02y_masked_l2_$h.csv python bench.py --steps 1 --warmup 3 --parts masked > /dev/null 2>&1
grep -o '"dram__bytes_read.sum","[a-zA-Z]*","[0-9.,]*"\|"gpu__time_duration.sum","[a-z]*","[0-9.,]*"' gpurun_out/r02y_masked_l2_$h.csv
done
echo "=== done"
