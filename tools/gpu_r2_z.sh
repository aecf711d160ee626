#!/bin/bash
# Round 2, call Z (1 GPU): the masked stored-type fit on compacted quads -- parity (new test + every stored / mask test),
# time beside the plain kernel (RMG_PACKED_NOCOMPACT=1) and the unmasked fit.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_z.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
export RMG_BENCH_SPIN_S=0
echo "=== pytest -m gpu -k 'stored or mask or minc2'"
timeout -k 10 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "stored or mask or minc2" 2>&1 | tail -6
run() { timeout -k 10 600 python bench.py --steps 10 --warmup 3 --parts stored 2>gpurun_out/r2z_err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stored_u16']; print('$1', round(s['ms_per_step'],4), 'ms stored;', s.get('masked'))" || tail -5 gpurun_out/r2z_err.txt; }
run default
RMG_PACKED_NOCOMPACT=1 run plain_kernel
RMG_STORED_EXACT=1 run exact
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k regex:"quad|packed" -c 12 --csv --log-file gpurun_out/r02z_masked_stored.csv python bench.py --steps 1 --warmup 3 --parts stored > /dev/null 2>&1
grep -o '"[a-z_]*kernel[^"]*".*' gpurun_out/r02z_masked_stored.csv | cut -d, -f1,9- | tail -24
echo "=== done"
