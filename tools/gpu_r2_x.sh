#!/bin/bash
# Round 2, call X (1 GPU): MINC 2 reader in the loop -- whole GPU suite (new: mincLm on .mnc files, rmincgpu_minc2_model against
# the reference's minc2_model), the from-files bench leg, the stored-type leg.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_x.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
export RMG_BENCH_SPIN_S=0
echo "=== pytest -m gpu (all)"
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -8
echo "=== bench files,stored"
timeout -k 10 900 python bench.py --steps 10 --warmup 3 --parts files,stored > gpurun_out/r2x_bench_files.json 2>gpurun_out/r2x_err.txt || tail -5 gpurun_out/r2x_err.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2x_bench_files.json').read().strip().splitlines()[-1])
print(json.dumps(d['from_files'], indent=1)); print('stored', d['stored_u16']['ms_per_step'], 'fit', d['ms_per_step'])
PY
echo "=== ncu --set full of the stored-type kernel (after its plain run above exited 0)"
ncu --set full --clock-control none --import-source on -k regex:"lm_fit_packed" -s 3 -c 1 -o /tmp/r02x_stored -f python bench.py --steps 1 --warmup 3 --parts stored > gpurun_out/ncu_r02x_stored.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/r02x_stored.ncu-rep --page raw --csv > gpurun_out/r02x_prof_stored_raw.csv 2>/dev/null
ncu -i /tmp/r02x_stored.ncu-rep --page source --csv > gpurun_out/r02x_stored_source.csv 2>/dev/null
python tools/ncu_source_summary.py gpurun_out/r02x_stored_source.csv 16 > gpurun_out/r02x_stored_source_summary.md 2>/dev/null
head -40 gpurun_out/r02x_stored_source_summary.md
echo "=== done"
