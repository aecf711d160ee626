#!/bin/bash
# Kernel-variant experiments (one gpurun call).  Prints one compact line per configuration.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/experiment.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
run() {  # label, then env assignments
  local label="$1"; shift
  env "$@" timeout -k 10 600 python bench.py --steps 6 --warmup 3 --parts vertex 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; v=d['vertexLm']
print('$label', '| cfg2', r['kernel'][7:10], 'ms', round(d['ms_per_step'],3), '| cfg3', v['kernel'][7:10], 'ms', round(v['ms_per_step'],3), 'frac', round(v['frac_of_hbm_peak'],3), 'host_us', round(v['host_call_us']))"
}
echo "=== pytest -m gpu"
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -5
echo "=== cfg3 device time per variant"
run "auto             " RMG_BENCH_SPIN_S=0




for s in 0 4; do
  run "ldg split=$s       " RMG_FIT_VARIANT=ldg RMG_FIT_SPLIT=$s RMG_BENCH_SPIN_S=0
done
echo "=== bench full"
unset RMG_BENCH_SPIN_S
timeout -k 10 1500 python bench.py --steps 10 --warmup 3 | tee gpurun_out/bench_full.json
echo "=== done"
