#!/bin/bash
# Round 2, call L (1 GPU): lazy TFCE sweep -- whole GPU suite, smoke, TFCE bench legs in both formulations.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_l.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "=== pytest -m gpu -k tfce"
timeout -k 10 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "tfce or TFCE" 2>&1 | tail -30
echo "=== pytest -m gpu (all)"
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -8
echo "=== smoke"
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
export RMG_BENCH_SPIN_S=0
for sw in 0 1; do
  echo "=== bench tfce, RMG_TFCE_SWEEP=$sw"
  RMG_TFCE_SWEEP=$sw RMG_TIMING=1 timeout -k 10 900 python bench.py --steps 3 --warmup 3 --parts tfce > gpurun_out/r2l_bench_tfce_$sw.json 2> gpurun_out/r2l_bench_tfce_$sw.err
  echo "rc=$?"; grep "rmg tfce" gpurun_out/r2l_bench_tfce_$sw.err | tail -6
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2l_bench_tfce_$sw.json").read().strip().splitlines()[-1])
print(json.dumps(d.get("vertexTFCE"), indent=1))
PY
done
echo "=== done"
