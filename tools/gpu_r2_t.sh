#!/bin/bash
# Round 2, call T (1 GPU): TFCE legs in both formulations (lazy sweep = default, RMG_TFCE_SWEEP=1 = level sweep).
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_t.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
export RMG_BENCH_SPIN_S=0
for sw in 0 1; do
  for rep in 1 2; do
  echo "=== bench tfce, RMG_TFCE_SWEEP=$sw rep $rep"
  RMG_TFCE_SWEEP=$sw RMG_TIMING=1 timeout -k 10 900 python bench.py --steps 3 --warmup 3 --parts tfce --skip-cpu > gpurun_out/r2t_bench_tfce_$sw.json 2> gpurun_out/r2t_bench_tfce_$sw.err
  echo "rc=$?"; grep "rmg tfce" gpurun_out/r2t_bench_tfce_$sw.err | tail -8
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2t_bench_tfce_$sw.json").read().strip().splitlines()[-1])
t = d.get("vertexTFCE"); t.pop("timing", None)
print(json.dumps(t))
PY
  done
done
echo "=== done"
