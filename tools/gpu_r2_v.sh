#!/bin/bash
# Round 2, call V (1 GPU): lazy TFCE with de-contended settle / flow, automatic choice between the two sweeps.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_v.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "=== pytest -m gpu -k tfce"
timeout -k 10 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "tfce or TFCE" 2>&1 | tail -3
for sw in 0 1 auto; do
  echo "=== volume probe, RMG_TFCE_SWEEP=$sw"
  if [ $sw = auto ]; then unset RMG_TFCE_SWEEP; else export RMG_TFCE_SWEEP=$sw; fi
  RMG_TIMING=1 python tools/vol_tfce_probe.py 3 2>&1 | grep -v "^\[rmg tfce\] [a-z]* *[a-z,]* [a-z]" | tail -8
done
export RMG_BENCH_SPIN_S=0
for sw in 0 1 auto; do
  if [ $sw = auto ]; then unset RMG_TFCE_SWEEP; else export RMG_TFCE_SWEEP=$sw; fi
  echo "=== bench tfce, RMG_TFCE_SWEEP=$sw"
  RMG_TIMING=1 timeout -k 10 900 python bench.py --steps 3 --warmup 3 --parts tfce --skip-cpu > gpurun_out/r2v_bench_tfce_$sw.json 2> gpurun_out/r2v_bench_tfce_$sw.err
  grep "visits" gpurun_out/r2v_bench_tfce_$sw.err | sort | uniq -c | head -4
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2v_bench_tfce_$sw.json").read().strip().splitlines()[-1])
t = d["vertexTFCE"]
print("pageable", [round(x, 4) for x in t["seconds_each_call"]], "pinned", [round(x, 4) for x in t["from_pinned"]["seconds_each_call"]], "mincTFCE", [round(x, 4) for x in t["mincTFCE"]["seconds_each_call"]])
PY
done
echo "=== done"
