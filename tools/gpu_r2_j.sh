#!/bin/bash
# Round 2, call J (1 GPU): GPU suite; bench legs stored / tfce (with host-side laps) / masked / config 5 at N = 1; ncu raw +
# source pages of the permutation GEMM as committed (epilogue constants in shared memory).
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_j.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "=== pytest -m gpu"
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -8
export RMG_BENCH_SPIN_S=0
echo "=== bench N=1: stored, masked, tfce (RMG_TIMING)"
RMG_TIMING=1 timeout -k 10 1200 python bench.py --steps 5 --warmup 3 --parts stored,masked,tfce > gpurun_out/r2j_bench_a.json 2> gpurun_out/r2j_bench_a.err
echo "rc=$?"; grep "rmg tfce" gpurun_out/r2j_bench_a.err | tail -12
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2j_bench_a.json").read().strip().splitlines()[-1])
for k in ("stored_u16", "masked_variant", "vertexTFCE"):
    print(k, json.dumps(d.get(k), indent=1))
PY
echo "=== bench N=1: config 5"
timeout -k 10 1200 python bench.py --steps 5 --warmup 3 --parts cfg5 > gpurun_out/r2j_bench_cfg5.json 2> gpurun_out/r2j_bench_cfg5.err
echo "rc=$?"; tail -3 gpurun_out/r2j_bench_cfg5.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2j_bench_cfg5.json").read().strip().splitlines()[-1])
print(json.dumps(d.get("config5"), indent=1))
PY
echo "=== ncu perm (committed kernel)"
CMD="python bench.py --steps 1 --warmup 3 --perms 256 --parts randomize"
$CMD > gpurun_out/plain_perm.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"perm_gemm_kernel" -s 1 -c 1 -o /tmp/r2j_perm -f $CMD > gpurun_out/ncu_perm.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/r2j_perm.ncu-rep --page raw --csv > gpurun_out/r2j_perm_raw.csv 2>/dev/null
ncu -i /tmp/r2j_perm.ncu-rep --page source --csv > gpurun_out/r2j_perm_source.csv 2>/dev/null
echo "=== done"
