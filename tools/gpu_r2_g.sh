#!/bin/bash
# Round 2, call G (1 GPU): permutation GEMM with statistics warps that stage their constants -- randomize parity tests,
# bench leg, the MMA-side ceiling (statistics skipped), ncu raw + source pages.
set -u
mkdir -p gpurun_out
TAG=${1:-r2g}
exec > >(tee gpurun_out/${TAG}.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "=== pytest -m gpu -k randomize"
timeout -k 10 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -x -k "randomize or perm or dataset or smoke" 2>&1 | tail -8
export RMG_BENCH_SPIN_S=0
echo "=== bench N=1: randomize"
timeout -k 10 1200 python bench.py --steps 5 --warmup 3 --parts randomize > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
echo "rc=$?"; tail -5 gpurun_out/${TAG}_bench_n1.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_n1.json").read().strip().splitlines()[-1])
r = d.get("randomize")
print({k: r[k] for k in ("ms", "value", "frac_of_dgemm")})
PY
echo "=== ceiling: statistics skipped (results void)"
RMG_PERM_DEBUG_SKIP=1 timeout -k 10 600 python bench.py --steps 1 --warmup 3 --perms 512 --parts randomize 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('skip:', d['randomize']['ms'], 'ms for 512')"
timeout -k 10 600 python bench.py --steps 1 --warmup 3 --perms 512 --parts randomize 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('full:', d['randomize']['ms'], 'ms for 512')"
echo "=== ncu perm"
CMD="python bench.py --steps 1 --warmup 3 --perms 256 --parts randomize"
$CMD > gpurun_out/plain_perm.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"perm_gemm_kernel" -s 1 -c 1 -o /tmp/${TAG}_perm -f $CMD > gpurun_out/ncu_perm.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/${TAG}_perm.ncu-rep --page raw --csv > gpurun_out/${TAG}_perm_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_perm.ncu-rep --page source --csv > gpurun_out/${TAG}_perm_source.csv 2>/dev/null
echo "=== done"
