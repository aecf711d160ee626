#!/bin/bash
# gpurun call: the randomize / stored-type tests, then the bench legs that time mincRandomize end to end from host memory.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/perm_e2e.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "=== pytest (randomize, stored)"
timeout -k 10 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "randomize or stored" 2>&1 | tail -15
echo "=== bench: randomize + e2e legs"
export RMG_BENCH_SPIN_S=0
timeout -k 10 1200 python bench.py --steps 3 --warmup 3 --parts stored,randomize,e2e,e2e_u16 > gpurun_out/bench_perm_e2e.json 2> gpurun_out/bench_perm_e2e.err
echo "rc=$?"; tail -5 gpurun_out/bench_perm_e2e.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_perm_e2e.json").read().strip().splitlines()[-1])
print(json.dumps(d["randomize"], indent=1))
print(json.dumps({k: v for k, v in d["e2e"].items() if k in ("value", "pinned_h2d_gbs_probe")}, indent=1))
PY
echo "=== done"
