#!/bin/bash
# Round 2, call E (1 GPU): full GPU suite (FDR methods added), smoke, randomize bench leg with the epilogue constants
# staged in shared memory.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_e.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "=== pytest -m gpu"
timeout -k 10 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -25
echo "=== smoke"
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
export RMG_BENCH_SPIN_S=0
echo "=== bench N=1: randomize"
timeout -k 10 1200 python bench.py --steps 5 --warmup 3 --parts randomize > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err
echo "rc=$?"; tail -5 gpurun_out/r2e_bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2e_bench_n1.json").read().strip().splitlines()[-1])
print(json.dumps(d.get("randomize"), indent=1))
PY
echo "=== done"
