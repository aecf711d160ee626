#!/bin/bash
# Round 2: the randomize leg of the bench alone at N ranks (gpurun --gpus N; N from the first argument), after the warm-up fix.
set -u
N=${1:-2}
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_rand_n$N.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
export RMG_BENCH_SPIN_S=0
for rep in 1 2; do
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$rep \
  bench.py --gpus $N --steps 5 --warmup 3 --parts randomize,multi,nocfg5 > gpurun_out/r2f_rand_n$N.json 2> gpurun_out/r2f_rand_n$N.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2f_rand_n$N.json").read().strip().splitlines()[-1])
r = d.get("randomize") or {}
m = d.get("single_process_multi") or {}
ds = m.get("resident_dataset") or {}
print("N=$N randomize ms", r.get("ms"), "perms/s", r.get("value"), "| one-process e2e s", (m.get("randomize_e2e") or {}).get("seconds"), "| dataset first/again", ds.get("randomize_seconds_first"), ds.get("randomize_seconds_again"))
PY
done
echo "=== done"
