#!/bin/bash
# Round 2, final 2-GPU record (gpurun --gpus 2): the whole GPU suite with 2 devices visible (every n_gpus = 2 branch, the
# rmincgpu_minc2_model / Dataset.from_minc2 two-GPU branches), then the driver's N = 2 bench command.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_two.log) 2>&1
nvidia-smi --query-gpu=name,memory.total --format=csv
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "=== pytest -m gpu (2 GPUs visible)"
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -6
export RMG_BENCH_SPIN_S=0
echo "=== bench N=2 (driver's command)"
( time timeout -k 10 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err ) 2>&1 | grep real
tail -2 gpurun_out/r2f_bench_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2f_bench_n2.json").read().strip().splitlines()[-1])
print(json.dumps({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "multi_parity")}))
r = d.get("randomize") or {}
print("randomize", {k: r.get(k) for k in ("ms", "value", "frac_of_dgemm", "frac_of_dgemm_executed")}, (r.get("e2e") or {}).get("seconds"), (r.get("e2e_stored_u16") or {}).get("seconds"))
m = d.get("single_process_multi") or {}
print("multi", (m.get("fit_e2e") or {}).get("value"), (m.get("randomize_e2e") or {}).get("seconds"), m.get("resident_dataset"), (m.get("h2d_probe") or {}).get("aggregate_gbs"))
print("e2e", (d.get("e2e") or {}).get("value"))
PY
echo "=== done"
