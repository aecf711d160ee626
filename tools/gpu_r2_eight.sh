#!/bin/bash
# Round 2, final 8-GPU record (gpurun --gpus 8): the driver's N = 8 bench command (config 5 on by default at N = 8).
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_eight.log) 2>&1
nvidia-smi --query-gpu=name --format=csv | sort | uniq -c; nproc
python -c "import __graft_entry__ as g; g.build()" || exit 1
export RMG_BENCH_SPIN_S=0
( time timeout -k 10 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2f_bench_n8.json 2> gpurun_out/r2f_bench_n8.err ) 2>&1 | grep real
tail -2 gpurun_out/r2f_bench_n8.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2f_bench_n8.json").read().strip().splitlines()[-1])
print(json.dumps({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "multi_parity")}))
r = d.get("randomize") or {}
print("randomize", {k: r.get(k) for k in ("ms", "value", "frac_of_dgemm", "frac_of_dgemm_executed")}, (r.get("e2e") or {}).get("seconds"), (r.get("e2e_stored_u16") or {}).get("seconds"))
m = d.get("single_process_multi") or {}
print("multi fit", (m.get("fit_e2e") or {}), "\nrand", (m.get("randomize_e2e") or {}).get("seconds"), "\nds", m.get("resident_dataset"), "\nprobe", (m.get("h2d_probe") or {}))
print("e2e", (d.get("e2e") or {}).get("value")); print("cfg5", d.get("config5"))
PY
echo "=== done"
