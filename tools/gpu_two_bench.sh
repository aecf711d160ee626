#!/bin/bash
# gpurun --gpus 2 call: the 2-rank bench exactly as the driver launches it (default parts).
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/two_bench.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
export RMG_BENCH_SPIN_S=0
timeout -k 10 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2_new.json 2> gpurun_out/bench_n2_new.err
echo "rc=$?"; tail -3 gpurun_out/bench_n2_new.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n2_new.json").read().strip().splitlines()[-1])
print(json.dumps({k: d.get(k) for k in ("value", "n_gpus", "ms_per_step", "randomize")}, indent=1))
print(json.dumps({k: v for k, v in d["e2e"].items() if not isinstance(v, dict)}, indent=1))
PY
echo "=== reference arm under torchrun"
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>/dev/null | tail -2 | cut -c1-400
echo "=== done"
