#!/bin/bash
# Round 2, call A (gpurun --gpus 2): the whole GPU suite with two devices visible (every n_gpus = 2 path, the resident
# dataset), smoke, then the torchrun bench at N = 2 with the single-process multi-GPU legs, then N = 1 multi legs.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_a.log) 2>&1
nvidia-smi --query-gpu=name,memory.total --format=csv
nvidia-smi topo -m 2>/dev/null | head -12
lscpu | grep -E "^CPU\(s\)|NUMA|Model name|Socket" ; cat /sys/devices/system/node/node*/cpulist 2>/dev/null; grep -E "Cpus_allowed_list|Mems_allowed_list" /proc/self/status
nproc; free -g | head -2
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "=== pytest -m gpu (2 GPUs visible)"
timeout -k 10 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -25
echo "=== smoke"
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
export RMG_BENCH_SPIN_S=0
echo "=== bench N=2 (torchrun): randomize + e2e + multi"
RMG_TIMING=1 timeout -k 10 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --steps 5 --warmup 3 --parts randomize,e2e,multi > gpurun_out/r2a_bench_n2.json 2> gpurun_out/r2a_bench_n2.err
echo "rc=$?"; tail -30 gpurun_out/r2a_bench_n2.err
echo "=== bench N=1: randomize + e2e + multi"
RMG_TIMING=1 timeout -k 10 1200 python bench.py --steps 5 --warmup 3 --parts randomize,e2e,multi,tfce > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err
echo "rc=$?"; tail -30 gpurun_out/r2a_bench_n1.err
python - <<'PY'
import json
for f in ("gpurun_out/r2a_bench_n2.json", "gpurun_out/r2a_bench_n1.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, json.dumps({k: d.get(k) for k in ("value", "randomize", "e2e", "single_process_multi", "vertexTFCE")}, indent=1))
    except Exception as e:
        print(f, "unreadable:", e)
PY
echo "=== done"
