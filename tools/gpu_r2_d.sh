#!/bin/bash
# Round 2, call D (gpurun --gpus 8): the bench exactly as the driver launches it at N = 8 (torchrun), with the legs that
# matter on 8 GPUs: strong-scaled mincRandomize, the single-process multi-GPU form R calls (rmg_*_multi, resident dataset,
# concurrent host-link probe) and config 5.  Topology of the box is printed first.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_d.log) 2>&1
nvidia-smi --query-gpu=name,memory.total --format=csv | head -3
nvidia-smi topo -m 2>/dev/null | head -14
lscpu | grep -E "^CPU\(s\)|NUMA|Model name|Socket"; grep -E "Cpus_allowed_list|Mems_allowed_list" /proc/self/status
nproc; free -g | head -2
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "=== multi-GPU tests (8 GPUs visible)"
timeout -k 10 600 python -m pytest tests/test_gpu_dataset.py tests/test_gpu_rshim.py -m gpu -q -p no:cacheprovider 2>&1 | tail -6
export RMG_BENCH_SPIN_S=0
echo "=== bench N=8 (torchrun)"
timeout -k 10 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 8 --steps 5 --warmup 3 --parts randomize,e2e,multi,cfg5 > gpurun_out/r2d_bench_n8.json 2> gpurun_out/r2d_bench_n8.err
echo "rc=$?"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/r2d_bench_n8.err | tail -12
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2d_bench_n8.json").read().strip().splitlines()[-1])
    print(json.dumps({k: d.get(k) for k in ("value", "ms_per_step", "randomize", "e2e", "single_process_multi", "config5", "multi_parity")}, indent=1))
except Exception as e:
    print("unreadable:", e)
PY
echo "=== done"
