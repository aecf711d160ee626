#!/bin/bash
# Multi-GPU validation (gpurun --gpus N): NCCL path of bench.py at N ranks, then N=1 on the same box.
set -u
N=${1:-2}
mkdir -p gpurun_out
exec > >(tee gpurun_out/multi_$N.log) 2>&1
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8; free -g | head -2; nproc
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "=== torchrun N=$N"
timeout -k 10 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 | tee gpurun_out/bench_n$N.json
echo "=== N=1 on the same box"
timeout -k 10 1500 python bench.py --gpus 1 --steps 10 --warmup 3 --parts randomize,e2e | tee gpurun_out/bench_n1_same_box.json
echo "=== reference arm under torchrun (rank 0 only prints)"
RMG_BENCH_CPU_SAMPLE=200000 timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 1 --warmup 0 | tail -2 | cut -c1-300
echo "=== done"
