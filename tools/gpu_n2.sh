#!/bin/bash
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/n2.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "=== fdr edge test"
timeout -k 10 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "fdr_edge or stored_tma" 2>&1 | tail -5
echo "=== torchrun N=2"
timeout -k 10 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 | tee gpurun_out/bench_n2.json
echo "=== done"
