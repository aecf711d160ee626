#!/bin/bash
# Multi-GPU validation (gpurun --gpus 2): NCCL path of bench.py + single-process multi-GPU C ABI.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/multi.log) 2>&1
nvidia-smi --query-gpu=index,name --format=csv
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "=== single-process multi-GPU C ABI (rmg_lm_fit_multi / rmg_randomize_multi) vs 1 GPU"
timeout -k 10 600 python - <<'PY'
import numpy as np, sys
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
from helpers import cfg1_data
from rminc_b200 import core
print('devices', core.device_count())
Y, X, mask = cfg1_data(dims=(40, 50, 33))
a = core.lm_fit(Y, X, mask=mask)
b = core.lm_fit(Y, X, mask=mask, n_gpus=2)
assert np.array_equal(a, b, equal_nan=True), 'multi-GPU rows differ'
rng = np.random.default_rng(1)
idx = np.column_stack([rng.permutation(20) for _ in range(50)])
d1 = core.randomize(Y, X, idx, [4, 5], mask=mask)
d2 = core.randomize(Y, X, idx, [4, 5], mask=mask, n_gpus=2)
np.testing.assert_allclose(d2, d1, rtol=1e-12)
print('single-process 2-GPU paths match 1 GPU')
PY
echo "=== torchrun N=2 bench (NCCL all-reduce(MAX) for randomize; weak-scaled mincLm)"
timeout -k 10 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 --parts randomize,e2e | tee gpurun_out/bench_n2.json
echo "=== N=1 for comparison"
timeout -k 10 1500 python bench.py --gpus 1 --steps 10 --warmup 3 --parts randomize,e2e | tee gpurun_out/bench_n1.json
echo "=== done"
