#!/bin/bash
# Round 2, call I (1 GPU): whole GPU suite (randomize with a voxel-wise covariate, adjacency symmetry check, FDR methods),
# smoke, timing of the covariate permutation test at config-2 size.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_i.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "=== pytest -m gpu"
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -25
echo "=== smoke"
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "=== covariate permutation test, config-2 shape"
timeout -k 10 900 python - <<'PY'
import time, numpy as np, torch
from rminc_b200 import core
V, n, ps, R = 7076160 // 4, 500, 3, 16            # a quarter of config 2 (host buffers: 2 x 7 GB)
rng = np.random.default_rng(0)
Y = np.empty((V, n), order="F"); Z = np.empty((V, n), order="F")
for j in range(n):
    Y[:, j] = rng.standard_normal(V, dtype=np.float32)
    Z[:, j] = 50 + 3 * rng.standard_normal(V, dtype=np.float32)
Xs = np.column_stack([np.ones(n), np.arange(n) % 2, rng.normal(60, 10, n)])
idx = np.column_stack([rng.permutation(n) for _ in range(R)])
cols = [6, 7, 8, 9]
core.randomize_cov(Y[:4096], Z[:4096], Xs, idx[:, :2], cols)      # warm-up (context, pinned pool)
t0 = time.perf_counter()
d = core.randomize_cov(Y, Z, Xs, idx, cols)
t1 = time.perf_counter()
print(f"randomize_cov: V={V} n={n} R={R}: {t1 - t0:.3f} s end to end from pageable host memory, last kernel window {core.last_kernel_ms():.2f} ms")
print("null 95% |t|:", np.quantile(d, 0.95, axis=0))
PY
echo "=== done"
