#!/bin/bash
# One gpurun call: parity tests of the voxel-wise covariate path + a config-2-sized timing.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/cov.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "=== pytest covariate"
timeout -k 10 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "covariate" 2>&1 | tail -40
echo "=== timing"
timeout -k 10 600 python - <<'PY'
import numpy as np, torch, time
from rminc_b200 import core
import sys; sys.path.insert(0, 'tests'); from helpers import design_cfg2
n, V = 500, 180*252*156 // 2
X = design_cfg2(n)
g = torch.Generator(device="cuda").manual_seed(1)
Y = torch.empty((n, V), dtype=torch.float64, device="cuda").normal_(0, 0.2, generator=g)
Z = torch.empty((n, V), dtype=torch.float64, device="cuda").normal_(100, 3, generator=g)
out = torch.empty((13, V), dtype=torch.float64, device="cuda")
for ps, Xs in ((3, X[:, :3]), (1, None)):
    o = out[:2 * (ps + 1) + 3]
    for _ in range(3): core.lm_fit_cov_torch(Y, Z, Xs, out=o)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): core.lm_fit_cov_torch(Y, Z, Xs, out=o)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    byt = V * (16 * n + 8 * (2 * (ps + 1) + 3))
    print(f"cov ps={ps}: {ms:.3f} ms, {byt/ms/1e6:.0f} GB/s, {V/ms/1e3:.1f} Mvox/s")
PY
echo "=== done"
