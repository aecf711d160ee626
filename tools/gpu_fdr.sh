#!/bin/bash
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/fdr.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "=== pytest fdr / stored tma / api"
timeout -k 10 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "fdr or FDR or stored or file_covariate" 2>&1 | tail -30
echo "=== fdr timing at config-2 size"
timeout -k 10 600 python - <<'PY'
import numpy as np, torch, time
from rminc_b200 import core
V = 180*252*156
g = torch.Generator(device="cuda").manual_seed(1)
t = torch.empty(V, dtype=torch.float64, device="cuda").normal_(0, 1.2, generator=g)
for _ in range(2): q, th = core.fdr_torch(t, core.STAT_T, 496)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): q, th = core.fdr_torch(t, core.STAT_T, 496)
torch.cuda.synchronize(); print("fdr t column, V=%d: %.2f ms/call" % (V, (time.perf_counter()-t0)/5*1e3), th)
PY
echo "=== done"
