#!/bin/bash
# gpurun call: launch list (gpu__time_duration) of one volume TFCE call at config-2 volume size.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/voltfce_prof.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
cat > tools/_vt.py <<'PY'
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time, numpy as np, torch
from rminc_b200 import core
dims = (180, 252, 156)
g = torch.Generator(device="cuda").manual_seed(11)
vols = torch.randn((2, 1) + dims, generator=g, device="cuda", dtype=torch.float32)
for _ in range(2):
    vols = torch.nn.functional.avg_pool3d(vols, 3, stride=1, padding=1)
vols = (vols / vols.std(dim=(2, 3, 4), keepdim=True)).to(torch.float64)
M = np.asfortranarray(vols.reshape(2, -1).T.cpu().numpy())
noise = np.random.default_rng(0).standard_normal(M.shape[0])          # an iid map: many tiny clusters
for name, m in (("smooth", M[:, 0]), ("iid", noise)):
    for it in range(2):
        t0 = time.perf_counter(); out = core.volume_tfce(m, dims, d=0.1, side="positive"); dt = time.perf_counter() - t0
    print(f"{name}: one map, positive side: {dt:.3f} s; max {m.max():.2f}")
PY
python tools/_vt.py
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_voltfce.csv python tools/_vt.py > /dev/null 2>&1
echo "ncu rc=$?"
python - <<'PY'
import csv, collections, re
rows = [r for r in csv.reader(open("gpurun_out/launches_voltfce.csv")) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki]); v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else v * (1e3 if r[ui] == "ms" else 1)    # -> us
    tot[name] += v; cnt[name] += 1
s = sum(tot.values())
for k, v in tot.most_common(12):
    print(f"{k:50s} {cnt[k]:6d} launches {v / 1e3:10.3f} ms {100 * v / s:5.1f} %")
PY
echo "=== done"
