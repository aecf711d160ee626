#!/usr/bin/env python
"""mincTFCE at config-2 volume size (the bench's leg alone): 4 smooth unit-variance maps, d = 0.1, both sides.
python tools/vol_tfce_probe.py [calls]   -- prints the wall clock of each rmg_volume_tfce call."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rminc_b200 import core  # noqa: E402

DIMS = (180, 252, 156)
calls = int(sys.argv[1]) if len(sys.argv) > 1 else 3
g = torch.Generator(device="cuda:0").manual_seed(11)
vols = torch.randn((4, 1) + DIMS, generator=g, device="cuda:0", dtype=torch.float32)
for _ in range(2):
    vols = torch.nn.functional.avg_pool3d(vols, 3, stride=1, padding=1)
vols = (vols / vols.std(dim=(2, 3, 4), keepdim=True)).to(torch.float64)
Mh = np.asfortranarray(vols.reshape(4, -1).T.cpu().numpy())
del vols
core.volume_tfce(Mh[:, :1], DIMS, d=0.1)
for _ in range(calls):
    t0 = time.perf_counter()
    tv = core.volume_tfce(Mh, DIMS, d=0.1)
    print(f"rmg_volume_tfce, 4 maps, both sides: {time.perf_counter() - t0:.4f} s", flush=True)
print("max", float(np.abs(tv).max()))
