#!/bin/bash
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/u16_debug.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
cat > tools/_u16.py <<'PY'
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rminc_b200 import core, _lib
DIMS = (180, 252, 156); V = int(np.prod(DIMS)); n, p = 500, 4
lr = int(os.environ.get("LOCAL_RANK", 0)); torch.cuda.set_device(lr); dev = torch.device(f"cuda:{lr}")
g = torch.Generator(device=dev); g.manual_seed(101)
hU = torch.empty((n, V), dtype=torch.uint16, pin_memory=True)
for j0 in range(0, n, 20):
    blk = torch.empty((20, V), dtype=torch.float32, device=dev).normal_(30000.0, 400.0, generator=g)
    hU[j0:j0 + 20].copy_(blk.round_().clamp_(0, 32767).to(torch.int16).view(torch.uint16))
del blk
nsl, slice_len = DIMS[0], DIMS[1] * DIMS[2]
rs = np.random.default_rng(5)
smin = rs.normal(-0.8, 0.05, (n, nsl)); smax = smin + rs.uniform(1.5, 2.5, (n, nsl))
sc_h, of_h = np.ascontiguousarray((smax - smin) / 65535.0), np.ascontiguousarray(smin)
rng = np.random.default_rng(1)
X = np.asfortranarray(np.column_stack([np.ones(n), np.arange(n) % 2, np.round(rng.normal(60, 10, n), 1), rng.integers(0, 2, n)]).astype(float))
R = 1000
idx = np.asfortranarray(np.column_stack([np.random.default_rng(4).permutation(n) for _ in range(R)]).astype(np.int32))
cols = np.arange(p + 2, 2 * p + 2, dtype=np.int32)
Uh = hU.numpy().T
lib = _lib.load(); err = _lib.errbuf()
dist = np.empty((R, p), order="F")
def run(sa, sb, R_):
    s0, s1 = sa // slice_len, (sb + slice_len - 1) // slice_len
    sc_r, of_r = np.ascontiguousarray(sc_h[:, s0:s1]), np.ascontiguousarray(of_h[:, s0:s1])
    t0 = time.perf_counter()
    rc = lib.rmg_randomize_stored(Uh.ctypes.data + 2 * sa, core.STORED_U16, sb - sa, V, n, sc_r.ctypes.data, of_r.ctypes.data, 0.0,
                                  slice_len, X.ctypes.data, p, None, 1.0, 99999999.0, idx.ctypes.data, R_, cols.ctypes.data, p, 1,
                                  dist.ctypes.data, lr, err, len(err))
    _lib.check(rc, err)
    return time.perf_counter() - t0
run(0, V, 8)
for name, (sa, sb) in (("whole", (0, V)), ("first half", (0, V // 2)), ("second half", (V // 2, V))):
    print(f"[rank {lr}] {name}: {run(sa, sb, R):.3f} s, window {core.last_kernel_ms():.1f} ms", flush=True)
PY
echo "--- two processes, one per GPU, concurrently (no NCCL)"
LOCAL_RANK=0 timeout -k 10 300 python tools/_u16.py > gpurun_out/u16_r0.log 2>&1 &
bg=$!
LOCAL_RANK=1 timeout -k 10 300 python tools/_u16.py > gpurun_out/u16_r1.log 2>&1
wait $bg      # never a bare `wait` here: it would also wait for the tee of the exec redirection above, which never ends
cat gpurun_out/u16_r0.log gpurun_out/u16_r1.log | tail -20
echo "=== done"
