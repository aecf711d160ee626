#!/usr/bin/env python
"""Benchmark of the voxel-wise OLS hot path (BASELINE.json metric: mincLm voxels/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the hot path (one whole mincLm fit) over one batch of synthetic
input.  Workload at N = 1: BASELINE config 2 -- mincLm(y ~ group + age + sex) on 500 synthetic
180x252x156 volumes (V = 7,076,160 voxels, n = 500, p = 4, FP64, no mask), Y resident in HBM
when the timed region starts.  Under torchrun (N > 1) every rank fits its own V-voxel shard
(weak scaling: the path shards over voxels with no data-path collective) and rank 0 prints
the ONE JSON line.  Extra keys: roofline (HBM), cpu_baseline (the oracle port on this box's
host cores), e2e (host buffers through the C ABI, H2D/D2H inside the timed region), clocks,
gpu_launches, and `randomize` (mincRandomize permutations/s, voxel-sharded, strong scaling,
one NCCL max all-reduce).

`--impl reference` times the reference's own CPU implementation of the path (oracle/_ref when
it was built from /root/reference, else the oracle port) with all host cores, forked over
contiguous voxel chunks exactly as mincLm(parallel = c("local", N)) does.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIMS = (180, 252, 156)
N_SUBJ = 500
P = 4
METRIC = "mincLm voxels/s"
UNIT = "voxels/s"
WORKLOAD = "mincLm(y ~ group + age + sex), 500 x 180x252x156 volumes (V=7076160, n=500, p=4, FP64, no mask)"


def design(n=N_SUBJ, seed=1):
    rng = np.random.default_rng(seed)
    group = (np.arange(n) % 2).astype(float)
    age = np.round(rng.normal(60, 10, n), 1)
    sex = rng.integers(0, 2, n).astype(float)
    return np.column_stack([np.ones(n), group, age, sex])


# --------------------------------------------------------------------------- CPU reference
def _cpu_worker(args):
    kind, a, b = args
    from oracle import oracle as O
    Y, X, out = _cpu_worker.Y, _cpu_worker.X, _cpu_worker.out
    if b <= a:
        return 0
    res = O.vertex_lm_loop(Y[a:b], X, use_ref=(kind == "reference"))
    out[a:b] = res
    return b - a


def cpu_reference_run(sample_V, procs, steps=1, warmup=0, seed=0):
    """Times the reference's CPU path on `sample_V` voxels of the benchmark workload with
    `procs` forked workers over contiguous voxel chunks (R/minc_voxel_statistics.R:865-873).
    Returns (voxels_per_s, kind, seconds_per_step)."""
    import multiprocessing as mp
    from oracle import oracle as O
    O.build(ref=True)
    kind = "reference" if O.have_ref() else "port"
    rng = np.random.default_rng(seed)
    X = design()
    Y = np.asfortranarray(0.2 * rng.standard_normal((sample_V, N_SUBJ)))
    shm = mp.RawArray("d", sample_V * (2 * P + 3))
    out = np.frombuffer(shm, dtype=np.float64).reshape(sample_V, 2 * P + 3)
    _cpu_worker.Y, _cpu_worker.X, _cpu_worker.out = Y, X, out
    bounds = np.linspace(0, sample_V, procs + 1).astype(int)
    tasks = [(kind, int(bounds[g]), int(bounds[g + 1])) for g in range(procs)]
    ctx = mp.get_context("fork")
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        with ctx.Pool(procs) as pool:       # a fresh fork per call, as mclapply does
            done = sum(pool.map(_cpu_worker, tasks))
        dt = time.perf_counter() - t0
        assert done == sample_V
        if it >= warmup:
            times.append(dt)
    assert np.isfinite(out[:, 0]).all()
    sec = float(np.mean(times))
    return sample_V / sec, kind, sec


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample_V = int(os.environ.get("RMG_BENCH_CPU_SAMPLE", 60000 * cores))
    vps, kind, sec = cpu_reference_run(sample_V, cores, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": vps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n": N_SUBJ, "p": P},
        "cpu_baseline": {"value": vps, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{sample_V} contiguous voxels of the workload per step, in memory (no MINC I/O), "
                                   f"{cores} forked workers over contiguous chunks like mincLm(parallel=c('local',{cores}))"},
        "e2e": {"value": vps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw"

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[6]))
            except ValueError:
                continue
            for nm, val in zip(names, f[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "power_w_max": max(pw) if pw else None, "samples": len(sm)}


# --------------------------------------------------------------------------- GPU arm
def bind_to_gpu_numa(index):
    """Run this process on the CPUs NVML reports as local to GPU `index`, so that the pinned host
    buffers of the end-to-end leg are first-touched on the GPU's NUMA node.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = ((os.cpu_count() or 64) + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {w * 64 + i for w in range(words) for i in range(64) if (mask[w] >> i) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return sorted(cpus)
    except Exception:
        pass
    return None


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def gen_Y(torch, V, n, device, seed, mean=0.0, sd=0.2):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    Yt = torch.empty((n, V), dtype=torch.float64, device=device)
    for j0 in range(0, n, 20):
        j1 = min(n, j0 + 20)
        Yt[j0:j1].normal_(mean, sd, generator=g)
    return Yt


def pinned_numpy(lib, shape, dtype):
    """A page-locked, NUMA-interleaved host array from the library's own allocator (rmg_host_alloc); (array, pointer)."""
    import ctypes as C
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    ptr = lib.rmg_host_alloc(nbytes)
    if not ptr:
        raise SystemExit("bench.py: rmg_host_alloc failed")
    buf = (C.c_char * nbytes).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape), ptr


def files_leg(args, core):
    """From FILES to the result: a study of MINC 2 volumes (config-2 dimensions, uint16, one deflated chunk per slice, per-slice
    real ranges) written by the test-side writer, decoded by the library's own reader on all host cores into page-locked
    memory and fitted in its stored type.  What the reference does here: libminc + HDF5, one thread, doubles
    (src/modelling_functions.c:1037-1055)."""
    import shutil
    import tempfile
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests"))
    import h5mini
    from rminc_b200 import minc2
    nf = int(os.environ.get("RMG_BENCH_FILES", 32))
    root = tempfile.mkdtemp(prefix="rmg_files_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        rng = np.random.default_rng(11)
        base = rng.normal(30000.0, 400.0, DIMS)
        paths, t_w = [], time.perf_counter()
        for j in range(nf):
            vol = np.clip(np.rint(base + rng.normal(0.0, 150.0, DIMS)), 0, 65535).astype(np.uint16)
            lo = rng.normal(-0.8, 0.05, DIMS[0])
            f = os.path.join(root, f"s{j:03d}.mnc")
            h5mini.write_minc2(f, vol, image_min=lo, image_max=lo + rng.uniform(1.5, 2.5, DIMS[0]), valid_range=(0, 65535),
                               chunks=(1, DIMS[1], DIMS[2]), compress=1)
            paths.append(f)
        t_w = time.perf_counter() - t_w
        fbytes = sum(os.path.getsize(f) for f in paths)
        V = int(np.prod(DIMS))
        X = design(nf) if nf > 8 else design(N_SUBJ)[:nf]
        threads = minc2.host_threads()
        t0 = time.perf_counter()
        U1 = minc2.read_table(paths[:2], n_threads=1)[0]
        t_one = (time.perf_counter() - t0) / 2
        del U1
        best = None
        t0 = time.perf_counter()
        U = minc2.read_table(paths, pinned=True)[0]              # first call: page-locks the staging matrix (kept below)
        t_first = time.perf_counter() - t0
        for _ in range(3):
            t0 = time.perf_counter()
            U, lo, hi, vr, info = minc2.read_table(paths, out=U)
            t1 = time.perf_counter()
            out = core.lm_fit_stored(U, X, (hi - lo) / (vr[1] - vr[0]), lo, voxel_min=vr[0], slice_len=info.slice_len)
            t2 = time.perf_counter()
            if best is None or t2 - t0 < best[0]:
                best = (t2 - t0, t1 - t0, t2 - t1)
        del U
        assert np.isfinite(out[:, 0]).all()
        return {"workload": f"{nf} MINC 2 files of 180x252x156 uint16 voxels (deflate, one chunk per slice, per-slice ranges) "
                            "-> decode on the host -> stored-type fit", "files": nf, "file_bytes": fbytes, "host_threads": threads,
                "seconds": best[0], "decode_s": best[1], "fit_s": best[2], "voxels_per_s": V / best[0],
                "decode_stored_gbs": 2.0 * V * nf / best[1] / 1e9, "decode_file_gbs": fbytes / best[1] / 1e9,
                "one_thread_s_per_file": t_one, "speedup_vs_one_thread": t_one * nf / best[1],
                "first_call_s_incl_page_locking": t_first,
                "writer_s": t_w, "timing": "wall clock, best of 3 into a page-locked matrix kept across calls (files in /dev/shm: no disk)"}
    finally:
        shutil.rmtree(root, ignore_errors=True)


def single_process_multi(args, core, Yt, X, G, d_ref, perms, idx):
    """Rank 0 alone drives G GPUs through the C ABI (what an R session does): rmg_h2d_probe (the box's concurrent
    host-link ceiling), rmg_lm_fit_multi and rmg_randomize_multi on ONE pinned config-2 matrix strong-sharded over the
    G links, then the resident handle: one upload -> fit -> randomize -> FDR."""
    import ctypes as C
    import torch
    from rminc_b200 import _lib
    lib = _lib.load()
    err = _lib.errbuf()
    n, V = Yt.shape
    p = X.shape[1]
    res = {"n_gpus": G, "api": "one process, one host thread; voxel shards over devices 0..G-1"}
    Yh_nv, yptr = pinned_numpy(lib, (n, V), np.float64)           # (n, V) C-order == V x n column-major
    Oh_cv, optr = pinned_numpy(lib, (2 * p + 3, V), np.float64)
    torch.from_numpy(Yh_nv).copy_(Yt)
    Xf = np.asfortranarray(X)
    # -- the ceiling: every link at once, 2 GiB per device per repeat
    per = min(2 << 30, (V * n * 8) // G // 4096 * 4096)
    gbs = (C.c_double * G)()
    agg = C.c_double(0.0)
    _lib.check(lib.rmg_h2d_probe(yptr, per, G, 3, gbs, C.byref(agg), err, len(err)), err)
    one = C.c_double(0.0)
    _lib.check(lib.rmg_h2d_probe(yptr, per, 1, 3, None, C.byref(one), err, len(err)), err)
    res["h2d_probe"] = {"aggregate_gbs": agg.value, "per_gpu_gbs": [round(x, 2) for x in gbs], "one_gpu_alone_gbs": one.value,
                        "bytes_per_gpu": per, "memory": "rmg_host_alloc: pinned, pages interleaved over the NUMA nodes"}
    # -- mincLm, strong-sharded: one 28.3 GB volume over G links
    def fit():
        _lib.check(lib.rmg_lm_fit_multi(yptr, V, V, n, Xf.ctypes.data, p, None, 1.0, 99999999.0, optr, V, G, err, len(err)), err)
    fit()
    k = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(k):
        fit()
    dt = (time.perf_counter() - t0) / k
    h2d = V * n * 8
    res["fit_e2e"] = {"value": V / dt, "unit": UNIT, "seconds": dt, "n_gpus": G, "scaling": "strong", "h2d_bytes_per_step": h2d,
                      "d2h_bytes_per_step": V * (2 * p + 3) * 8, "host_link_gbs": h2d / dt / 1e9,
                      "frac_of_h2d_probe": h2d / dt / 1e9 / agg.value,
                      "api": "rmg_lm_fit_multi (C ABI; pinned host Y, V x n doubles, contiguous voxel shards over G links)"}
    ref = core.lm_fit_torch(Yt[:, :4096], X)
    far = core.lm_fit_torch(Yt[:, V - 4096:], X)
    torch.cuda.synchronize()
    parity = True
    for blk, r in ((slice(0, 4096), ref), (slice(V - 4096, V), far)):
        chk = torch.from_numpy(np.ascontiguousarray(Oh_cv[:, blk])).to(r.device)
        scale = torch.maximum(r.abs(), r.square().mean(dim=1, keepdim=True).sqrt())
        parity = parity and bool(((chk - r).abs() <= 1e-9 * scale).all())
    # -- mincRandomize from the same host matrix, strong-sharded; upload hidden behind the permutation GEMM
    if idx is not None:
        cols_a = np.arange(p + 2, 2 * p + 2, dtype=np.int32)
        idx_f = np.asfortranarray(idx, dtype=np.int32)
        dist_h = np.empty((perms, p), order="F")

        def perm(R_):
            _lib.check(lib.rmg_randomize_multi(yptr, V, V, n, Xf.ctypes.data, p, None, 1.0, 99999999.0, idx_f.ctypes.data, R_,
                                               cols_a.ctypes.data, p, 1, dist_h.ctypes.data, G, err, len(err)), err)
        perm(8)
        t0 = time.perf_counter()
        perm(perms)
        sec = time.perf_counter() - t0
        res["randomize_e2e"] = {"value": perms / sec, "unit": "perms/s", "seconds": sec, "R": perms, "n_gpus": G, "h2d_bytes": h2d,
                                "gemm_window_ms": core.last_kernel_ms(),
                                "api": "rmg_randomize_multi (C ABI; pinned host Y -> voxel blocks per device on copy streams -> "
                                       "all R permutations per block as it lands; designs factored and packed once for all devices)"}
        # reference: the device-resident one-GPU path on rank 0's own matrix (under torchrun every rank holds a
        # different synthetic volume, so the distributed result is not comparable), first 64 permutations
        d64 = core.randomize_torch(Yt, X, idx[:, :64], list(range(p + 2, 2 * p + 2)))
        torch.cuda.synchronize()
        d64 = d64.cpu().numpy().T
        parity = parity and bool(np.allclose(dist_h[:64], d64, rtol=1e-9, atol=0))
    # -- the resident handle: one upload, then fit -> randomize -> FDR on the shards
    h = C.c_void_p()
    t0 = time.perf_counter()
    _lib.check(lib.rmg_dataset_create(yptr, V, V, n, None, 1.0, 99999999.0, G, C.byref(h), err, len(err)), err)
    _lib.check(lib.rmg_dataset_sync(h, err, len(err)), err)
    t_up = time.perf_counter() - t0
    t0 = time.perf_counter()
    _lib.check(lib.rmg_dataset_lm_fit(h, Xf.ctypes.data, p, optr, V, err, len(err)), err)
    t_fit = time.perf_counter() - t0
    fit_kernel_ms = core.last_kernel_ms()
    t0 = time.perf_counter()
    _lib.check(lib.rmg_dataset_lm_fit(h, Xf.ctypes.data, p, None, V, err, len(err)), err)
    t_fit_resident = time.perf_counter() - t0
    ds = {"upload_seconds": t_up, "upload_gbs": h2d / t_up / 1e9, "lm_fit_seconds_incl_d2h": t_fit,
          "lm_fit_seconds_result_kept_on_device": t_fit_resident, "lm_fit_kernel_ms": fit_kernel_ms}
    if idx is not None:
        for name in ("randomize_seconds_first", "randomize_seconds_again"):
            t0 = time.perf_counter()
            _lib.check(lib.rmg_dataset_randomize(h, Xf.ctypes.data, p, idx_f.ctypes.data, perms, cols_a.ctypes.data, p, 1,
                                                 dist_h.ctypes.data, err, len(err)), err)
            ds[name] = time.perf_counter() - t0
        ds["randomize_perms_per_s"] = perms / ds["randomize_seconds_again"]
        ds["randomize_gemm_window_ms"] = core.last_kernel_ms()
        parity = parity and bool(np.allclose(dist_h[:64], d64, rtol=1e-9, atol=0))
    q = np.empty(V)
    th = np.empty(5)
    t0 = time.perf_counter()
    _lib.check(lib.rmg_dataset_fdr(h, p + 3, core.STAT_T, float(n - p), 0.0, q.ctypes.data, th.ctypes.data, err, len(err)), err)
    ds["fdr_seconds_one_column"] = time.perf_counter() - t0
    parity = parity and bool(np.isfinite(q).all()) and bool((q >= 0).all() and (q <= 1).all())
    lib.rmg_dataset_free(h)
    ds["sequence"] = "rmg_dataset_create -> rmg_dataset_lm_fit -> rmg_dataset_randomize (x2) -> rmg_dataset_fdr -> rmg_dataset_free"
    res["resident_dataset"] = ds
    res["multi_parity"] = parity
    lib.rmg_host_free(yptr)
    lib.rmg_host_free(optr)
    return res


def config5(args, torch, dist, core, dev, rank, world, local_rank, barrier, design_fn):
    """BASELINE config 5: mincLm on a 400x560x340 volume x 200 subjects, voxel-sharded across 8 GPUs.  Every rank fits
    shard `rank` of 8 (9.52 M voxels, 15.23 GB of doubles generated on the device); rank 0 also fits the WHOLE volume
    as uint16 on one GPU (30.5 GB; int64 offsets) and every rank times its shard end to end from pinned host memory."""
    import psutil
    from rminc_b200 import _lib
    lib = _lib.load()
    err = _lib.errbuf()
    dims5, n5, p5, shards = (400, 560, 340), 200, 4, 8
    V5 = int(np.prod(dims5))
    Vs = V5 // shards
    X5 = design_fn(n5, seed=5)
    stream = torch.cuda.current_stream()
    torch.cuda.empty_cache()                      # blocks cached by the earlier legs would hide the allocations below
    torch.cuda.reset_peak_memory_stats()
    free0, total = torch.cuda.mem_get_info()
    Ys = gen_Y(torch, Vs, n5, dev, seed=500 + rank)
    outs = torch.empty((2 * p5 + 3, Vs), dtype=torch.float64, device=dev)
    for _ in range(3):
        core.lm_fit_torch(Ys, X5, out=outs)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        core.lm_fit_torch(Ys, X5, out=outs)
    e1.record(stream)
    barrier()
    tt = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = float(tt.item())
    bytes_shard = Vs * (8 * n5 + 8 * (2 * p5 + 3))
    peak = measured_hbm_peak()[0]
    res = {"workload": f"mincLm, 400x560x340 x 200 subjects, p = 4: {shards} voxel shards of {Vs} voxels (15.23 GB each)",
           "shards_fitted": world, "shards_total": shards, "ms_per_shard": ms, "voxels_per_s": world * Vs / (ms * 1e-3),
           "whole_volume_seconds_on_8_gpus": ms * 1e-3 * max(1, shards // world),
           "achieved_gbs_per_gpu": bytes_shard / (ms * 1e-3) / 1e9, "frac": bytes_shard / (ms * 1e-3) / 1e9 / peak,
           "kernel": core.last_fit_variant(), "hbm_used_gb_per_gpu": (free0 - torch.cuda.mem_get_info()[0]) / 1e9,
           "hbm_resident_gb_per_gpu": (Ys.numel() + outs.numel()) * 8 / 1e9, "hbm_total_gb": total / 1e9}
    # end to end from this rank's pinned host shard
    need = Vs * n5 * 8
    if psutil.virtual_memory().available > need * world * 1.2 + (8 << 30):
        bind_to_gpu_numa(local_rank)
        hY = torch.empty((n5, Vs), dtype=torch.float64, pin_memory=True)
        hY.copy_(Ys)
        hO = torch.empty((2 * p5 + 3, Vs), dtype=torch.float64, pin_memory=True)
        Xf = np.asfortranarray(X5)

        def step():
            _lib.check(lib.rmg_lm_fit(hY.data_ptr(), Vs, Vs, n5, Xf.ctypes.data, p5, None, 1.0, 99999999.0, hO.data_ptr(), Vs,
                                      local_rank, err, len(err)), err)
        step()
        barrier()
        t0 = time.perf_counter()
        step()
        barrier()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        sec = float(tt.item())
        res["e2e"] = {"value": world * Vs / sec, "unit": UNIT, "seconds": sec, "h2d_bytes_per_gpu": need,
                      "host_link_gbs_per_gpu": need / sec / 1e9, "api": "rmg_lm_fit per rank on its pinned host shard"}
        chk = hO[:, :4096].to(dev)
        scl = torch.maximum(outs[:, :4096].abs(), outs[:, :4096].square().mean(dim=1, keepdim=True).sqrt())
        assert bool(((chk - outs[:, :4096]).abs() <= 1e-9 * scl).all()), "config-5 e2e result differs from the device result"
        del hY, hO
    del Ys, outs
    torch.cuda.empty_cache()
    # the whole volume as uint16 on ONE GPU (rank 0): 15.2e9 samples -> 64-bit offsets everywhere
    if rank == 0:
        g = torch.Generator(device=dev).manual_seed(55)
        U = torch.empty((n5, V5), dtype=torch.uint16, device=dev)
        for j0 in range(0, n5, 10):
            j1 = min(n5, j0 + 10)
            blk = torch.empty((j1 - j0, V5), dtype=torch.float32, device=dev).normal_(30000.0, 400.0, generator=g)
            U[j0:j1] = blk.round_().clamp_(0, 32767).to(torch.int16).view(torch.uint16)
        del blk
        nsl, slice_len = dims5[0], dims5[1] * dims5[2]
        rs = np.random.default_rng(6)
        smin = rs.normal(-0.8, 0.05, (n5, nsl))
        sc = torch.from_numpy(np.ascontiguousarray(rs.uniform(1.5, 2.5, (n5, nsl)) / 65535.0)).to(dev)
        of = torch.from_numpy(np.ascontiguousarray(smin)).to(dev)
        ou = torch.empty((2 * p5 + 3, V5), dtype=torch.float64, device=dev)
        for _ in range(2):
            core.lm_fit_stored_torch(U, X5, sc, of, slice_len=slice_len, out=ou)
        variant_u = core.last_fit_variant()
        torch.cuda.synchronize()
        eu0, eu1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(1, args.steps // 2)
        tw = time.perf_counter()
        eu0.record(stream)
        for _ in range(reps):
            core.lm_fit_stored_torch(U, X5, sc, of, slice_len=slice_len, out=ou)
        eu1.record(stream)
        torch.cuda.synchronize()
        ms_wall = (time.perf_counter() - tw) * 1e3 / reps
        msu = eu0.elapsed_time(eu1) / reps
        # int64 sanity row: the last 4096 voxels start 15.2e9 samples into U; refit them alone (as a shard that keeps its
        # global slice index) and compare
        tail = core.lm_fit_stored_torch(U[:, V5 - 4096:], X5, sc, of, slice_len=slice_len, v_offset=V5 - 4096)
        torch.cuda.synchronize()
        scl = torch.maximum(tail.abs(), tail.square().mean(dim=1, keepdim=True).sqrt())
        ok = bool(((ou[:, V5 - 4096:] - tail).abs() <= 1e-9 * scl).all()) and bool(torch.isfinite(ou[0]).all())
        bytes_u = V5 * (2 * n5 + 8 * (2 * p5 + 3))
        res["whole_volume_uint16_one_gpu"] = {"voxels": V5, "ms": msu, "ms_wall_clock": ms_wall, "voxels_per_s": V5 / (msu * 1e-3),
                                              "achieved_gbs": bytes_u / (msu * 1e-3) / 1e9,
                                              "hbm_resident_gb": (U.numel() * 2 + ou.numel() * 8) / 1e9,
                                              "hbm_peak_allocated_gb": torch.cuda.max_memory_allocated() / 1e9,
                                              "first_sample_offset_of_the_checked_rows": int((V5 - 4096)),
                                              "samples_before_the_last_subject_row": int((n5 - 1) * V5),
                                              "int64_offset_rows_match": ok, "kernel": variant_u}
        del U, ou, sc, of
        torch.cuda.empty_cache()
    barrier()
    return res


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from rminc_b200 import build
    build.build()
    from rminc_b200 import core, parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if core.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device; rminc_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    host_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        host_group = dist.new_group(backend="gloo")      # host-side barrier: waiting ranks must not spin a kernel

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    V = int(np.prod(DIMS)) if not args.small else 64 * 1024
    n, p = N_SUBJ, P
    X = design()
    Yt = gen_Y(torch, V, n, dev, seed=1 + rank)
    out = torch.empty((2 * p + 3, V), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream()

    # ---- timed region 1: Y resident in HBM.  Clocks are sampled from the warm-up on (the timed
    #      region alone is a few tens of ms, shorter than nvidia-smi's sampling period).
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        core.lm_fit_torch(Yt, X, out=out)
    torch.cuda.synchronize()
    t_w = time.perf_counter()
    while time.perf_counter() - t_w < float(os.environ.get("RMG_BENCH_SPIN_S", "0.6")):      # keep the GPU under the same load while nvidia-smi samples
        core.lm_fit_torch(Yt, X, out=out)
        torch.cuda.synchronize()
    barrier()
    launches0 = core.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record(stream)
    for i in range(args.steps):
        core.lm_fit_torch(Yt, X, out=out)
        ev[i + 1].record(stream)
    barrier()
    launches = core.launch_count() - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * V * args.steps / (total_ms * 1e-3)

    # roofline of the dominant (only) kernel: algorithmic bytes per launch / its launch duration
    alg_bytes = V * (8 * n + 8 * (2 * p + 3))
    kern_ms = float(np.mean(step_ms))            # one fit kernel per step; the 2 tiny H2D copies precede it
    peak, peak_src = measured_hbm_peak()
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch, from the committed ncu --set full capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json"))).get(core.last_fit_variant())
        if tj and not args.small:
            traffic, traffic_src = tj["dram_bytes_read"] + tj["dram_bytes_write"], tj["source"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                "kernel": core.last_fit_variant(), "kernel_ms": kern_ms,
                "step_ms": [round(x, 4) for x in step_ms]}

    # ---- masked variant (ellipsoid, ~38 % inside), reported beside the headline
    masked = None
    if rank == 0 and not args.small and "masked" in args.parts:
        z, y, x = np.meshgrid(*[np.arange(d) for d in DIMS], indexing="ij")
        c = [(d - 1) / 2 for d in DIMS]
        inside = sum(((a - cc) / (0.45 * d)) ** 2 for a, cc, d in zip((z, y, x), c, DIMS)) <= 1
        mt = torch.from_numpy(inside.reshape(-1).astype(np.float64)).to(dev)
        core.lm_fit_torch(Yt, X, mask=mt, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            core.lm_fit_torch(Yt, X, mask=mt, out=out)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        fin = float(inside.mean())
        mb = V * (8 + 8 * (2 * p + 3)) + int(inside.sum()) * 8 * n
        masked = {"inside_frac": fin, "ms_per_step": ms, "voxels_per_s": V / (ms * 1e-3),
                  "achieved_gbs": mb / (ms * 1e-3) / 1e9}
        del mt

    # ---- vertexLm (config 3): 81,924 vertices x 2000 subjects, p = 7 -- small V, large n (split-n)
    vertex = None
    if rank == 0 and "vertex" in args.parts:
        Vv, nv, pv = 81924, 2000, 7
        rngv = np.random.default_rng(3)
        Xv = np.column_stack([np.ones(nv)] + [rngv.normal(size=nv) for _ in range(pv - 1)])
        Yv = gen_Y(torch, Vv, nv, dev, seed=3, mean=2.5, sd=0.4)
        ov = torch.empty((2 * pv + 3, Vv), dtype=torch.float64, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # 1.3 GB input < 10x L2: flush between steps
        times, host_us = [], []
        for i in range(3 + args.steps):
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(10_000_000)      # ~1.5 ms of GPU spin: the host-side part of the call (design
            e0.record(stream)                 # factorisation, upload, launch) overlaps it, so e0..e1 is the
            t0 = time.perf_counter()          # device time of the call, not the launch latency
            core.lm_fit_torch(Yv, Xv, out=ov)
            host_us.append((time.perf_counter() - t0) * 1e6)
            e1.record(stream)
            torch.cuda.synchronize()
            if i >= 3:
                times.append(e0.elapsed_time(e1))
        ms = float(np.median(times))
        vb = Vv * (8 * nv + 8 * (2 * pv + 3))
        vertex = {"workload": "vertexLm, 81,924 vertices x 2000 subjects, p = 7", "ms_per_step": ms,
                  "vertices_per_s": Vv / (ms * 1e-3), "achieved_gbs": vb / (ms * 1e-3) / 1e9,
                  "frac_of_hbm_peak": vb / (ms * 1e-3) / 1e9 / measured_hbm_peak()[0],
                  "l2": "256 MiB write between steps flushes L2", "kernel": core.last_fit_variant(),
                  "host_call_us": float(np.median(host_us[3:]))}
        del Yv, ov, flush

    # ---- voxel-wise covariate: mincLm(y ~ group + age + otherVolume), two streamed operands
    cov = None
    if rank == 0 and not args.small and "cov" in args.parts:
        Zt = gen_Y(torch, V, n, dev, seed=77, mean=100.0, sd=3.0)
        oc = torch.empty((2 * p + 3, V), dtype=torch.float64, device=dev)
        Xs = X[:, :p - 1]
        core.lm_fit_cov_torch(Yt, Zt, Xs, out=oc)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            core.lm_fit_cov_torch(Yt, Zt, Xs, out=oc)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        cb = V * (16 * n + 8 * (2 * p + 3))
        cov = {"workload": "mincLm(y ~ group + age + volume): per-voxel design [Xs | z_v], p = 4", "ms_per_step": ms,
               "voxels_per_s": V / (ms * 1e-3), "achieved_gbs": cb / (ms * 1e-3) / 1e9,
               "frac_of_hbm_peak": cb / (ms * 1e-3) / 1e9 / measured_hbm_peak()[0], "kernel": core.last_fit_variant()}
        del Zt, oc

    # ---- volumes in their stored type: uint16 voxels + per-slice real ranges, expanded in registers
    stored = None
    Ut = sc_t = of_t = None
    if not args.small and ("stored" in args.parts or "e2e_u16" in args.parts):
        g = torch.Generator(device=dev)
        g.manual_seed(100 + rank)
        Ut = torch.empty((n, V), dtype=torch.uint16, device=dev)
        for j0 in range(0, n, 20):
            j1 = min(n, j0 + 20)
            blk = torch.empty((j1 - j0, V), dtype=torch.float32, device=dev).normal_(30000.0, 400.0, generator=g)
            Ut[j0:j1] = blk.round_().clamp_(0, 32767).to(torch.int16).view(torch.uint16)
        del blk
        if os.environ.get("RMG_BENCH_EDGES"):
            # tissue / background transitions inside the 4 voxels a thread owns (an unmasked whole-head intensity image):
            # every RMG_BENCH_EDGES-th voxel is background (about 100 +- 30 counts next to 30000 +- 400)
            step = int(os.environ["RMG_BENCH_EDGES"])
            bg = torch.empty((n, (V + step - 1) // step), dtype=torch.float32, device=dev).normal_(100.0, 30.0, generator=g)
            Ut[:, ::step] = bg.round_().clamp_(0, 32767).to(torch.int16).view(torch.uint16)
            del bg
        if os.environ.get("RMG_BENCH_BACKGROUND"):       # the first third of the volume is background: stored 0 in every file
            Ut[:, :V // 3] = 0
        nsl, slice_len = DIMS[0], DIMS[1] * DIMS[2]
        rs = np.random.default_rng(5)
        smin = rs.normal(-0.8, 0.05, (n, nsl))
        smax = smin + rs.uniform(1.5, 2.5, (n, nsl))
        if os.environ.get("RMG_BENCH_FLAT_RANGES"):          # every file and slice with the same real range
            smin[:] = -0.8
            smax[:] = 1.2
        sc_h, of_h = np.ascontiguousarray((smax - smin) / 65535.0), np.ascontiguousarray(smin)   # (n, nslices) == [nslices x n] col-major
        sc_t, of_t = torch.from_numpy(sc_h).to(dev), torch.from_numpy(of_h).to(dev)
    if Ut is not None and "stored" in args.parts:
        os_ = torch.empty((2 * p + 3, V), dtype=torch.float64, device=dev)
        for _ in range(3):
            core.lm_fit_stored_torch(Ut, X, sc_t, of_t, slice_len=slice_len, out=os_)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            core.lm_fit_stored_torch(Ut, X, sc_t, of_t, slice_len=slice_len, out=os_)
        e1.record(stream)
        barrier()
        tt = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
        sb = V * (2 * n + 8 * (2 * p + 3))
        stored = {"workload": "the same fit on uint16 voxels + per-slice real ranges (180 slices x 500 files)",
                  "ms_per_step": ms, "voxels_per_s": world * V / (ms * 1e-3), "achieved_gbs": sb / (ms * 1e-3) / 1e9,
                  "frac_of_hbm_peak": sb / (ms * 1e-3) / 1e9 / measured_hbm_peak()[0],
                  "fp64_ops_per_s": V * n * (2.25 + p) / (ms * 1e-3), "kernel": core.last_fit_variant(),
                  # scalar FP64 issue limit measured by tools/dfma_bench.cu: one warp-wide DFMA per 2.2 cycles and scheduler
                  "fp64_issue_floor_ms": V * n * (2.25 + p) / (148 * 4 * 32 / 2.2 * 1.965e9) * 1e3,
                  "bound": "FP64 pipe co-limits: (2.25+p) FP64 operations per 2-byte sample (one FMA to expand with the "
                           "thread's shift folded in, square, p projections, one offset subtraction per 4 samples); "
                           "RMG_STORED_EXACT=1 (the reader's three roundings): (5+p)"}
        # the stored-type fit equals the double fit of the expanded samples
        Ye = ((Ut[:, :4096].to(torch.float64) - 0.0) * sc_t[:, :1]) + of_t[:, :1]
        ref = core.lm_fit_torch(Ye, X)
        torch.cuda.synchronize()
        scl = torch.maximum(ref.abs(), ref.square().mean(dim=1, keepdim=True).sqrt())
        if not os.environ.get("RMG_BENCH_BACKGROUND"):
            assert bool(((os_[:, :4096] - ref).abs() <= 1e-9 * scl).all()), "stored-type result differs from the double fit"
        if rank == 0:
            # the everyday form of the call: uint16 files AND a mask (38 % ellipsoid)
            z, y, x = np.meshgrid(*[np.arange(d) for d in DIMS], indexing="ij")
            c = [(d - 1) / 2 for d in DIMS]
            inside = sum(((a - cc) / (0.45 * d)) ** 2 for a, cc, d in zip((z, y, x), c, DIMS)) <= 1
            mt = torch.from_numpy(inside.reshape(-1).astype(np.float64)).to(dev)
            for _ in range(2):
                core.lm_fit_stored_torch(Ut, X, sc_t, of_t, slice_len=slice_len, mask=mt, out=os_)
            torch.cuda.synchronize()
            m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            m0.record(stream)
            for _ in range(args.steps):
                core.lm_fit_stored_torch(Ut, X, sc_t, of_t, slice_len=slice_len, mask=mt, out=os_)
            m1.record(stream)
            torch.cuda.synchronize()
            stored["masked"] = {"inside_frac": float(inside.mean()), "ms_per_step": m0.elapsed_time(m1) / args.steps,
                                "kernel": core.last_fit_variant()}
            sel = torch.nonzero(mt[:65536] > 0.5).flatten()[:2048]
            if sel.numel():
                Ym = ((Ut[:, sel].to(torch.float64) - 0.0) * sc_t[:, :1]) + of_t[:, :1]
                refm = core.lm_fit_torch(Ym, X)
                torch.cuda.synchronize()
                sclm = torch.maximum(refm.abs(), refm.square().mean(dim=1, keepdim=True).sqrt())
                assert bool(((os_[:, sel] - refm).abs() <= 1e-9 * sclm).all()), "masked stored-type result differs from the double fit"
            assert bool((os_[:, mt < 0.5] == 0).all()), "masked rows of the stored-type fit are not +0.0"
            del mt
        del os_, Ye

    # ---- vertexTFCE randomisation: config-3-sized surface (grid graph stand-in for the mesh), all t columns
    tfce = None
    if rank == 0 and "tfce" in args.parts:
        side_len, nv, pv, Rt = 286, 400, 4, args.tfce_perms
        Vt = side_len * side_len
        ii, jj = np.divmod(np.arange(Vt), side_len)
        nbrs = [np.where(ii > 0, np.arange(Vt) - side_len, -1), np.where(ii < side_len - 1, np.arange(Vt) + side_len, -1),
                np.where(jj > 0, np.arange(Vt) - 1, -1), np.where(jj < side_len - 1, np.arange(Vt) + 1, -1)]
        nb = np.stack(nbrs, axis=1)
        keep = nb >= 0
        aptr = np.zeros(Vt + 1, dtype=np.int32)
        aptr[1:] = np.cumsum(keep.sum(axis=1))
        aidx = nb[keep].astype(np.int32)
        rt = np.random.default_rng(8)
        Xt = np.column_stack([np.ones(nv)] + [rt.normal(size=nv) for _ in range(pv - 1)])
        Yt_h = np.asfortranarray(rt.normal(3.0, 0.4, (Vt, nv)))
        idx_t = np.column_stack([rt.permutation(nv) for _ in range(Rt)]).astype(np.int32)
        cols_t = list(range(pv + 2, 2 * pv + 2))
        core.randomize_tfce(Yt_h, Xt, idx_t[:, :4], cols_t, (aptr, aidx))          # warm-up
        # wall clock of single host calls moves with the box (page faults, host copy threads): median of three
        secs = []
        for _ in range(3):
            t0 = time.perf_counter()
            dt_ = core.randomize_tfce(Yt_h, Xt, idx_t, cols_t, (aptr, aidx))
            secs.append(time.perf_counter() - t0)
        sec = float(np.median(secs))
        from oracle import oracle as O
        O.build(ref=False)
        tmap = core.lm_fit(Yt_h[:, :nv], Xt)[:, cols_t[1]]
        c0 = time.perf_counter()
        O.graph_tfce(tmap, aptr, aidx)
        cpu_ms = (time.perf_counter() - c0) * 1e3
        # the same call with Y in page-locked memory (the library's allocator): what is left is refits + enhancement
        from rminc_b200 import _lib as _L
        flat, hptr = pinned_numpy(_L.load(), (Vt * nv,), np.float64)
        Yp = flat.reshape((Vt, nv), order="F")
        Yp[...] = Yt_h
        core.randomize_tfce(Yp, Xt, idx_t[:, :4], cols_t, (aptr, aidx))
        secs_pinned = []
        for _ in range(3):
            t0 = time.perf_counter()
            dp_ = core.randomize_tfce(Yp, Xt, idx_t, cols_t, (aptr, aidx))
            secs_pinned.append(time.perf_counter() - t0)
        sec_pinned = float(np.median(secs_pinned))
        # (not bit-equal: which root a concurrent hook attaches under decides the rounding of the offset sums)
        assert np.allclose(dp_, dt_, rtol=1e-10, atol=0), "TFCE randomisation: pinned and pageable input disagree"
        del Yp, flat
        _L.load().rmg_host_free(hptr)
        tfce = {"workload": f"vertexTFCE.vertexLm randomisation: {Vt} vertices x {nv} subjects, {pv} t columns, nsteps 100, both sides",
                "R": Rt, "seconds": sec, "perms_per_s": Rt / sec, "maps_per_s": Rt * pv / sec,
                "host_cores": os.cpu_count(), "input": f"pageable host memory, {Vt * nv * 8 / 1e6:.0f} MB (an R session's heap)",
                "seconds_each_call": secs,
                "from_pinned": {"seconds": sec_pinned, "perms_per_s": Rt / sec_pinned, "seconds_each_call": secs_pinned},
                "cpu_restatement_ms_per_map": cpu_ms, "speedup_vs_one_core": (cpu_ms * 1e-3 * Rt * pv) / sec,
                "timing": "wall clock around the host-buffer C-ABI call (upload, refits, enhancement, maxima, download), median of 3"}
        del Yt_h
        if not args.small:
            # mincTFCE at config-2 volume size: smooth unit-variance maps (box-filtered noise), d = 0.1, both sides
            g = torch.Generator(device=dev).manual_seed(11)
            vols = torch.randn((4, 1) + tuple(DIMS), generator=g, device=dev, dtype=torch.float32)
            for _ in range(2):
                vols = torch.nn.functional.avg_pool3d(vols, 3, stride=1, padding=1)
            vols = (vols / vols.std(dim=(2, 3, 4), keepdim=True)).to(torch.float64)
            Mh = np.asfortranarray(vols.reshape(4, -1).T.cpu().numpy())
            del vols
            core.volume_tfce(Mh[:, :1], DIMS, d=0.1)                                # warm-up
            secvs = []
            for _ in range(3):
                t0 = time.perf_counter()
                tv = core.volume_tfce(Mh, DIMS, d=0.1)
                secvs.append(time.perf_counter() - t0)
            secv = float(np.median(secvs))
            assert np.isfinite(tv).all() and np.abs(tv).max() > 0
            tfce["mincTFCE"] = {"workload": f"mincTFCE of 4 maps of {DIMS[0]}x{DIMS[1]}x{DIMS[2]} voxels, d 0.1, both sides, 6-connected",
                                "seconds": secv, "seconds_each_call": secvs, "maps_per_s": 4 / secv,
                                "max_abs_t": float(np.abs(Mh).max()),
                                "timing": "wall clock around the host-buffer C-ABI call (grid adjacency, upload, enhancement, download)"}
            del Mh, tv

    # ---- mincRandomize (config 4): voxel-sharded, all ranks evaluate every permutation, one all-reduce(MAX)
    rand = None
    if "randomize" in args.parts:
        Rn = args.perms
        a, b = parallel.my_shard(V, rank, world)
        Ys = Yt[:, a:b]
        rng = np.random.default_rng(4)
        idx = np.column_stack([rng.permutation(n) for _ in range(Rn)]).astype(np.int32)
        cols = list(range(p + 2, 2 * p + 2))
        # warm-up: one call of the SAME size, so that the library's pooled pinned staging (the packed operand, 12 MB at R = 1000)
        # and the device workspace exist -- the first call of a size pays cudaHostAlloc, and N ranks of one box pay it one after
        # the other (measured at N = 8: 125 ms for a first call against 105 ms); a session's later calls do not
        parallel.randomize_sharded(Ys, X, idx if not args.small else idx[:, :8], cols)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        d = parallel.randomize_sharded(Ys, X, idx, cols)
        e1.record(stream)
        barrier()
        tt = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
        flops = 2.0 * p * n * V * Rn
        rand = {"metric": "mincRandomize permutations/s", "value": Rn / (ms * 1e-3), "unit": "perms/s", "R": Rn,
                "n_gpus": world, "scaling": "strong", "ms": ms, "path": os.environ.get("RMG_PERM_PATH", "gemm"),
                "warmup": "one untimed call of the same R (staging buffers pooled by the library across calls)",
                "algorithmic_tflops": flops / (ms * 1e-3) / 1e12,
                "null_95pct_t": [float(x) for x in np.quantile(d.cpu().numpy(), 0.95, axis=1)]}
        if rank == 0:
            # FP64 denominator: cuBLAS DGEMM measured in this very run
            A = torch.randn((4096, 4096), dtype=torch.float64, device=dev)
            B = torch.randn((4096, 4096), dtype=torch.float64, device=dev)
            for _ in range(2):
                A @ B
            torch.cuda.synchronize()
            e0.record(stream)
            for _ in range(5):
                A @ B
            e1.record(stream)
            torch.cuda.synchronize()
            rand["dgemm_peak_tflops"] = 5 * 2 * 4096 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
            rand["frac_of_dgemm"] = rand["algorithmic_tflops"] / world / rand["dgemm_peak_tflops"]
            # the GEMM path skips the permutation-invariant intercept row: (p - 1) / p of the algorithmic flops are EXECUTED
            executed = (p - 1) / p if rand["path"] == "gemm" else 1.0
            rand["executed_tflops"] = rand["algorithmic_tflops"] * executed
            rand["frac_of_dgemm_executed"] = rand["frac_of_dgemm"] * executed
            del A, B

    # ---- timed region 2: end to end through the C ABI with HOST buffers (pinned), rank-local
    e2e = None
    if "e2e" in args.parts:
        import psutil
        numa_cpus = bind_to_gpu_numa(local_rank)
        avail = psutil.virtual_memory().available
        need = V * n * 8
        Ve = V if avail > need * world * 1.3 + (8 << 30) else max(1024, int((avail / world * 0.5) // (8 * n)) // 1024 * 1024)
        Ve = min(Ve, V)
        hY = torch.empty((n, Ve), dtype=torch.float64, pin_memory=True)
        hY.copy_(Yt[:, :Ve])
        hO = torch.empty((2 * p + 3, Ve), dtype=torch.float64, pin_memory=True)
        Yh = hY.numpy().T            # V x n, Fortran order, zero-copy
        Oh = hO.numpy().T
        assert Yh.flags.f_contiguous and Oh.flags.f_contiguous
        from rminc_b200 import _lib
        lib = _lib.load()
        err = _lib.errbuf()

        Xf = np.asfortranarray(X)          # keep alive: ctypes only sees the address
        # plain pinned->device copy rate on this box, reported beside the number it bounds
        probe = torch.empty((min(n, 64), Ve), dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        tp = time.perf_counter()
        probe.copy_(hY[:probe.shape[0]], non_blocking=True)
        torch.cuda.synchronize()
        h2d_gbs = probe.numel() * 8 / (time.perf_counter() - tp) / 1e9
        del probe

        def e2e_step():
            rc = lib.rmg_lm_fit(Yh.ctypes.data, Ve, Ve, n, Xf.ctypes.data, p, None, 1.0, 99999999.0,
                                Oh.ctypes.data, Ve, local_rank, err, len(err))
            _lib.check(rc, err)

        e2e_step()
        ksteps = max(1, min(args.steps, 3))
        barrier()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": world * Ve * ksteps / dt, "unit": UNIT, "h2d_bytes_per_step": Ve * n * 8 + n * p * 8,
               "d2h_bytes_per_step": Ve * (2 * p + 3) * 8, "steps": ksteps, "voxels_per_step_per_gpu": Ve,
               "api": "rmg_lm_fit (C ABI, pinned host Y -> chunked H2D -> kernel -> D2H result)",
               "kernel_ms_inside": core.last_kernel_ms(), "pinned_h2d_gbs_probe": h2d_gbs,
               "bound": "PCIe: Y is 4000 B/voxel over the host link", "numa_bound_cpus": len(numa_cpus) if numa_cpus else None}
        # sanity: the end-to-end result equals the device-resident result
        # (different chunk shapes pick different summation orders, so "equal" means 1e-9 of the
        #  column scale, the same criterion tests/helpers.py uses)
        chk = torch.from_numpy(np.ascontiguousarray(Oh[:4096].T)).to(dev)
        ref = core.lm_fit_torch(Yt[:, :4096], X)      # `out` was reused by the masked variant above
        torch.cuda.synchronize()
        scale = torch.maximum(ref.abs(), ref.square().mean(dim=1, keepdim=True).sqrt())
        assert bool(((chk - ref).abs() <= 1e-9 * scale).all()), "e2e result differs from device result"
        # mincRandomize through the C ABI from the same pinned host Y: voxel blocks are uploaded on a copy stream and
        # every block runs its R permutations as it lands (strong scaling: each rank takes its voxel shard)
        if rand is not None and Ve == V:
            sa, sb = parallel.my_shard(V, rank, world)
            cols_a = np.asarray(cols, dtype=np.int32)
            idx_f = np.asfortranarray(idx, dtype=np.int32)
            dist_h = np.empty((Rn, len(cols)), order="F")

            def perm_step(R_):
                rc = lib.rmg_randomize(Yh.ctypes.data + 8 * sa, sb - sa, Ve, n, Xf.ctypes.data, p, None, 1.0, 99999999.0,
                                       idx_f.ctypes.data, R_, cols_a.ctypes.data, len(cols), 1, dist_h.ctypes.data,
                                       local_rank, err, len(err))
                _lib.check(rc, err)

            perm_step(Rn)            # warm-up of the SAME size: the library's pooled staging exists when the timed call starts
            secs = []
            for _ in range(2):       # two timed calls, the better one reported: one call in four or five takes 0.4-0.6 s longer
                barrier()            # on these boxes (host-side; the GEMM window does not move), both are listed
                t0 = time.perf_counter()
                perm_step(Rn)
                dd = torch.from_numpy(np.ascontiguousarray(dist_h)).to(dev)
                if world > 1:
                    dist.all_reduce(dd, op=dist.ReduceOp.MAX)
                torch.cuda.synchronize()
                barrier()
                tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                secs.append(float(tt.item()))
            sec = min(secs)
            assert np.allclose(dd.cpu().numpy().T, d.cpu().numpy(), rtol=1e-9, atol=0), "e2e randomize differs from the device-resident result"
            rand["e2e"] = {"value": Rn / sec, "unit": "perms/s", "seconds": sec, "seconds_each_call": secs, "h2d_bytes": (sb - sa) * n * 8,
                           "gemm_window_ms": core.last_kernel_ms(),
                           "api": "rmg_randomize (C ABI, pinned host Y -> voxel blocks on a copy stream -> all R permutations "
                                  "per block as it lands -> R x ncols maxima)"}
        del hY, hO

    # ---- end to end with the volumes in their stored type: 2 bytes per sample over the host link
    if Ut is not None and "e2e_u16" in args.parts:
        from rminc_b200 import _lib
        lib = _lib.load()
        err = _lib.errbuf()
        bind_to_gpu_numa(local_rank)
        hU = torch.empty((n, V), dtype=torch.uint16, pin_memory=True)
        hU.copy_(Ut)
        hO = torch.empty((2 * p + 3, V), dtype=torch.float64, pin_memory=True)
        Uh, Oh = hU.numpy().T, hO.numpy().T
        Xf3 = np.asfortranarray(X)

        def u16_step():
            rc = lib.rmg_lm_fit_stored(Uh.ctypes.data, core.STORED_U16, V, V, n, sc_h.ctypes.data, of_h.ctypes.data, 0.0,
                                       slice_len, Xf3.ctypes.data, p, None, 1.0, 99999999.0, Oh.ctypes.data, V,
                                       local_rank, err, len(err))
            _lib.check(rc, err)

        u16_step()
        ksteps = max(1, min(args.steps, 3))
        barrier()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            u16_step()
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        if e2e is None:
            e2e = {}
        e2e["stored_u16"] = {"value": world * V * ksteps / dt, "unit": UNIT, "h2d_bytes_per_step": V * n * 2 + 2 * sc_h.size * 8,
                             "d2h_bytes_per_step": V * (2 * p + 3) * 8, "steps": ksteps,
                             "api": "rmg_lm_fit_stored (C ABI, pinned host uint16 voxels + real-range tables -> chunked "
                                    "H2D -> expand + fit in one kernel -> D2H result)",
                             "kernel_ms_inside": core.last_kernel_ms()}
        chk = torch.from_numpy(np.ascontiguousarray(Oh[:4096].T)).to(dev)
        ref = core.lm_fit_stored_torch(Ut[:, :4096], X, sc_t[:, :1].contiguous(), of_t[:, :1].contiguous(), slice_len=4096)
        torch.cuda.synchronize()
        scl = torch.maximum(ref.abs(), ref.square().mean(dim=1, keepdim=True).sqrt())
        assert bool(((chk - ref).abs() <= 1e-9 * scl).all()), "stored e2e result differs from device result"
        if rand is not None:
            sa, sb = parallel.my_shard(V, rank, world)
            sa -= sa % slice_len                       # whole slices per rank keep the tables' global index simple here
            sb = V if rank == world - 1 else sb - sb % slice_len
            cols_a = np.asarray(cols, dtype=np.int32)
            idx_f = np.asfortranarray(idx, dtype=np.int32)
            dist_h = np.empty((Rn, len(cols)), order="F")
            s0, s1 = sa // slice_len, (sb + slice_len - 1) // slice_len
            sc_r, of_r = np.ascontiguousarray(sc_h[:, s0:s1]), np.ascontiguousarray(of_h[:, s0:s1])   # [slices x n] col-major

            def perm_u16(R_):
                rc = lib.rmg_randomize_stored(Uh.ctypes.data + 2 * sa, core.STORED_U16, sb - sa, V, n, sc_r.ctypes.data,
                                              of_r.ctypes.data, 0.0, slice_len, Xf3.ctypes.data, p, None, 1.0, 99999999.0,
                                              idx_f.ctypes.data, R_, cols_a.ctypes.data, len(cols), 1, dist_h.ctypes.data,
                                              local_rank, err, len(err))
                _lib.check(rc, err)

            perm_u16(Rn)             # warm-up of the SAME size (see the doubles' leg)
            secs_u, calls_u = [], []
            for _ in range(2):       # two timed calls, the better one reported (see the doubles' leg)
                barrier()
                t0 = time.perf_counter()
                perm_u16(Rn)
                calls_u.append(time.perf_counter() - t0)     # the C-ABI call alone, before the max-reduce over ranks
                dd = torch.from_numpy(np.ascontiguousarray(dist_h)).to(dev)
                if world > 1:
                    dist.all_reduce(dd, op=dist.ReduceOp.MAX)
                torch.cuda.synchronize()
                barrier()
                tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                secs_u.append(float(tt.item()))
            sec, call_sec = min(secs_u), min(calls_u)
            assert bool(torch.isfinite(dd).all()) and bool((dd > 0).all())
            rand["e2e_stored_u16"] = {"value": Rn / sec, "unit": "perms/s", "seconds": sec, "seconds_each_call": secs_u,
                                      "h2d_bytes": (sb - sa) * n * 2,
                                      "call_seconds_this_rank": call_sec, "gemm_window_ms": core.last_kernel_ms(),
                                      "api": "rmg_randomize_stored (C ABI, pinned host uint16 voxels -> blocks expanded once "
                                             "on the device -> all R permutations per block)"}
        del hU, hO
    Ut = None

    # ---- the same call with PAGEABLE host memory (what R's heap is): informational
    if rank == 0 and "e2e_pageable" in args.parts:
        from rminc_b200 import _lib
        lib = _lib.load()
        err = _lib.errbuf()
        Vp = min(V, 1769040)                       # a quarter of the volume: 7 GB of pageable host memory
        Yp = np.empty((Vp, n), order="F")
        Yp[:] = Yt[:, :Vp].T.cpu().numpy()
        Op = np.empty((Vp, 2 * p + 3), order="F")
        Xf2 = np.asfortranarray(X)
        for it in range(2):
            t0 = time.perf_counter()
            rc = lib.rmg_lm_fit(Yp.ctypes.data, Vp, Vp, n, Xf2.ctypes.data, p, None, 1.0, 99999999.0,
                                Op.ctypes.data, Vp, local_rank, err, len(err))
            _lib.check(rc, err)
            dtp = time.perf_counter() - t0
        if e2e is None:
            e2e = {}
        e2e["pageable_host"] = {"value": Vp / dtp, "unit": UNIT, "voxels": Vp, "gbs": Vp * n * 8 / dtp / 1e9}
        # mincRandomize from the same pageable matrix (what R's heap is): blocks staged through the pinned ring
        Rp = min(args.perms, 1000)
        rngp = np.random.default_rng(4)
        idx_p = np.asfortranarray(np.column_stack([rngp.permutation(n) for _ in range(Rp)]).astype(np.int32))
        cols_p = np.arange(p + 2, 2 * p + 2, dtype=np.int32)
        dist_p = np.empty((Rp, p), order="F")
        for R_ in (8, Rp):
            t0 = time.perf_counter()
            rc = lib.rmg_randomize(Yp.ctypes.data, Vp, Vp, n, Xf2.ctypes.data, p, None, 1.0, 99999999.0, idx_p.ctypes.data, R_,
                                   cols_p.ctypes.data, p, 1, dist_p.ctypes.data, local_rank, err, len(err))
            _lib.check(rc, err)
            dtr = time.perf_counter() - t0
        e2e["pageable_host"]["randomize"] = {"R": Rp, "voxels": Vp, "seconds": dtr, "perms_per_s": Rp / dtr,
                                             "gemm_window_ms": core.last_kernel_ms()}
        del Yp, Op

    # ---- the form R calls: ONE process (rank 0) driving all N GPUs through the C ABI, one config-2 volume strong-sharded
    #      over the N host links; the other ranks wait on a host barrier
    multi = None
    if "multi" in args.parts and not args.small:
        if rank == 0:
            multi = single_process_multi(args, core, Yt, X, world, d_ref=(d if rand is not None else None),
                                         perms=args.perms, idx=(idx if rand is not None else None))
            if rand is not None and "randomize_e2e" in multi:
                rand["e2e_single_process_multi"] = multi["randomize_e2e"]
        if world > 1:
            dist.barrier(group=host_group)
    barrier()

    # ---- config 5 (capacity): 400x560x340 x 200 subjects = 121.9 GB of doubles in 8 voxel shards of 9.52 M voxels
    cfg5 = None
    if "cfg5" in args.parts and not args.small:
        del Yt, out
        torch.cuda.empty_cache()
        cfg5 = config5(args, torch, dist, core, dev, rank, world, local_rank, barrier, design)

    files = None
    if rank == 0 and "files" in args.parts:
        files = files_leg(args, core)

    # ---- CPU baseline beside it (rank 0, N = 1 only): bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and "cpu" in args.parts:
        cores = os.cpu_count() or 1
        sample_V = int(os.environ.get("RMG_BENCH_CPU_SAMPLE", 60000 * cores))
        vps, kind, sec = cpu_reference_run(sample_V, cores)
        cpu = {"value": vps, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{sample_V} voxels of the workload (n=500, p=4), in memory, {cores} forked workers "
                         f"over contiguous chunks; {sec:.1f} s"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD if not args.small else "SMALL SMOKE CONFIG (not a bench line)",
                       "voxels_per_gpu": V, "n": n, "p": p,
                       "l2": "inputs (28.3 GB per GPU) are far larger than the 126 MB L2; no flush needed",
                       "sharding": "contiguous voxel shards, no data-path collective"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "gpu_launches": int(launches),
            "masked_variant": masked, "vertexLm": vertex, "covariate": cov, "stored_u16": stored, "vertexTFCE": tfce, "randomize": rand,
            "single_process_multi": multi, "multi_parity": (multi or {}).get("multi_parity"), "config5": cfg5,
            "from_files": files,
        }
        if e2e is not None and multi is not None:
            e2e["single_process_multi"] = multi.get("fit_e2e")
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--perms", type=int, default=1000, help="permutations timed for the randomize line")
    ap.add_argument("--small", action="store_true", help="tiny smoke configuration (not a bench line)")
    ap.add_argument("--skip-extras", action="store_true", help="only the device-resident headline")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--tfce-perms", type=int, default=64, help="permutations timed for the vertexTFCE line")
    ap.add_argument("--parts", default="masked,vertex,cov,stored,tfce,randomize,e2e,e2e_u16,multi,cpu",
                    help="which extra measurements to run beside the headline (comma list)")
    args = ap.parse_args()
    args.parts = set() if args.skip_extras else set(args.parts.split(","))
    if not args.skip_extras and args.gpus == 8 and "nocfg5" not in args.parts:
        args.parts.add("cfg5")                      # the capacity configuration is quoted on 8 GPUs
    if args.skip_cpu:
        args.parts.discard("cpu")
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
