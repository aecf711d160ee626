"""Thin numpy / torch front-ends over the C ABI (include/rminc_b200.h).

Layout convention: every matrix is column-major as in R.  A V x n column-major
matrix is the same memory as a C-contiguous array of shape (n, V); the torch entry
points therefore take Y as an (n, V) float64 CUDA tensor and return the result as a
(2p+3, V) tensor whose row c is result column c.

Result columns (src/modelling_functions.c:1159-1174):
  0 F-statistic | 1 R-squared | 2..p+1 beta | p+2..2p+1 t | 2p+2 logLik
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import RmincGpuError, SingularDesignError  # noqa: F401  (re-exported)

DEFAULT_MASK_LO = 1.0          # R/minc_voxel_statistics.R:781-787
DEFAULT_MASK_HI = 99999999.0


def mask_range(maskval=None):
    """mincLm's maskval handling (R/minc_voxel_statistics.R:781-787)."""
    if maskval is None:
        return DEFAULT_MASK_LO, DEFAULT_MASK_HI
    return float(maskval), float(maskval)


def device_count() -> int:
    return _lib.load().rmg_device_count()


def design_factor(X):
    """Host factorisation of the design (pivot, rank, Q, Rinv, cdiag).  No GPU needed."""
    X = _lib.as_f64_fortran(X)
    if X.ndim != 2:
        raise ValueError("X must be n x p")
    n, p = X.shape
    lib = _lib.load()
    pivot = np.zeros(p, dtype=np.int32)
    rank = C.c_int(0)
    info = C.c_int(0)
    Q = np.zeros((n, p), order="F")
    Rinv = np.zeros((p, p), order="F")
    cdiag = np.zeros(p)
    err = _lib.errbuf()
    rc = lib.rmg_design_factor(X.ctypes.data, n, p, pivot.ctypes.data, C.byref(rank), Q.ctypes.data,
                               Rinv.ctypes.data, cdiag.ctypes.data, C.byref(info), err, len(err))
    if rc not in (_lib.RMG_OK, _lib.RMG_ERR_SINGULAR):
        _lib.check(rc, err)
    return dict(pivot=pivot, rank=rank.value, Q=Q, Rinv=Rinv, cdiag=cdiag, info=info.value)


def _prep_host(Y, X, mask):
    X = _lib.as_f64_fortran(X)
    if X.ndim != 2:
        raise ValueError("X must be n x p")
    Y = np.asarray(Y)
    if Y.ndim != 2 or Y.shape[1] != X.shape[0]:
        raise ValueError(f"Y must be V x n with n == nrow(X); got {Y.shape} vs {X.shape}")
    if not (Y.dtype == np.float64 and Y.flags.f_contiguous):
        Y = _lib.as_f64_fortran(Y)
    m = None
    if mask is not None:
        m = np.ascontiguousarray(np.asarray(mask, dtype=np.float64).reshape(-1))
        if m.shape[0] != Y.shape[0]:
            raise ValueError("mask length must equal the number of voxels")
    return Y, X, m


def lm_fit(Y, X, mask=None, mask_lo=DEFAULT_MASK_LO, mask_hi=DEFAULT_MASK_HI, device=0, n_gpus=None):
    """mincLm's native step on host data: V x (2p+3) Fortran-ordered result."""
    Y, X, m = _prep_host(Y, X, mask)
    V, n = Y.shape
    p = X.shape[1]
    out = np.empty((V, 2 * p + 3), order="F")
    err = _lib.errbuf()
    lib = _lib.load()
    mp = m.ctypes.data if m is not None else None
    ld = max(V, 1)
    if n_gpus is None:
        rc = lib.rmg_lm_fit(Y.ctypes.data, V, ld, n, X.ctypes.data, p, mp, mask_lo, mask_hi,
                            out.ctypes.data, ld, device, err, len(err))
    else:
        rc = lib.rmg_lm_fit_multi(Y.ctypes.data, V, ld, n, X.ctypes.data, p, mp, mask_lo, mask_hi,
                                  out.ctypes.data, ld, int(n_gpus), err, len(err))
    _lib.check(rc, err)
    return out


def vertex_lm(Y, X, device=0):
    """vertexLm / anatLm's native step (vertex_lm_loop, src/vertex_modelling.c:8)."""
    Y, X, _ = _prep_host(Y, X, None)
    V, n = Y.shape
    p = X.shape[1]
    out = np.empty((V, 2 * p + 3), order="F")
    err = _lib.errbuf()
    ld = max(V, 1)
    rc = _lib.load().rmg_vertex_lm(Y.ctypes.data, V, ld, n, X.ctypes.data, p, out.ctypes.data, ld, device,
                                   err, len(err))
    _lib.check(rc, err)
    return out


def lm_fit_cov(Y, Z, Xs=None, mask=None, mask_lo=DEFAULT_MASK_LO, mask_hi=DEFAULT_MASK_HI, device=0, n_gpus=None):
    """Voxel-wise covariate design [Xs | z_v] ([1 | z_v] when Xs is None): a file term on the right-hand side
    (src/vertex_modelling.c:112-140, src/modelling_functions.c:1118-1144).  V x (2p+3), p = ncol(Xs) + 1."""
    Y = np.asarray(Y)
    if Y.ndim != 2:
        raise ValueError("Y must be V x n")
    V, n = Y.shape
    Xs_ = None if Xs is None else _lib.as_f64_fortran(Xs)
    if Xs_ is not None and (Xs_.ndim != 2 or Xs_.shape[0] != n):
        raise ValueError(f"Xs must be n x ps with n == ncol(Y); got {Xs_.shape} vs {Y.shape}")
    if not (Y.dtype == np.float64 and Y.flags.f_contiguous):
        Y = _lib.as_f64_fortran(Y)
    Z = np.asarray(Z)
    if Z.shape != Y.shape:
        raise ValueError("the voxel-wise covariate must have the shape of Y")
    if not (Z.dtype == np.float64 and Z.flags.f_contiguous):
        Z = _lib.as_f64_fortran(Z)
    m = None
    if mask is not None:
        m = np.ascontiguousarray(np.asarray(mask, dtype=np.float64).reshape(-1))
        if m.shape[0] != V:
            raise ValueError("mask length must equal the number of voxels")
    ps = 0 if Xs_ is None else Xs_.shape[1]
    p = ps + 1 if ps else 2
    out = np.empty((V, 2 * p + 3), order="F")
    err = _lib.errbuf()
    ld = max(V, 1)
    lib = _lib.load()
    fn, where = (lib.rmg_lm_fit_cov, device) if n_gpus is None else (lib.rmg_lm_fit_cov_multi, int(n_gpus))
    rc = fn(Y.ctypes.data, Z.ctypes.data, V, ld, n, Xs_.ctypes.data if ps else None, ps,
            m.ctypes.data if m is not None else None, mask_lo, mask_hi, out.ctypes.data, ld, where, err, len(err))
    _lib.check(rc, err)
    return out


def lm_fit_cov_torch(Yt, Zt, Xs=None, mask=None, mask_lo=DEFAULT_MASK_LO, mask_hi=DEFAULT_MASK_HI, out=None,
                     status=None):
    """Device-resident form: Yt, Zt (n, V) float64 CUDA tensors with the same row stride -> (2p+3, V) tensor.
    `status`: optional zeroed int32 CUDA tensor of one element, raised when a fitted covariate row is all zero."""
    import torch
    for t in (Yt, Zt):
        if not (t.is_cuda and t.dtype == torch.float64 and t.dim() == 2 and t.stride(1) == 1):
            raise ValueError("Yt / Zt must be float64 CUDA tensors of shape (n, V) with unit stride along V")
    n, V = Yt.shape
    if Zt.shape != Yt.shape or (n > 1 and Zt.stride(0) != Yt.stride(0)):
        raise ValueError("Zt must have Yt's shape and row stride")
    Xs_ = None if Xs is None else _lib.as_f64_fortran(Xs)
    ps = 0 if Xs_ is None else Xs_.shape[1]
    p = ps + 1 if ps else 2
    if out is None:
        out = torch.empty((2 * p + 3, V), dtype=torch.float64, device=Yt.device)
    mp = mask.data_ptr() if mask is not None else None
    err = _lib.errbuf()
    rc = _lib.load().rmg_lm_fit_cov_device(Yt.data_ptr(), Zt.data_ptr(), V, Yt.stride(0) if n > 1 else max(V, 1), n,
                                           Xs_.ctypes.data if ps else None, ps, mp, mask_lo, mask_hi, out.data_ptr(),
                                           out.stride(0), status.data_ptr() if status is not None else None,
                                           Yt.device.index or 0, _torch_stream(Yt), err, len(err))
    _lib.check(rc, err)
    return out


STORED_U16, STORED_F32 = 1, 2          # RMG_STORED_U16 / RMG_STORED_F32


def _stored_tables(U_dtype, V, n, scale, offset, slice_len):
    if U_dtype == np.float32:
        return STORED_F32, None, None, max(V, 1)
    if U_dtype != np.uint16:
        raise ValueError("stored volumes must be uint16 or float32")
    slice_len = max(V, 1) if slice_len is None else int(slice_len)
    nsl = (V + slice_len - 1) // slice_len if V else 0
    sc = np.asfortranarray(np.asarray(scale, dtype=np.float64).reshape(nsl, n))
    of = np.asfortranarray(np.asarray(offset, dtype=np.float64).reshape(nsl, n))
    return STORED_U16, sc, of, slice_len


def lm_fit_stored(U, X, scale=None, offset=None, voxel_min=0.0, slice_len=None, mask=None,
                  mask_lo=DEFAULT_MASK_LO, mask_hi=DEFAULT_MASK_HI, device=0, n_gpus=None):
    """mincLm's native step on volumes in their stored type: U is V x n uint16 (with [nslices x n] scale / offset
    tables from the files' real ranges) or float32.  The doubles ((u - voxel_min) * scale + offset) are formed on
    the device; the result equals lm_fit on the expanded matrix."""
    X = _lib.as_f64_fortran(X)
    U = np.asarray(U)
    if U.ndim != 2 or U.shape[1] != X.shape[0]:
        raise ValueError(f"U must be V x n with n == nrow(X); got {U.shape} vs {X.shape}")
    if not U.flags.f_contiguous:
        U = np.asfortranarray(U)
    V, n = U.shape
    p = X.shape[1]
    dtype, sc, of, slice_len = _stored_tables(U.dtype, V, n, scale, offset, slice_len)
    m = None
    if mask is not None:
        m = np.ascontiguousarray(np.asarray(mask, dtype=np.float64).reshape(-1))
        if m.shape[0] != V:
            raise ValueError("mask length must equal the number of voxels")
    out = np.empty((V, 2 * p + 3), order="F")
    err = _lib.errbuf()
    ld = max(V, 1)
    lib = _lib.load()
    fn, where = (lib.rmg_lm_fit_stored, device) if n_gpus is None else (lib.rmg_lm_fit_stored_multi, int(n_gpus))
    rc = fn(U.ctypes.data, dtype, V, ld, n, sc.ctypes.data if sc is not None else None,
            of.ctypes.data if of is not None else None, float(voxel_min), slice_len,
            X.ctypes.data, p, m.ctypes.data if m is not None else None, mask_lo, mask_hi,
            out.ctypes.data, ld, where, err, len(err))
    _lib.check(rc, err)
    return out


def lm_fit_stored_torch(Ut, X, scale=None, offset=None, voxel_min=0.0, slice_len=None, mask=None,
                        mask_lo=DEFAULT_MASK_LO, mask_hi=DEFAULT_MASK_HI, out=None, v_offset=0):
    """Device-resident form: Ut (n, V) uint16 / float32 CUDA tensor; scale / offset (n, nslices) float64 CUDA
    tensors (the column-major [nslices x n] tables).  Returns a (2p+3, V) CUDA tensor; current stream.
    v_offset: global index of Ut's first voxel when Ut is a voxel shard of a larger volume (the tables then cover
    the whole volume)."""
    import torch
    if not (Ut.is_cuda and Ut.dim() == 2 and Ut.stride(1) == 1 and Ut.dtype in (torch.uint16, torch.float32)):
        raise ValueError("Ut must be a uint16 / float32 CUDA tensor of shape (n, V) with unit stride along V")
    X = _lib.as_f64_fortran(X)
    n, V = Ut.shape
    p = X.shape[1]
    dtype = STORED_U16 if Ut.dtype == torch.uint16 else STORED_F32
    slice_len = max(V, 1) if slice_len is None else int(slice_len)
    nsl = 0
    if dtype == STORED_U16:
        for t in (scale, offset):
            if not (t is not None and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and t.dim() == 2
                    and t.shape[0] == n):
                raise ValueError("scale / offset must be contiguous float64 CUDA tensors of shape (n, nslices)")
        nsl = scale.shape[1]
        if offset.shape[1] != nsl or nsl * slice_len < v_offset + V:
            raise ValueError("scale / offset do not cover the voxels [v_offset, v_offset + V)")
    if out is None:
        out = torch.empty((2 * p + 3, V), dtype=torch.float64, device=Ut.device)
    err = _lib.errbuf()
    rc = _lib.load().rmg_lm_fit_stored_device(
        Ut.data_ptr(), dtype, V, Ut.stride(0) if n > 1 else max(V, 1), n,
        scale.data_ptr() if dtype == STORED_U16 else None, offset.data_ptr() if dtype == STORED_U16 else None,
        float(voxel_min), slice_len, nsl, int(v_offset), X.ctypes.data, p, mask.data_ptr() if mask is not None else None,
        mask_lo, mask_hi, out.data_ptr(), out.stride(0), Ut.device.index or 0, _torch_stream(Ut), err, len(err))
    _lib.check(rc, err)
    return out


STAT_T, STAT_F = 1, 2                  # RMG_STAT_T / RMG_STAT_F
FDR_LEVELS = (0.01, 0.05, 0.10, 0.15, 0.20)      # R/minc_FDR.R:370


def pvalue(stat, stat_type, df1, df2=0.0) -> float:
    """pt2(t, df1) or pf(F, df1, df2, lower.tail = FALSE) of one statistic (host; no GPU needed)."""
    return _lib.load().rmg_pvalue(float(stat), int(stat_type), float(df1), float(df2))


FDR_METHODS = {"FDR": 0, "p.adjust": 0, "BY": 1, "pFDR": 2, "fastqvalue": 2, "qvalue": 3}   # R/minc_FDR.R:429-438


def _fdr_inputs(stat, mask):
    s = np.ascontiguousarray(stat, dtype=np.float64).reshape(-1)
    V = s.shape[0]
    m = None
    if mask is not None:
        m = np.ascontiguousarray(mask, dtype=np.float64).reshape(-1)
        if m.shape[0] != V:
            raise ValueError("mask length must equal the number of voxels")
    return s, m, V


def fdr(stat, stat_type, df1, df2=0.0, mask=None, device=0, method="FDR", pi0=1.0):
    """mincFDR's native step for one column: (q-values [V], thresholds [5]) over the voxels with mask > 0.5
    (R/minc_FDR.R:254-520).  method: "FDR" / "p.adjust" (Benjamini-Hochberg), "BY", "pFDR" / "fastqvalue" (the
    reference's fast.qvalue with its qvalue_min step as written) or "qvalue" (the qvalue package's published form);
    the last two take the estimated pi0 (see fdr_pi0_fractions)."""
    if method not in FDR_METHODS:
        raise ValueError(f"unknown FDR method {method!r}")
    s, m, V = _fdr_inputs(stat, mask)
    q = np.empty(V)
    th = np.empty(5)
    err = _lib.errbuf()
    rc = _lib.load().rmg_fdr_method(s.ctypes.data, V, int(stat_type), float(df1), float(df2), m.ctypes.data if m is not None else None,
                                    q.ctypes.data, th.ctypes.data, FDR_METHODS[method], float(pi0), device, err, len(err))
    _lib.check(rc, err)
    return q, th


def fdr_pi0_fractions(stat, stat_type, df1, df2=0.0, mask=None, lambdas=None, device=0):
    """mean(p >= lambda) for every lambda over the voxels with mask > 0.5: the grid fast.qvalue smooths into pi0
    (R/minc_misc.R:439-441).  The p-values never leave the device."""
    lam = np.ascontiguousarray(np.arange(0.0, 0.91, 0.05) if lambdas is None else lambdas, dtype=np.float64).reshape(-1)
    s, m, V = _fdr_inputs(stat, mask)
    out = np.empty(lam.shape[0])
    err = _lib.errbuf()
    rc = _lib.load().rmg_fdr_pi0_fractions(s.ctypes.data, V, int(stat_type), float(df1), float(df2),
                                           m.ctypes.data if m is not None else None, lam.ctypes.data, lam.shape[0],
                                           out.ctypes.data, device, err, len(err))
    _lib.check(rc, err)
    return out


def fdr_torch(stat_t, stat_type, df1, df2=0.0, mask=None, out=None):
    """Device-resident form: stat_t a contiguous float64 CUDA vector -> (q CUDA vector, thresholds numpy[5])."""
    import torch
    if not (stat_t.is_cuda and stat_t.dtype == torch.float64 and stat_t.is_contiguous()):
        raise ValueError("stat_t must be a contiguous float64 CUDA tensor")
    V = stat_t.numel()
    if out is None:
        out = torch.empty(V, dtype=torch.float64, device=stat_t.device)
    th = np.empty(5)
    err = _lib.errbuf()
    rc = _lib.load().rmg_fdr_device(stat_t.data_ptr(), V, int(stat_type), float(df1), float(df2),
                                    mask.data_ptr() if mask is not None else None, out.data_ptr(), th.ctypes.data,
                                    stat_t.device.index or 0, _torch_stream(stat_t), err, len(err))
    _lib.check(rc, err)
    return out, th


TFCE_SIDES = {"both": 0, "positive": 1, "negative": 2}


def adjacency_csr(adjacencies):
    """The 0-based adjacency list vertexTFCE takes (one neighbour array per vertex) -> CSR (ptr int32[V+1], idx int32)."""
    if isinstance(adjacencies, tuple) and len(adjacencies) == 2:
        return (np.ascontiguousarray(adjacencies[0], dtype=np.int32), np.ascontiguousarray(adjacencies[1], dtype=np.int32))
    ptr = np.zeros(len(adjacencies) + 1, dtype=np.int32)
    ptr[1:] = np.cumsum([len(a) for a in adjacencies])
    idx = (np.concatenate([np.asarray(a, dtype=np.int32).reshape(-1) for a in adjacencies])
           if len(adjacencies) and ptr[-1] else np.zeros(0, np.int32))
    return ptr, np.ascontiguousarray(idx, dtype=np.int32)


def _tfce_weights(weights, V):
    if weights is None:
        return None
    w = np.asarray(weights, dtype=np.float64).reshape(-1)
    if w.size == 1:
        w = np.full(V, float(w[0]))
    if w.size != V:
        raise ValueError("Please supply 1 or enough weights for each element of x")      # minc_vertex_statistics.R:763
    return np.ascontiguousarray(w)


def graph_tfce(maps, adjacency, E=0.5, H=2.0, nsteps=100, side="both", weights=None, device=0):
    """vertexTFCE.numeric / .matrix native step: maps is V or V x nmaps; returns the same shape (graph_tfce,
    src/vertex_tfce.cpp:92-157)."""
    M = np.asarray(maps, dtype=np.float64)
    one = M.ndim == 1
    M = np.asfortranarray(M.reshape(-1, 1) if one else M)
    V, nmaps = M.shape
    ptr, idx = adjacency_csr(adjacency)
    if ptr.shape[0] != V + 1:
        raise ValueError("Mismatch between adjacency list and data")                        # src/vertex_tfce.cpp:97
    w = _tfce_weights(weights, V)
    out = np.empty((V, nmaps), order="F")
    err = _lib.errbuf()
    rc = _lib.load().rmg_graph_tfce(M.ctypes.data, V, nmaps, ptr.ctypes.data, idx.ctypes.data if idx.size else None,
                                    w.ctypes.data if w is not None else None, float(E), float(H), int(nsteps),
                                    TFCE_SIDES[side], out.ctypes.data, device, err, len(err))
    _lib.check(rc, err)
    return out[:, 0] if one else out


def randomize_tfce(Y, X, idx, cols, adjacency, two_sided=True, E=0.5, H=2.0, nsteps=100, side="both", weights=None,
                   device=0, n_gpus=None):
    """vertexTFCE.vertexLm's permutation loop: R x len(cols) maxima of the enhanced statistic maps."""
    Y, X, _ = _prep_host(Y, X, None)
    V, n = Y.shape
    p = X.shape[1]
    idx, cols = _prep_perm(idx, n, cols)
    ptr, aidx = adjacency_csr(adjacency)
    if ptr.shape[0] != V + 1:
        raise ValueError("Mismatch between adjacency list and data")
    w = _tfce_weights(weights, V)
    R = idx.shape[1]
    dist = np.empty((R, len(cols)), order="F")
    err = _lib.errbuf()
    lib = _lib.load()
    fn, where = (lib.rmg_randomize_tfce, device) if n_gpus is None else (lib.rmg_randomize_tfce_multi, int(n_gpus))
    rc = fn(Y.ctypes.data, V, max(V, 1), n, X.ctypes.data, p, idx.ctypes.data, R, cols.ctypes.data, len(cols),
            int(bool(two_sided)), ptr.ctypes.data, aidx.ctypes.data if aidx.size else None,
            w.ctypes.data if w is not None else None, float(E), float(H), int(nsteps), TFCE_SIDES[side],
            dist.ctypes.data, where, err, len(err))
    _lib.check(rc, err)
    return dist


def mesh_area(vertex_matrix, triangle_matrix):
    """RMINC:::mesh_area (src/vertex_tfce.cpp:221-257): per-vertex areas of a triangle mesh -- vertexTFCE's default weights
    for a bic_obj surface.  vertex_matrix 3 x nv, triangle_matrix 3 x nt with 1-based vertex numbers.  Host only."""
    vm = np.asfortranarray(vertex_matrix, dtype=np.float64)
    tm = np.asfortranarray(triangle_matrix, dtype=np.float64)
    if vm.ndim != 2 or vm.shape[0] != 3 or tm.ndim != 2 or tm.shape[0] != 3:
        raise ValueError("vertex_matrix must be 3 x nvertices and triangle_matrix 3 x ntriangles")
    areas = np.empty(vm.shape[1])
    err = _lib.errbuf()
    _lib.check(_lib.load().rmg_mesh_area(vm.ctypes.data, vm.shape[1], tm.ctypes.data, tm.shape[1], areas.ctypes.data, err, len(err)), err)
    return areas


def neighbour_list(x, y, z, n=6):
    """RMINC:::neighbour_list (src/vertex_tfce.cpp:178-214) as a CSR pair (ptr, idx) that vertexTFCE / graph_tfce take:
    the adjacency of an x*y*z grid, vertex = k*y*x + j*x + i.  Host only."""
    import ctypes as C
    lib = _lib.load()
    err = _lib.errbuf()
    nnz = C.c_int64(0)
    _lib.check(lib.rmg_neighbour_list(int(x), int(y), int(z), int(n), None, None, 0, C.byref(nnz), err, len(err)), err)
    ptr = np.empty(int(x) * int(y) * int(z) + 1, dtype=np.int32)
    idx = np.empty(max(nnz.value, 1), dtype=np.int32)
    _lib.check(lib.rmg_neighbour_list(int(x), int(y), int(z), int(n), ptr.ctypes.data, idx.ctypes.data, nnz.value,
                                      C.byref(nnz), err, len(err)), err)
    return ptr, idx[:nnz.value]


def _dims3(dims, V=None):
    d = np.ascontiguousarray(dims, dtype=np.int64).reshape(-1)
    if d.size != 3 or (d < 1).any():
        raise ValueError("dims must be three positive sizes (file dimension order)")
    if V is not None and int(d.prod()) != V:
        raise ValueError(f"dims {tuple(int(x) for x in d)} do not match {V} voxels")
    return d


def volume_tfce(maps, dims, d=0.1, E=0.5, H=2.0, side="both", connectivity=6, voxel_volume=1.0, device=0):
    """mincTFCE's native step: maps is V or V x nmaps (V = prod(dims), file dimension order); returns the same shape.
    The tree's graph TFCE on the voxel grid with the height step d (rmg_volume_tfce)."""
    M = np.asarray(maps, dtype=np.float64)
    one = M.ndim == 1
    M = np.asfortranarray(M.reshape(-1, 1) if one else M)
    V, nmaps = M.shape
    dd = _dims3(dims, V)
    out = np.empty((V, nmaps), order="F")
    err = _lib.errbuf()
    rc = _lib.load().rmg_volume_tfce(M.ctypes.data, dd.ctypes.data, nmaps, int(connectivity), float(voxel_volume), float(d),
                                     float(E), float(H), TFCE_SIDES[side], out.ctypes.data, device, err, len(err))
    _lib.check(rc, err)
    return out[:, 0] if one else out


def randomize_volume_tfce(Y, X, idx, cols, dims, two_sided=True, d=0.1, E=0.5, H=2.0, side="both", connectivity=6,
                          voxel_volume=1.0, mask=None, mask_lo=DEFAULT_MASK_LO, mask_hi=DEFAULT_MASK_HI, n_gpus=None):
    """mincTFCE.mincLm's permutation loop: R x len(cols) maxima of the enhanced statistic volumes."""
    Y, X, m = _prep_host(Y, X, mask)
    V, n = Y.shape
    p = X.shape[1]
    idx, cols = _prep_perm(idx, n, cols)
    dd = _dims3(dims, V)
    R = idx.shape[1]
    dist = np.empty((R, len(cols)), order="F")
    err = _lib.errbuf()
    rc = _lib.load().rmg_randomize_volume_tfce(
        Y.ctypes.data, V, max(V, 1), n, X.ctypes.data, p, m.ctypes.data if m is not None else None, mask_lo, mask_hi,
        idx.ctypes.data, R, cols.ctypes.data, len(cols), int(bool(two_sided)), dd.ctypes.data, int(connectivity),
        float(voxel_volume), float(d), float(E), float(H), TFCE_SIDES[side], dist.ctypes.data,
        int(n_gpus) if n_gpus is not None else 1, err, len(err))
    _lib.check(rc, err)
    return dist


def vertex_wlm(Y, X, weights, device=0):
    """anatLm(weights=)'s native step (vertex_wlm_loop, src/weighted_least_squares.c:168)."""
    Y, X, _ = _prep_host(Y, X, None)
    V, n = Y.shape
    p = X.shape[1]
    w = np.ascontiguousarray(weights, dtype=np.float64)
    if w.shape != (n,):
        raise ValueError("Incorrect number of weights supplied, need one per observation")   # minc_anatomy.R:1133
    out = np.empty((V, 2 * p + 3), order="F")
    err = _lib.errbuf()
    ld = max(V, 1)
    rc = _lib.load().rmg_vertex_wlm(Y.ctypes.data, V, ld, n, X.ctypes.data, p, w.ctypes.data, out.ctypes.data, ld,
                                    device, err, len(err))
    _lib.check(rc, err)
    return out


def anova_fit(Y, X, asgn, mask=None, mask_lo=DEFAULT_MASK_LO, mask_hi=DEFAULT_MASK_HI, device=0):
    """mincAnova / vertexAnova / anatAnova native step: V x nterms sequential F-statistics."""
    Y, X, m = _prep_host(Y, X, mask)
    V, n = Y.shape
    p = X.shape[1]
    asgn = np.ascontiguousarray(asgn, dtype=np.int32)
    if asgn.shape != (p,):
        raise ValueError("asgn must have one entry per design column")
    nterms = int(asgn.max()) if p else 0
    out = np.empty((V, nterms), order="F")
    err = _lib.errbuf()
    ld = max(V, 1)
    rc = _lib.load().rmg_anova_fit(Y.ctypes.data, V, ld, n, X.ctypes.data, p, asgn.ctypes.data,
                                   m.ctypes.data if m is not None else None, mask_lo, mask_hi,
                                   out.ctypes.data, ld, device, err, len(err))
    _lib.check(rc, err)
    return out


def anova_fit_torch(Yt, X, asgn, mask=None, mask_lo=DEFAULT_MASK_LO, mask_hi=DEFAULT_MASK_HI, out=None):
    """Device-resident anova: Yt (n, V) CUDA tensor -> (nterms, V) CUDA tensor, current stream."""
    import torch
    X = _lib.as_f64_fortran(X)
    n, V = Yt.shape
    p = X.shape[1]
    asgn = np.ascontiguousarray(asgn, dtype=np.int32)
    nterms = int(asgn.max()) if p else 0
    if out is None:
        out = torch.empty((nterms, V), dtype=torch.float64, device=Yt.device)
    mp = mask.data_ptr() if mask is not None else None
    err = _lib.errbuf()
    rc = _lib.load().rmg_anova_fit_device(Yt.data_ptr(), V, Yt.stride(0) if n > 1 else max(V, 1), n, X.ctypes.data, p,
                                          asgn.ctypes.data, mp, mask_lo, mask_hi, out.data_ptr(),
                                          out.stride(0) if nterms > 1 else max(V, 1), Yt.device.index or 0,
                                          _torch_stream(Yt), err, len(err))
    _lib.check(rc, err)
    return out


def _prep_perm(idx, n, cols):
    idx = np.asfortranarray(idx, dtype=np.int32)
    if idx.ndim != 2 or idx.shape[0] != n:
        raise ValueError("idx must be n x R")
    if idx.size and (idx.min() < 0 or idx.max() >= n):
        raise ValueError("idx entries must be in 0..n-1")
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    return idx, cols


def randomize(Y, X, idx, cols, two_sided=True, mask=None, mask_lo=DEFAULT_MASK_LO, mask_hi=DEFAULT_MASK_HI,
              device=0, n_gpus=None):
    """mincRandomize's loop on host data: R x len(cols) matrix of per-permutation maxima."""
    Y, X, m = _prep_host(Y, X, mask)
    V, n = Y.shape
    p = X.shape[1]
    idx, cols = _prep_perm(idx, n, cols)
    if cols.size and (cols.min() < 0 or cols.max() >= 2 * p + 3):
        raise ValueError("cols entries must be in 0..2p+2")
    R = idx.shape[1]
    dist = np.empty((R, len(cols)), order="F")
    err = _lib.errbuf()
    lib = _lib.load()
    mp = m.ctypes.data if m is not None else None
    ld = max(V, 1)
    if n_gpus is None:
        rc = lib.rmg_randomize(Y.ctypes.data, V, ld, n, X.ctypes.data, p, mp, mask_lo, mask_hi, idx.ctypes.data, R,
                               cols.ctypes.data, len(cols), int(bool(two_sided)), dist.ctypes.data, device,
                               err, len(err))
    else:
        rc = lib.rmg_randomize_multi(Y.ctypes.data, V, ld, n, X.ctypes.data, p, mp, mask_lo, mask_hi,
                                     idx.ctypes.data, R, cols.ctypes.data, len(cols), int(bool(two_sided)),
                                     dist.ctypes.data, int(n_gpus), err, len(err))
    _lib.check(rc, err)
    return dist


def randomize_cov(Y, Z, Xs, idx, cols, two_sided=True, mask=None, mask_lo=DEFAULT_MASK_LO, mask_hi=DEFAULT_MASK_HI,
                  device=0, n_gpus=None):
    """mincRandomize's loop for a model with a voxel-wise covariate: permutation r fits y_v on [Xs | z_v][idx[:, r], :]
    (the covariate files are re-sampled with the static predictors, R/minc_voxel_statistics.R:1468-1477).
    R x len(cols) maxima; cols index the 2p+3 columns of the p = ncol(Xs) + 1 column result."""
    Y = _lib.as_f64_fortran(np.asarray(Y))
    Z = _lib.as_f64_fortran(np.asarray(Z))
    if Y.ndim != 2 or Z.shape != Y.shape:
        raise ValueError("Y must be V x n and the voxel-wise covariate must have its shape")
    V, n = Y.shape
    Xs_ = None if Xs is None else _lib.as_f64_fortran(Xs)
    if Xs_ is not None and (Xs_.ndim != 2 or Xs_.shape[0] != n):
        raise ValueError(f"Xs must be n x ps with n == ncol(Y); got {Xs_.shape} vs {Y.shape}")
    ps = 0 if Xs_ is None else Xs_.shape[1]
    p = ps + 1 if ps else 2
    m = None
    if mask is not None:
        m = np.ascontiguousarray(np.asarray(mask, dtype=np.float64).reshape(-1))
        if m.shape[0] != V:
            raise ValueError("mask length must equal the number of voxels")
    idx, cols = _prep_perm(idx, n, cols)
    if cols.size and (cols.min() < 0 or cols.max() >= 2 * p + 3):
        raise ValueError("cols entries must be in 0..2p+2")
    R = idx.shape[1]
    dist = np.empty((R, len(cols)), order="F")
    err = _lib.errbuf()
    lib = _lib.load()
    fn, where = (lib.rmg_randomize_cov, device) if n_gpus is None else (lib.rmg_randomize_cov_multi, int(n_gpus))
    rc = fn(Y.ctypes.data, Z.ctypes.data, V, max(V, 1), n, Xs_.ctypes.data if ps else None, ps,
            m.ctypes.data if m is not None else None, mask_lo, mask_hi, idx.ctypes.data, R, cols.ctypes.data, len(cols),
            int(bool(two_sided)), dist.ctypes.data, where, err, len(err))
    _lib.check(rc, err)
    return dist


def randomize_stored(U, X, idx, cols, scale=None, offset=None, voxel_min=0.0, slice_len=None, two_sided=True, mask=None,
                     mask_lo=DEFAULT_MASK_LO, mask_hi=DEFAULT_MASK_HI, device=0, n_gpus=None):
    """mincRandomize's loop on volumes in their stored type (U, scale, offset, voxel_min, slice_len as lm_fit_stored):
    the samples are expanded once on the device; the result equals randomize() on the expanded doubles."""
    X = _lib.as_f64_fortran(X)
    U = np.asarray(U)
    if U.ndim != 2 or U.shape[1] != X.shape[0]:
        raise ValueError(f"U must be V x n with n == nrow(X); got {U.shape} vs {X.shape}")
    if not U.flags.f_contiguous:
        U = np.asfortranarray(U)
    V, n = U.shape
    p = X.shape[1]
    dtype, sc, of, slice_len = _stored_tables(U.dtype, V, n, scale, offset, slice_len)
    m = None
    if mask is not None:
        m = np.ascontiguousarray(np.asarray(mask, dtype=np.float64).reshape(-1))
        if m.shape[0] != V:
            raise ValueError("mask length must equal the number of voxels")
    idx, cols = _prep_perm(idx, n, cols)
    if cols.size and (cols.min() < 0 or cols.max() >= 2 * p + 3):
        raise ValueError("cols entries must be in 0..2p+2")
    R = idx.shape[1]
    dist = np.empty((R, len(cols)), order="F")
    err = _lib.errbuf()
    lib = _lib.load()
    fn, where = (lib.rmg_randomize_stored, device) if n_gpus is None else (lib.rmg_randomize_stored_multi, int(n_gpus))
    rc = fn(U.ctypes.data, dtype, V, max(V, 1), n, sc.ctypes.data if sc is not None else None,
            of.ctypes.data if of is not None else None, float(voxel_min), slice_len, X.ctypes.data, p,
            m.ctypes.data if m is not None else None, mask_lo, mask_hi, idx.ctypes.data, R, cols.ctypes.data, len(cols),
            int(bool(two_sided)), dist.ctypes.data, where, err, len(err))
    _lib.check(rc, err)
    return dist


# ------------------------------------------------------------------ the resident dataset
class Dataset:
    """Y uploaded once, sharded over `n_gpus` devices of the box inside this process, resident until close().

    mincLm keeps `data` and the call so mincRandomize / mincAnova / mincFDR can be run on the same files later
    (R/minc_voxel_statistics.R:911-912, :1457-1477) and each re-reads every file; here the staged matrix stays in
    HBM (rmg_dataset_*).  Y: V x n doubles, or uint16 / float32 stored samples with scale / offset tables
    (lm_fit_stored's arguments).  mask / mask_lo / mask_hi as lm_fit.  n_gpus None: every visible device."""

    def __init__(self, Y, mask=None, mask_lo=DEFAULT_MASK_LO, mask_hi=DEFAULT_MASK_HI, n_gpus=None, scale=None, offset=None,
                 voxel_min=0.0, slice_len=None):
        Y = np.asarray(Y)
        if Y.ndim != 2:
            raise ValueError("Y must be V x n")
        stored = Y.dtype in (np.uint16, np.float32)
        if not stored and not (Y.dtype == np.float64 and Y.flags.f_contiguous):
            Y = _lib.as_f64_fortran(Y)
        if stored and not Y.flags.f_contiguous:
            Y = np.asfortranarray(Y)
        V, n = Y.shape
        m = None
        if mask is not None:
            m = np.ascontiguousarray(np.asarray(mask, dtype=np.float64).reshape(-1))
            if m.shape[0] != V:
                raise ValueError("mask length must equal the number of voxels")
        self._handle = C.c_void_p()
        self.V, self.n = V, n
        self.out_cols = 0
        err = _lib.errbuf()
        lib = _lib.load()
        g = 0 if n_gpus is None else int(n_gpus)
        mp = m.ctypes.data if m is not None else None
        if stored:
            dtype, sc, of, slice_len = _stored_tables(Y.dtype, V, n, scale, offset, slice_len)
            rc = lib.rmg_dataset_create_stored(Y.ctypes.data, dtype, V, max(V, 1), n, sc.ctypes.data if sc is not None else None,
                                               of.ctypes.data if of is not None else None, float(voxel_min), slice_len, mp,
                                               mask_lo, mask_hi, g, C.byref(self._handle), err, len(err))
        else:
            rc = lib.rmg_dataset_create(Y.ctypes.data, V, max(V, 1), n, mp, mask_lo, mask_hi, g, C.byref(self._handle), err, len(err))
        _lib.check(rc, err)
        self._keep = (Y, m)          # pinned sources may still be read by the queued upload until sync()
        self.n_gpus = lib.rmg_dataset_gpus(self._handle)

    @classmethod
    def from_minc2(cls, filenames, mask=None, mask_lo=DEFAULT_MASK_LO, mask_hi=DEFAULT_MASK_HI, n_gpus=None, n_threads=None):
        """The resident dataset of a study of MINC 2 files: decoded by the library's reader on the host cores into pinned
        memory (uint16 / float32 studies in their stored type, expanded once on the devices; other types as doubles) and
        uploaded in voxel shards.  `mask` may be a MINC 2 file, an array or None.  What mincLm(files) followed by
        mincRandomize / mincAnova / mincFDR re-reads from disk every time (R/minc_voxel_statistics.R:911-912, 1457-1477)."""
        from . import minc2
        U, lo, hi, vr, info = minc2.read_table(list(filenames), n_threads=n_threads, pinned="auto")
        if isinstance(mask, (str, os.PathLike)):
            with minc2.Minc2Volume(mask) as mv:
                mask = mv.real()
        if U.dtype == np.uint16:
            return cls(U, mask=mask, mask_lo=mask_lo, mask_hi=mask_hi, n_gpus=n_gpus, scale=(hi - lo) / (vr[1] - vr[0]), offset=lo,
                       voxel_min=vr[0], slice_len=info.slice_len if info.nslices > 1 else max(U.shape[0], 1))
        return cls(U, mask=mask, mask_lo=mask_lo, mask_hi=mask_hi, n_gpus=n_gpus)

    def _h(self):
        if not self._handle:
            raise RmincGpuError(_lib.RMG_ERR_INVALID, "the dataset has been closed")
        return self._handle

    def sync(self):
        err = _lib.errbuf()
        _lib.check(_lib.load().rmg_dataset_sync(self._h(), err, len(err)), err)
        self._keep = None
        return self

    def lm_fit(self, X, want_result=True):
        """mincLm's native step on the resident matrix: V x (2p+3) (None when want_result is False; the result also
        stays resident for fdr())."""
        X = _lib.as_f64_fortran(X)
        if X.ndim != 2 or X.shape[0] != self.n:
            raise ValueError("X must be n x p with n == ncol(Y)")
        p = X.shape[1]
        out = np.empty((self.V, 2 * p + 3), order="F") if want_result else None
        err = _lib.errbuf()
        rc = _lib.load().rmg_dataset_lm_fit(self._h(), X.ctypes.data, p, out.ctypes.data if want_result else None, max(self.V, 1),
                                            err, len(err))
        _lib.check(rc, err)
        self.out_cols = 2 * p + 3
        return out

    def anova(self, X, asgn):
        X = _lib.as_f64_fortran(X)
        p = X.shape[1]
        asgn = np.ascontiguousarray(asgn, dtype=np.int32)
        if asgn.shape != (p,):
            raise ValueError("asgn must have one entry per design column")
        nterms = int(asgn.max()) if p else 0
        out = np.empty((self.V, nterms), order="F")
        err = _lib.errbuf()
        rc = _lib.load().rmg_dataset_anova(self._h(), X.ctypes.data, p, asgn.ctypes.data, out.ctypes.data, max(self.V, 1), err, len(err))
        _lib.check(rc, err)
        self.out_cols = nterms
        return out

    def randomize(self, X, idx, cols, two_sided=True):
        X = _lib.as_f64_fortran(X)
        p = X.shape[1]
        idx, cols = _prep_perm(idx, self.n, cols)
        if cols.size and (cols.min() < 0 or cols.max() >= 2 * p + 3):
            raise ValueError("cols entries must be in 0..2p+2")
        R = idx.shape[1]
        dist = np.empty((R, len(cols)), order="F")
        err = _lib.errbuf()
        rc = _lib.load().rmg_dataset_randomize(self._h(), X.ctypes.data, p, idx.ctypes.data, R, cols.ctypes.data, len(cols),
                                               int(bool(two_sided)), dist.ctypes.data, err, len(err))
        _lib.check(rc, err)
        return dist

    def fdr(self, column, stat_type, df1, df2=0.0, method="FDR", pi0=1.0):
        """mincFDR's native step on one column of the resident result of the last fit: (q-values [V], thresholds [5]).
        method / pi0 as core.fdr (R/minc_FDR.R:429-438)."""
        q = np.empty(self.V)
        th = np.empty(5)
        err = _lib.errbuf()
        rc = _lib.load().rmg_dataset_fdr_method(self._h(), int(column), int(stat_type), float(df1), float(df2), q.ctypes.data,
                                                th.ctypes.data, int(FDR_METHODS[method]), float(pi0), err, len(err))
        _lib.check(rc, err)
        return q, th

    def fdr_pi0_fractions(self, column, stat_type, df1, df2=0.0, lambdas=None):
        """mean(p >= lambda) of the column's p-values (fast.qvalue's pi0 grid, R/minc_misc.R:439-441)."""
        lam = np.ascontiguousarray(np.arange(0.0, 0.91, 0.05) if lambdas is None else lambdas, dtype=np.float64)
        fr = np.empty(lam.size)
        err = _lib.errbuf()
        rc = _lib.load().rmg_dataset_fdr_pi0_fractions(self._h(), int(column), int(stat_type), float(df1), float(df2), lam.ctypes.data,
                                                       int(lam.size), fr.ctypes.data, err, len(err))
        _lib.check(rc, err)
        return fr

    def close(self):
        if self._handle:
            _lib.load().rmg_dataset_free(self._handle)
            self._handle = C.c_void_p()
            self._keep = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass


# ------------------------------------------------------------------ device-resident (torch) forms
def _torch_stream(t):
    import torch
    return torch.cuda.current_stream(t.device).cuda_stream


def lm_fit_torch(Yt, X, mask=None, mask_lo=DEFAULT_MASK_LO, mask_hi=DEFAULT_MASK_HI, out=None):
    """Fit with Y resident in HBM.  Yt: (n, V) float64 CUDA tensor (row stride = ldY).
    Returns a (2p+3, V) CUDA tensor.  Enqueued on torch's current stream; does not synchronise."""
    import torch
    if not (Yt.is_cuda and Yt.dtype == torch.float64 and Yt.dim() == 2 and Yt.stride(1) == 1):
        raise ValueError("Yt must be a float64 CUDA tensor of shape (n, V) with unit stride along V")
    X = _lib.as_f64_fortran(X)
    n, V = Yt.shape
    if X.shape[0] != n:
        raise ValueError("nrow(X) must equal Yt.shape[0]")
    p = X.shape[1]
    ldY = Yt.stride(0) if n > 1 else max(V, 1)
    if out is None:
        out = torch.empty((2 * p + 3, V), dtype=torch.float64, device=Yt.device)
    if not (out.is_cuda and out.dtype == torch.float64 and out.shape == (2 * p + 3, V) and out.stride(1) == 1):
        raise ValueError("out must be a (2p+3, V) float64 CUDA tensor")
    mp = None
    if mask is not None:
        if not (mask.is_cuda and mask.dtype == torch.float64 and mask.numel() == V and mask.is_contiguous()):
            raise ValueError("mask must be a contiguous float64 CUDA tensor of V elements")
        mp = mask.data_ptr()
    err = _lib.errbuf()
    rc = _lib.load().rmg_lm_fit_device(Yt.data_ptr(), V, ldY, n, X.ctypes.data, p, mp, mask_lo, mask_hi,
                                       out.data_ptr(), out.stride(0), Yt.device.index or 0, _torch_stream(Yt),
                                       err, len(err))
    _lib.check(rc, err)
    return out


def randomize_torch(Yt, X, idx, cols, two_sided=True, mask=None, mask_lo=DEFAULT_MASK_LO,
                    mask_hi=DEFAULT_MASK_HI, dist=None):
    """Permutation maxima over THIS device's voxels.  Returns an (ncols, R) CUDA tensor (the
    column-major R x ncols matrix); combine shards with an element-wise max."""
    import torch
    if not (Yt.is_cuda and Yt.dtype == torch.float64 and Yt.dim() == 2 and Yt.stride(1) == 1):
        raise ValueError("Yt must be a float64 CUDA tensor of shape (n, V) with unit stride along V")
    X = _lib.as_f64_fortran(X)
    n, V = Yt.shape
    p = X.shape[1]
    idx, cols = _prep_perm(idx, n, cols)
    R = idx.shape[1]
    ldY = Yt.stride(0) if n > 1 else max(V, 1)
    if dist is None:
        dist = torch.empty((len(cols), R), dtype=torch.float64, device=Yt.device)
    mp = None
    if mask is not None:
        if not (mask.is_cuda and mask.dtype == torch.float64 and mask.numel() == V and mask.is_contiguous()):
            raise ValueError("mask must be a contiguous float64 CUDA tensor of V elements")
        mp = mask.data_ptr()
    err = _lib.errbuf()
    rc = _lib.load().rmg_randomize_device(Yt.data_ptr(), V, ldY, n, X.ctypes.data, p, mp, mask_lo, mask_hi,
                                          idx.ctypes.data, R, cols.ctypes.data, len(cols), int(bool(two_sided)),
                                          dist.data_ptr(), Yt.device.index or 0, _torch_stream(Yt), err, len(err))
    _lib.check(rc, err)
    return dist


def last_kernel_ms() -> float:
    return _lib.load().rmg_last_kernel_ms()


def launch_count() -> int:
    return _lib.load().rmg_launch_count()


def last_fit_variant() -> str:
    return {0: "none", 1: "lm_fit_tma_kernel", 2: "lm_fit_ldg_kernel", 3: "lm_cov_kernel", 4: "lm_fit_packed_kernel",
            5: "lm_fit_packed_tma_kernel"}[_lib.load().rmg_last_fit_variant()]
