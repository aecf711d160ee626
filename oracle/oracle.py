"""ctypes binding for the CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  Nothing under rminc_b200/ does.

liboracle.so    : oracle/rminc_oracle.c, the plain-C restatement of
                  src/modelling_functions.c:448-576 (voxel_lm), :1012-1175
                  (minc2_model, "lm"), src/vertex_modelling.c:104-158 and
                  R/minc_voxel_statistics.R:1465-1497 (mincRandomize).
_ref/librminc_ref.so : the reference's own voxel_lm / vertex_lm_loop /
                  minc2_model object code (built by `make ref` where
                  /root/reference exists; see oracle/Makefile).

Parity unpinned (third-party code absent from the image, restated from published algorithms):
expand_stored (libminc's stored-voxel -> real conversion) and fdr / p_adjust_bh (R's nmath pt / pf,
stats::p.adjust); see the header of rminc_oracle.c and DESIGN.md section 3.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
_REF = os.path.join(_HERE, "_ref", "librminc_ref.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


def build(ref: bool = True) -> None:
    """Compile liboracle.so (and _ref/ when /root/reference is present)."""
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    if ref and os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-C", _HERE, "-s", "ref"], check=True)


def _load(path):
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    return C.CDLL(path)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build(ref=False)
        _lib = _load(_LIB)
        _lib.orc_voxel_lm.restype = C.c_int
        _lib.orc_vertex_lm_loop.restype = C.c_int
        _lib.orc_minc2_model_lm.restype = C.c_int
        _lib.orc_randomize.restype = C.c_int
        _lib.orc_design_probe.restype = C.c_int
        _lib.orc_dnrm2.restype = C.c_double
        _lib.orc_anova_loop.restype = C.c_int
        _lib.orc_vertex_wlm_loop.restype = C.c_int
        _lib.orc_lm_loop_cov.restype = C.c_int
    return _lib


def have_ref() -> bool:
    return os.path.exists(_REF)


def ref():
    global _ref
    if _ref is None:
        _ref = _load(_REF)
        for f in ("ref_voxel_lm", "ref_vertex_lm_loop", "ref_minc2_model_lm", "ref_vertex_anova_loop",
                  "ref_per_voxel_anova", "ref_vertex_wlm_loop", "ref_vertex_lm_loop_cov"):
            getattr(_ref, f).restype = C.c_int
    return _ref


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _F(a):
    return np.asfortranarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(_dp)


class OracleError(RuntimeError):
    """The reference raises an R error (dpotri info != 0, modelling_functions.c:551-557)."""


def voxel_lm(y, X, use_ref=False):
    """One voxel.  Returns dict(coef, F, t, R2, logLik, pivot, rank)."""
    y = _f64(y)
    X = _F(X)
    n, p = X.shape
    coef = np.zeros(p)
    stats = np.zeros(p + 3)
    pivot = np.zeros(p, dtype=np.int32)
    if use_ref:
        err = C.create_string_buffer(512)
        rc = ref().ref_voxel_lm(_p(y), _p(X), C.c_int(n), C.c_int(p), _p(coef), _p(stats),
                                pivot.ctypes.data_as(_ip), err, C.c_size_t(512))
        if rc:
            raise OracleError(err.value.decode())
        rank = None
    else:
        rk = C.c_int(0)
        info = lib().orc_voxel_lm(_p(y), _p(X), C.c_int(n), C.c_int(p), _p(coef), _p(stats),
                                  pivot.ctypes.data_as(_ip), C.byref(rk))
        if info:
            raise OracleError(f"element ({info}, {info}) is zero, so the inverse cannot be computed")
        rank = rk.value
    return dict(coef=coef, F=stats[0], t=stats[1:p + 1].copy(), R2=stats[p + 1], logLik=stats[p + 2],
                pivot=pivot, rank=rank)


def vertex_lm_loop(Y, X, use_ref=False):
    """Y: V x n (any layout; copied to column-major).  Returns V x (2p+3), Fortran order."""
    Y = _F(Y)
    X = _F(X)
    V, n = Y.shape
    p = X.shape[1]
    out = np.zeros((V, 2 * p + 3), order="F")
    if use_ref:
        err = C.create_string_buffer(512)
        rc = ref().ref_vertex_lm_loop(_p(Y), C.c_int64(V), C.c_int(n), _p(X), C.c_int(p), _p(out),
                                      err, C.c_size_t(512))
        if rc:
            raise OracleError(err.value.decode())
    else:
        info = lib().orc_vertex_lm_loop(_p(Y), C.c_int64(V), C.c_int64(V), C.c_int(n), _p(X), C.c_int(p),
                                        _p(out), C.c_int64(V))
        if info:
            raise OracleError(f"element ({info}, {info}) is zero, so the inverse cannot be computed")
    return out


def vertex_lm_loop_threads(Y, X, threads):
    """The reference's `parallel = c("local", N)` sharding (contiguous voxel chunks,
    R/minc_parallel.R:613-622) with N host threads (ctypes drops the GIL).  Used by
    bench.py's cpu_baseline only.  Y must be V x n Fortran order."""
    assert Y.flags.f_contiguous
    X = _F(X)
    V, n = Y.shape
    p = X.shape[1]
    out = np.zeros((V, 2 * p + 3), order="F")
    L = lib()
    bounds = np.linspace(0, V, threads + 1).astype(np.int64)

    def work(g):
        a, b = int(bounds[g]), int(bounds[g + 1])
        if b <= a:
            return 0
        yoff = C.cast(C.c_void_p(Y.ctypes.data + 8 * a), _dp)
        ooff = C.cast(C.c_void_p(out.ctypes.data + 8 * a), _dp)
        return L.orc_vertex_lm_loop(yoff, C.c_int64(b - a), C.c_int64(V), C.c_int(n), _p(X), C.c_int(p),
                                    ooff, C.c_int64(V))

    with ThreadPoolExecutor(max_workers=threads) as ex:
        rcs = list(ex.map(work, range(threads)))
    if any(rcs):
        raise OracleError("dpotri info != 0")
    return out


def minc2_model_lm(Y, sizes, X, mask=None, lo=1.0, hi=99999999.0, use_ref=False, fill=0.0):
    """Y: V x n with V = prod(sizes) rows in file (C) order.  mask: length-V doubles or None."""
    Y = _F(Y)
    X = _F(X)
    V, n = Y.shape
    p = X.shape[1]
    sz = (C.c_int64 * 3)(*[int(s) for s in sizes])
    assert int(np.prod(sizes)) == V
    out = np.zeros((V, 2 * p + 3), order="F")
    m = None if mask is None else _f64(np.asarray(mask).reshape(-1))
    mp = _p(m) if m is not None else None
    if use_ref:
        err = C.create_string_buffer(512)
        rc = ref().ref_minc2_model_lm(_p(Y), sz, C.c_int(n), _p(X), C.c_int(p), mp, C.c_double(lo),
                                      C.c_double(hi), _p(out), C.c_double(fill), err, C.c_size_t(512))
        if rc:
            raise OracleError(err.value.decode())
    else:
        info = lib().orc_minc2_model_lm(_p(Y), sz, C.c_int(n), _p(X), C.c_int(p), mp, C.c_double(lo),
                                        C.c_double(hi), _p(out))
        if info:
            raise OracleError(f"element ({info}, {info}) is zero, so the inverse cannot be computed")
    return out


def randomize(Y, X, idx, cols, two_sided=True, mask=None, lo=1.0, hi=99999999.0):
    """idx: n x R (0-based).  cols: 0-based output columns.  Returns R x len(cols)."""
    Y = _F(Y)
    X = _F(X)
    V, n = Y.shape
    p = X.shape[1]
    idx = np.asfortranarray(idx, dtype=np.int32)
    assert idx.shape[0] == n
    R = idx.shape[1]
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    dist = np.zeros((R, len(cols)), order="F")
    scratch = np.zeros(V * (2 * p + 3))
    m = None if mask is None else _f64(np.asarray(mask).reshape(-1))
    info = lib().orc_randomize(_p(Y), C.c_int64(V), C.c_int(n), _p(X), C.c_int(p),
                               _p(m) if m is not None else None, C.c_double(lo), C.c_double(hi),
                               idx.ctypes.data_as(_ip), C.c_int(R), cols.ctypes.data_as(_ip),
                               C.c_int(len(cols)), C.c_int(1 if two_sided else 0), _p(dist), _p(scratch))
    if info:
        raise OracleError(f"element ({info}, {info}) is zero, so the inverse cannot be computed")
    return dist


def randomize_cov(Y, Z, Xs, idx, cols, two_sided=True, mask=None, lo=1.0, hi=99999999.0, use_ref=False):
    """mincRandomize_core on a model with a voxel-wise covariate, literally (R/minc_voxel_statistics.R:1465-1497): per
    permutation shuf_data[, pred_cols] <- data[sample(n), pred_cols] re-samples the static predictors AND the column of
    covariate files together, the whole model is refitted (lm_loop_cov: the per-voxel design [Xs | z_v], restated from /
    equal to the reference's vertex_lm_loop with data_right), non-finite -> 0, abs when two-sided, column maximum."""
    Y = _F(Y)
    Z = _F(Z)
    idx = np.asarray(idx)
    out = np.empty((idx.shape[1], len(cols)))
    for r in range(idx.shape[1]):
        ix = idx[:, r]
        fit = lm_loop_cov(Y, np.asfortranarray(Z[:, ix]), None if Xs is None else np.asarray(Xs)[ix], mask=mask, lo=lo, hi=hi,
                          use_ref=use_ref)
        sel = fit[:, list(cols)].copy()
        sel[~np.isfinite(sel)] = 0.0
        out[r] = np.abs(sel).max(axis=0) if two_sided else sel.max(axis=0)
    return out


def design_probe(X):
    """pivot, rank, R (p x p upper), c = diag((R'R)^-1), dpotri info."""
    X = _F(X)
    n, p = X.shape
    pivot = np.zeros(p, dtype=np.int32)
    rank = C.c_int(0)
    Rm = np.zeros((p, p), order="F")
    c = np.zeros(p)
    info = lib().orc_design_probe(_p(X), C.c_int(n), C.c_int(p), pivot.ctypes.data_as(_ip), C.byref(rank),
                                  _p(Rm), _p(c))
    return dict(pivot=pivot, rank=rank.value, R=Rm, c=c, info=info)


def anova(Y, X, asgn, mask=None, lo=1.0, hi=99999999.0, use_ref=False, sizes=None):
    """mincAnova / vertexAnova / anatAnova native step: V x nterms F-statistics (voxel_anova,
    src/minc_anova.c:4-80).  asgn = model.matrix's "assign" attribute.  use_ref: the reference's
    vertex_anova_loop (no mask) or per_voxel_anova (mask or sizes given)."""
    Y = _F(Y)
    X = _F(X)
    V, n = Y.shape
    p = X.shape[1]
    asgn = np.ascontiguousarray(asgn, dtype=np.int32)
    nterms = int(asgn.max()) if len(asgn) else 0
    out = np.zeros((V, nterms), order="F")
    m = None if mask is None else _f64(np.asarray(mask).reshape(-1))
    if use_ref:
        err = C.create_string_buffer(512)
        if m is None and sizes is None:
            rc = ref().ref_vertex_anova_loop(_p(Y), C.c_int64(V), C.c_int(n), _p(X), C.c_int(p),
                                             asgn.ctypes.data_as(_ip), _p(out), err, C.c_size_t(512))
        else:
            sz = (C.c_int64 * 3)(*[int(s) for s in (sizes or (1, 1, V))])
            rc = ref().ref_per_voxel_anova(_p(Y), sz, C.c_int(n), _p(X), C.c_int(p), asgn.ctypes.data_as(_ip),
                                           _p(m) if m is not None else None, C.c_double(lo), C.c_double(hi), _p(out),
                                           err, C.c_size_t(512))
        if rc:
            raise OracleError(err.value.decode())
        return out
    info = lib().orc_anova_loop(_p(Y), C.c_int64(V), C.c_int64(V), C.c_int(n), _p(X), C.c_int(p),
                                asgn.ctypes.data_as(_ip), _p(m) if m is not None else None, C.c_double(lo),
                                C.c_double(hi), _p(out), C.c_int64(max(V, 1)))
    if info:
        raise OracleError(f"element ({info}, {info}) is zero, so the inverse cannot be computed")
    return out


def vertex_wlm_loop(Y, X, w, use_ref=False):
    """anatLm(weights=): vertex_wlm_loop / voxel_wlm (src/weighted_least_squares.c).  V x (2p+3)."""
    Y = _F(Y)
    X = _F(X)
    w = _f64(w)
    V, n = Y.shape
    p = X.shape[1]
    out = np.zeros((V, 2 * p + 3), order="F")
    if use_ref:
        err = C.create_string_buffer(512)
        rc = ref().ref_vertex_wlm_loop(_p(Y), C.c_int64(V), C.c_int(n), _p(X), C.c_int(p), _p(w), _p(out), err,
                                       C.c_size_t(512))
        if rc:
            raise OracleError(err.value.decode())
        return out
    info = lib().orc_vertex_wlm_loop(_p(Y), C.c_int64(V), C.c_int64(V), C.c_int(n), _p(X), C.c_int(p), _p(w), _p(out),
                                     C.c_int64(max(V, 1)))
    if info:
        raise OracleError(f"element ({info}, {info}) is zero, so the inverse cannot be computed")
    return out


def lm_loop_cov(Y, Z, Xs=None, mask=None, lo=1.0, hi=99999999.0, use_ref=False):
    """Voxel-wise covariate design [Xs | z_v] (or [1 | z_v] when Xs is None): a file term on the
    right-hand side of the formula (src/vertex_modelling.c:112-140).  V x (2p+3), p = ncol(Xs)+1."""
    Y = _F(Y)
    Z = _F(Z)
    V, n = Y.shape
    assert Z.shape == (V, n)
    Xs_ = None if Xs is None else _F(Xs)
    ps = 0 if Xs_ is None else Xs_.shape[1]
    p = ps + 1 if Xs_ is not None else 2
    out = np.zeros((V, 2 * p + 3), order="F")
    m = None if mask is None else _f64(np.asarray(mask).reshape(-1))
    if use_ref:
        assert m is None
        err = C.create_string_buffer(512)
        rc = ref().ref_vertex_lm_loop_cov(_p(Y), _p(Z), C.c_int64(V), C.c_int(n), _p(Xs_) if Xs_ is not None else None,
                                          C.c_int(ps), _p(out), err, C.c_size_t(512))
        if rc:
            raise OracleError(err.value.decode())
        return out
    info = lib().orc_lm_loop_cov(_p(Y), _p(Z), C.c_int64(V), C.c_int64(V), C.c_int(n),
                                 _p(Xs_) if Xs_ is not None else None, C.c_int(ps), _p(m) if m is not None else None,
                                 C.c_double(lo), C.c_double(hi), _p(out), C.c_int64(max(V, 1)))
    if info:
        raise OracleError(f"element ({info}, {info}) is zero, so the inverse cannot be computed")
    return out


def expand_stored(U, scale=None, offset=None, voxel_min=0.0, slice_len=None):
    """The doubles libminc's miget_real_value_hyperslab(MI_TYPE_DOUBLE) hands to voxel_lm for stored voxels U
    (V x n, uint16 with [nslices x n] scale / offset tables, or float32): see orc_expand_u16."""
    U = np.asfortranarray(U)
    V, n = U.shape
    out = np.empty((V, n), order="F")
    if U.dtype == np.float32:
        lib().orc_expand_f32(U.ctypes.data_as(C.c_void_p), C.c_int64(V), C.c_int64(max(V, 1)), C.c_int(n), _p(out),
                             C.c_int64(max(V, 1)))
        return out
    assert U.dtype == np.uint16
    slice_len = V if slice_len is None else int(slice_len)
    nsl = (V + slice_len - 1) // slice_len
    sc = _F(np.asarray(scale, dtype=np.float64).reshape(nsl, n))
    of = _F(np.asarray(offset, dtype=np.float64).reshape(nsl, n))
    lib().orc_expand_u16(U.ctypes.data_as(C.c_void_p), C.c_int64(V), C.c_int64(max(V, 1)), C.c_int(n), _p(sc), _p(of),
                         C.c_double(voxel_min), C.c_int64(slice_len), _p(out), C.c_int64(max(V, 1)))
    return out


# ------------------------------------------------------------------ mincFDR (R/minc_FDR.R:254-520)
FDR_LEVELS = (0.01, 0.05, 0.10, 0.15, 0.20)


def p_adjust_bh(p):
    """R's p.adjust(p, "fdr") (stats::p.adjust, method "BH"): NA p-values are dropped, n is the number of the
    remaining ones (the default n = length(p) is a promise evaluated after the NAs were removed), and
    q = pmin(1, cummin(n / i * p[o]))[ro] with o = order(p, decreasing = TRUE)."""
    p = np.asarray(p, dtype=np.float64)
    q = np.full(p.shape, np.nan)
    ok = ~np.isnan(p)
    pp = p[ok]
    n = pp.size
    if n == 0:
        return q
    o = np.argsort(-pp, kind="stable")
    i = np.arange(n, 0, -1)
    adj = np.minimum(1.0, np.minimum.accumulate(n / i * pp[o]))
    out = np.empty(n)
    out[o] = adj
    q[ok] = out
    return q


def p_adjust_by(p):
    """R's p.adjust(p, "BY"): q <- sum(1 / (1L:n)); pmin(1, cummin(q * n / i * p[o]))[ro].  R's sum() accumulates in long
    double and rounds once."""
    p = np.asarray(p, dtype=np.float64)
    q = np.full(p.shape, np.nan)
    ok = ~np.isnan(p)
    pp = p[ok]
    n = pp.size
    if n == 0:
        return q
    c = float(np.sum(1.0 / np.arange(1, n + 1, dtype=np.float64), dtype=np.longdouble))
    o = np.argsort(-pp, kind="stable")
    i = np.arange(n, 0, -1)
    adj = np.minimum(1.0, np.minimum.accumulate(c * n / i * pp[o]))
    out = np.empty(n)
    out[o] = adj
    q[ok] = out
    return q


def qvalue_min(qvalue, u, use_ref=False):
    """The reference's qvalue_min (src/modelling_functions.c:15-46), restated with its indices as written: u is R's
    1-based order(p).  The last-ranked element is tested through qvalue[u[m-1]] (1-based index used 0-based = the element
    AFTER it in the vector; one past the end when it is the last element: read as "not above 1" here, and the caller of
    the object code below pads the vector with a 0 there); every other element takes the smaller of its own and the
    next-ranked RAW value, or 1 when both exceed 1."""
    qv = np.ascontiguousarray(qvalue, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.int32)
    m = qv.shape[0]
    out = np.zeros(m)
    if use_ref:
        padded = np.concatenate([qv, [0.0]])
        length = np.array([m], dtype=np.int32)
        ref().qvalue_min(padded.ctypes.data_as(C.c_void_p), u.ctypes.data_as(C.c_void_p), length.ctypes.data_as(C.c_void_p),
                         out.ctypes.data_as(C.c_void_p))
        return out
    last = m - 1
    probe = qv[u[last]] if u[last] < m else 0.0
    out[u[last] - 1] = 1.0 if probe > 1 else qv[u[last] - 1]
    for i in range(last - 1, -1, -1):
        a, b = qv[u[i] - 1], qv[u[i + 1] - 1]
        out[u[i] - 1] = 1.0 if (a > 1 and b > 1) else (a if a < b else b)
    return out


def fast_qvalue(p, pi0, use_ref=False):
    """fast.qvalue's q-values for a given pi0 (R/minc_misc.R:503-531): v = qvalue.rank(p) = number of p' <= p,
    qvalue = pi0 * m * p / v, then qvalue_min over u = order(p)."""
    p = np.asarray(p, dtype=np.float64)
    m = p.shape[0]
    u = np.argsort(p, kind="stable") + 1
    v = np.searchsorted(np.sort(p), p, side="right")
    return qvalue_min(pi0 * m * p / v, u, use_ref=use_ref)


def pi0_fractions(p, lambdas):
    """mean(p >= lambda[i]) (R/minc_misc.R:439-441)."""
    p = np.asarray(p, dtype=np.float64)
    return np.array([np.mean(p >= l) for l in lambdas])


def qvalue_package(p, pi0):
    """qvalue::qvalue's q-values for a given pi0 (third party; its published form since 2.0):
    i <- m:1; o <- order(p, decreasing = TRUE); ro <- order(o); pi0 * pmin(1, cummin(p[o] * m / i))[ro]."""
    p = np.asarray(p, dtype=np.float64)
    m = p.shape[0]
    o = np.argsort(-p, kind="stable")
    i = np.arange(m, 0, -1)
    out = np.empty(m)
    out[o] = pi0 * np.minimum(1.0, np.minimum.accumulate(p[o] * m / i))
    return out


def fdr(stat, stat_type, df1, df2=None, mask=None, method="FDR", pi0=1.0):
    """mincFDR.mincMultiDim for one column (R/minc_FDR.R:385-505): p-values (pt2 = 2 pt(-|t|, df) for "t",
    pf(F, df1, df2, lower.tail = FALSE) for "F"; scipy's t / f distributions stand in for R's nmath), q-values by
    `method` over mask > 0.5 (:429-438), output 1 outside the mask, thresholds qt(max p / 2, df, lower = FALSE) /
    qf(max p, ...) of the largest p among q <= level (:444-475)."""
    from scipy import stats as S
    s = np.asarray(stat, dtype=np.float64).reshape(-1)
    sel = np.ones(s.shape, bool) if mask is None else (np.asarray(mask).reshape(-1) > 0.5)
    if stat_type == "t":
        p = 2.0 * S.t.sf(np.abs(s[sel]), df1)
    else:
        p = S.f.sf(s[sel], df1, df2)
    p = np.where(np.isnan(s[sel]), np.nan, p)
    if method in ("FDR", "p.adjust"):
        q = p_adjust_bh(p)
    elif method == "BY":
        q = p_adjust_by(p)
    elif method in ("pFDR", "fastqvalue"):
        q = fast_qvalue(p, pi0)
    elif method == "qvalue":
        q = qvalue_package(p, pi0)
    else:
        raise ValueError(method)
    out = np.ones(s.shape)
    out[sel] = q
    th = np.full(len(FDR_LEVELS), np.nan)
    for j, lev in enumerate(FDR_LEVELS):
        acc = p[(q <= lev) & ~np.isnan(p)]
        if acc.size:
            th[j] = S.t.isf(acc.max() / 2.0, df1) if stat_type == "t" else S.f.isf(acc.max(), df1, df2)
    return out, th, p


# ------------------------------------------------------------------ graph TFCE (src/vertex_tfce.cpp:92-157)
_REF_TFCE = os.path.join(_HERE, "_ref", "librminc_ref_tfce.so")
_ref_tfce = None
TFCE_SIDES = {"both": 0, "positive": 1, "negative": 2}


def have_ref_tfce() -> bool:
    return os.path.exists(_REF_TFCE)


def adjacency_csr(adjacencies):
    """0-based adjacency list (list of neighbour arrays, as vertexTFCE takes it) -> CSR (ptr int32[V+1], idx int32)."""
    ptr = np.zeros(len(adjacencies) + 1, dtype=np.int32)
    ptr[1:] = np.cumsum([len(a) for a in adjacencies])
    idx = np.concatenate([np.asarray(a, dtype=np.int32) for a in adjacencies]) if len(adjacencies) else np.zeros(0, np.int32)
    return ptr, np.ascontiguousarray(idx, dtype=np.int32)


def graph_tfce(x, adj_ptr, adj_idx, E=0.5, H=2.0, nsteps=100, side="both", weights=None, use_ref=False):
    """vertexTFCE.numeric's native step: graph_tfce_wqu / graph_tfce on one map (R/minc_vertex_statistics.R:788-796)."""
    global _ref_tfce
    x = _f64(np.asarray(x).reshape(-1))
    V = x.shape[0]
    w = np.ones(V) if weights is None else _f64(np.broadcast_to(np.asarray(weights, dtype=np.float64), (V,)).copy())
    ptr = np.ascontiguousarray(adj_ptr, dtype=np.int32)
    idx = np.ascontiguousarray(adj_idx, dtype=np.int32)
    out = np.zeros(V)
    args = (_p(x), C.c_int64(V), ptr.ctypes.data_as(C.c_void_p), idx.ctypes.data_as(C.c_void_p), _p(w), C.c_double(E),
            C.c_double(H), C.c_int(nsteps), C.c_int(TFCE_SIDES[side]), _p(out))
    if use_ref:
        if _ref_tfce is None:
            _ref_tfce = _load(_REF_TFCE)
            _ref_tfce.ref_graph_tfce.restype = C.c_int
        err = C.create_string_buffer(256)
        if _ref_tfce.ref_graph_tfce(*args, err, C.c_size_t(256)):
            raise OracleError(err.value.decode())
        return out
    lib().orc_graph_tfce(*args)
    return out


def neighbour_list(x, y, z, n, use_ref=False):
    """RMINC:::neighbour_list (src/vertex_tfce.cpp:178-214), literally: vertex = k*y*x + j*x + i (i fastest), neighbours
    within Manhattan distance 1 / 2 / 3 for n = 6 / 18 / 26.  Pure-Python loops: small grids only.
    use_ref: run the reference's own compiled function instead (oracle/_ref/librminc_ref_tfce.so)."""
    global _ref_tfce
    if use_ref:
        if _ref_tfce is None:
            _ref_tfce = _load(_REF_TFCE)
            _ref_tfce.ref_graph_tfce.restype = C.c_int
        ptr = np.zeros(x * y * z + 1, dtype=np.int32)
        nnz = C.c_int64(0)
        err = C.create_string_buffer(256)
        pp = ptr.ctypes.data_as(C.c_void_p)
        if _ref_tfce.ref_neighbour_list(C.c_int(x), C.c_int(y), C.c_int(z), C.c_int(n), pp, None, C.byref(nnz), err, C.c_size_t(256)):
            raise OracleError(err.value.decode())
        idx = np.zeros(max(nnz.value, 1), dtype=np.int32)
        _ref_tfce.ref_neighbour_list(C.c_int(x), C.c_int(y), C.c_int(z), C.c_int(n), pp, idx.ctypes.data_as(C.c_void_p),
                                     C.byref(nnz), err, C.c_size_t(256))
        return [idx[ptr[v]:ptr[v + 1]].tolist() for v in range(x * y * z)]
    maxdist = {6: 1, 18: 2, 26: 3}.get(n)
    if maxdist is None:
        raise OracleError("Nearest neighbour connectivity must be 6,18, or 26")
    nlist = [None] * (x * y * z)
    for i in range(x):
        for j in range(y):
            for k in range(z):
                nebs = []
                for oi in (-1, 0, 1):
                    if 0 <= i + oi < x:
                        for oj in (-1, 0, 1):
                            if 0 <= j + oj < y:
                                for ok in (-1, 0, 1):
                                    if 0 <= k + ok < z and 0 < abs(oi) + abs(oj) + abs(ok) <= maxdist:
                                        nebs.append((k + ok) * y * x + (j + oj) * x + (i + oi))
                nlist[k * y * x + j * x + i] = nebs
    return nlist


def mesh_area(vertex_matrix, triangle_matrix, use_ref=False):
    """RMINC:::mesh_area (src/vertex_tfce.cpp:221-257), literally -- including its first cross-product component
    d1y*d2z - d1z*d2x.  use_ref: the reference's own compiled function (oracle/_ref/librminc_ref_tfce.so)."""
    global _ref_tfce
    vm = np.asfortranarray(vertex_matrix, dtype=np.float64)
    tm = np.asfortranarray(triangle_matrix, dtype=np.float64)
    nv, nt = vm.shape[1], tm.shape[1]
    areas = np.zeros(nv)
    if use_ref:
        if _ref_tfce is None:
            _ref_tfce = _load(_REF_TFCE)
            _ref_tfce.ref_graph_tfce.restype = C.c_int
        _ref_tfce.ref_mesh_area(vm.ctypes.data_as(C.c_void_p), C.c_int64(nv), tm.ctypes.data_as(C.c_void_p), C.c_int64(nt),
                                areas.ctypes.data_as(C.c_void_p))
        return areas
    for i in range(nt):
        a, b, c = (int(t) - 1 for t in tm[:, i])
        d1, d2 = vm[:, b] - vm[:, a], vm[:, c] - vm[:, a]
        c0 = d1[1] * d2[2] - d1[2] * d2[0]
        c1 = d1[2] * d2[0] - d1[0] * d2[2]
        c2 = d1[0] * d2[1] - d1[1] * d2[0]
        area = .5 * np.sqrt(c0 * c0 + c1 * c1 + c2 * c2)
        for v in (a, b, c):
            areas[v] += (1 / 3.0) * area
    return areas


def volume_tfce(x, dims, d=0.1, E=0.5, H=2.0, side="both", connectivity=6, voxel_volume=1.0):
    """mincTFCE on one volume (flat, file dimension order).  PARITY UNPINNED against the reference's mincTFCE, which
    shells out to the external `TFCE` program of minc-stuffs (R/minc_voxel_statistics.R:1225-1237; not in the reference
    tree).  What the reference does pin is that vertexTFCE on RMINC:::neighbour_list(.., 6) with weights = the voxel
    volume agrees with it (tests/testthat/test_mincTFCE.R:153-170), so this is the tree's graph_tfce
    (src/vertex_tfce.cpp:92-157) on the voxel grid with the height step d that mincTFCE passes -- restated
    INDEPENDENTLY of the union-find formulation, with scipy's connected-component labelling: for t = d, 2d, ... <= max,
    every voxel of a cluster of {x >= t} receives t^H * (voxels * voxel_volume)^E * d."""
    from scipy import ndimage
    vol = _f64(np.asarray(x).reshape(-1)).reshape(tuple(int(v) for v in dims))
    structure = ndimage.generate_binary_structure(3, {6: 1, 18: 2, 26: 3}[connectivity])

    def positive(v):
        out = np.zeros(v.shape)
        hmax = v.max()
        if not np.isfinite(hmax) or hmax < d:
            return out
        for j in range(1, int(np.floor(hmax / d)) + 1):
            t = j * d
            lab, nlab = ndimage.label(v >= t, structure=structure)
            if nlab == 0:
                continue
            sizes = np.bincount(lab.reshape(-1), minlength=nlab + 1).astype(np.float64) * voxel_volume
            contrib = (t ** H) * (sizes ** E) * d
            contrib[0] = 0.0
            out += contrib[lab]
        return out

    if side == "positive":
        res = positive(vol)
    elif side == "negative":
        res = -positive(-vol)
    else:
        res = positive(vol) - positive(-vol)
    return res.reshape(-1)


def randomize_volume_tfce(Y, X, idx, cols, dims, two_sided=True, mask=None, lo=1.0, hi=99999999.0, **tf):
    """mincTFCE.mincLm's loop, literally (R/minc_voxel_statistics.R:1465-1497 with post_proc = mincTFCE.matrix): per
    permutation refit, non-finite -> 0, enhance every requested column, abs when two-sided, column maximum."""
    Y = np.asfortranarray(Y, dtype=np.float64)
    idx = np.asarray(idx)
    out = np.empty((idx.shape[1], len(cols)))
    for r in range(idx.shape[1]):
        fit = minc2_model_lm(Y, (1, 1, Y.shape[0]), np.asarray(X)[idx[:, r]], mask=mask, lo=lo, hi=hi)
        for c, col in enumerate(cols):
            m = fit[:, col].copy()
            m[~np.isfinite(m)] = 0.0
            e = volume_tfce(m, dims, **tf)
            out[r, c] = np.abs(e).max() if two_sided else e.max()
    return out

