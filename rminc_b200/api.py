"""Host-side mirror of RMINC's R interface for the voxel-wise OLS path.

R is not available in this image, so the layer RMINC implements in R
(R/minc_voxel_statistics.R, R/minc_vertex_statistics.R, R/minc_anatomy.R) is mirrored here in
Python with the same names, argument meaning, result layout, attributes and error behaviour,
so the parity tests read like the reference's own testthat files.  The R-side binding a
maintainer would add is in rminc_b200/R/ and INTEGRATION.md.

    mincLm(formula, data, subset, mask, maskval, parallel)   R/minc_voxel_statistics.R:751-950
    vertexLm(formula, data, subset, column)                  R/minc_vertex_statistics.R:365-446
    anatLm(formula, data, anat, subset, weights)             R/minc_anatomy.R:1048-1192
    mincAnova / vertexAnova / anatAnova                      R/minc_voxel_statistics.R:495-693,
                                                             R/minc_vertex_statistics.R:301-344
    mincRandomize(x, R, alternative, replace, parallel, columns)   :1326-1378, :1441-1531
    thresholds(x, probs)                                     :1577-1581
    mincFDR / vertexFDR / anatFDR                            R/minc_FDR.R:254-573
    vertexTFCE (numeric / matrix / vertexLm randomisation)   R/minc_vertex_statistics.R:731-948

Everything numerical goes through the C ABI (rminc_b200.core); there is no CPU path.
MINC I/O (libminc) is outside the hot path: volumes are staged by a `reader` callable
(default: numpy .npy files or in-memory arrays), exactly where mincTable / mincGetVolume sit in
the reference (R/minc_interface.R:1031-1077).
"""
from __future__ import annotations

import os
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from . import core
from .formula import FormulaError, model_matrix, parse_rhs, split_formula


# ----------------------------------------------------------------------------- result objects
class MincMatrix(np.ndarray):
    """A numeric matrix with R-style attributes: `colnames`, `rownames`, `attrs`, `r_class`.

    Mirrors the objects mincLm / vertexLm / anatLm return: a V x (2p+3) matrix of class
    c("mincLm","mincMultiDim","matrix") etc. with attributes likeVolume, filenames, model,
    data, call, stat-type and df (R/minc_voxel_statistics.R:908-944)."""

    def __new__(cls, array, colnames=None, rownames=None, attrs=None, r_class=()):
        obj = np.asarray(array).view(cls)
        obj.colnames = list(colnames) if colnames is not None else None
        obj.rownames = list(rownames) if rownames is not None else None
        obj.attrs = dict(attrs or {})
        obj.r_class = tuple(r_class)
        return obj

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self.colnames = getattr(obj, "colnames", None)
        self.rownames = getattr(obj, "rownames", None)
        self.attrs = getattr(obj, "attrs", {})
        self.r_class = getattr(obj, "r_class", ())

    def col(self, name: str) -> np.ndarray:
        return np.asarray(self)[:, self.colnames.index(name)]


def _stat_type(p: int) -> List[str]:
    return ["F", "R-squared"] + ["beta"] * p + ["t"] * p + ["logLik"]      # :921-927


def _df_list(n: int, p: int) -> list:
    fdf1, fdf2 = p - 1, n - p                                              # :929-935
    return [[fdf1, fdf2]] + [fdf2] * p


def _colnames(rows: Sequence[str]) -> List[str]:
    return ["F-statistic", "R-squared"] + [f"beta-{r}" for r in rows] + [f"tvalue-{r}" for r in rows] + ["logLik"]


def _subset_rows(data, subset):
    n = len(data)
    if subset is None:
        return list(range(n))
    s = np.asarray(subset)
    if s.dtype == bool:
        if len(s) != n:
            raise ValueError("logical subset must have one entry per row of data")
        return [i for i in range(n) if s[i]]
    return [int(i) for i in s]


def _column(data, name, rows):
    c = data[name]
    return [c.iloc[i] if hasattr(c, "iloc") else c[i] for i in rows]


# ----------------------------------------------------------------------------- staging (I/O side)
def default_reader(item):
    """Volume loader used where the reference calls mincGetVolume: in-memory arrays pass through, '.npy' paths are
    np.load'ed, MINC 2 files go through the library's own reader (rminc_b200/minc2.py: uint16 / float32 volumes come back
    as StoredVolume so that they travel in their stored type, other types as the real doubles).  Arrays are flattened in
    file (C) order."""
    if isinstance(item, np.ndarray):
        return np.asarray(item, dtype=np.float64).reshape(-1)
    if isinstance(item, (str, os.PathLike)):
        path = os.fspath(item)
        if not os.path.exists(path):
            raise FileNotFoundError(f"Error opening input file: {path}.")          # minc2_model :809
        if path.endswith(".npy"):
            return np.load(path).astype(np.float64, copy=False).reshape(-1)
        from . import minc2
        if minc2.is_minc2(path):
            return minc2.read_volume(path)
        raise ValueError(f"no reader for '{path}': not a MINC 2 (HDF5) file or .npy array; pass reader= (MINC 1 files: mincconvert -2)")
    raise TypeError(f"cannot stage a volume from {type(item).__name__}")


class StoredVolume:
    """A volume as the file stores it: `voxels` (uint16 or float32, any shape, file order) plus, for uint16, the
    per-slice real range of the slowest dimension as MINC keeps it (image-min / image-max) and the valid range.
    A reader that returns StoredVolume objects instead of double arrays makes mincLm ship 2 (or 4) bytes per
    sample to the GPU, where the doubles libminc's miget_real_value_hyperslab would produce are formed in
    registers (src/modelling_functions.c:1037-1055; rmg_lm_fit_stored)."""

    def __init__(self, voxels, image_min=None, image_max=None, valid_range=(0.0, 65535.0)):
        v = np.asarray(voxels)
        if v.dtype not in (np.uint16, np.float32):
            raise TypeError("StoredVolume voxels must be uint16 or float32")
        self.voxels = v
        self.nslices = 1
        self.scale = self.offset = None
        self.valid_min = float(valid_range[0])
        if v.dtype == np.uint16:
            lo = np.atleast_1d(np.asarray(image_min, dtype=np.float64))
            hi = np.atleast_1d(np.asarray(image_max, dtype=np.float64))
            if lo.shape != hi.shape or lo.ndim != 1 or (lo.size != 1 and lo.size != v.shape[0]):
                raise ValueError("image_min / image_max: one value, or one per slice of the slowest dimension")
            self.nslices = lo.size
            self.scale = (hi - lo) / (float(valid_range[1]) - float(valid_range[0]))
            self.offset = lo

    def real(self) -> np.ndarray:
        """The doubles the CPU reader would return (what mincGetVolume yields)."""
        v = self.voxels
        if v.dtype == np.float32:
            return v.astype(np.float64).reshape(-1)
        shape = (self.nslices,) + (1,) * (v.ndim - 1) if self.nslices > 1 else (1,) * v.ndim
        return ((v.astype(np.float64) - self.valid_min) * self.scale.reshape(shape) + self.offset.reshape(shape)).reshape(-1)


def mincTableStored(vols: Sequence["StoredVolume"]):
    """U (V x n, stored type), scale / offset ([nslices x n]), voxel_min, slice_len for rmg_lm_fit_stored."""
    first = vols[0]
    V, n = first.voxels.size, len(vols)
    U = np.empty((V, n), dtype=first.voxels.dtype, order="F")
    nsl = max(v.nslices for v in vols)
    scale = offset = None
    if first.voxels.dtype == np.uint16:
        scale, offset = np.empty((nsl, n)), np.empty((nsl, n))
    for j, v in enumerate(vols):
        if v.voxels.size != V or v.voxels.dtype != U.dtype:
            raise ValueError("all volumes must have the same number of voxels and the same stored type")
        if v.valid_min != first.valid_min:
            raise ValueError("all volumes must share one valid range")
        U[:, j] = v.voxels.reshape(-1)
        if scale is not None:
            scale[:, j] = np.broadcast_to(v.scale, (nsl,)) if v.nslices in (1, nsl) else np.nan
            offset[:, j] = np.broadcast_to(v.offset, (nsl,)) if v.nslices in (1, nsl) else np.nan
            if v.nslices not in (1, nsl):
                raise ValueError("volumes disagree on the number of slices")
    return U, scale, offset, first.valid_min, (V // nsl if nsl > 1 else max(V, 1))


def _expand_stored(U, scale, offset, voxel_min, slice_len):
    """The doubles the CPU reader would produce from mincTableStored's pieces (StoredVolume.real() for all volumes at
    once): ((double)u - valid_min) * scale[slice, j] + offset[slice, j], each operation rounded once."""
    if U.dtype == np.float32:
        return np.asfortranarray(U, dtype=np.float64)
    sl = np.minimum(np.arange(U.shape[0]) // int(slice_len), scale.shape[0] - 1)
    return np.asfortranarray((U.astype(np.float64) - float(voxel_min)) * scale[sl, :] + offset[sl, :])


def mincTable(filenames: Sequence, mask=None, reader: Callable = default_reader) -> np.ndarray:
    """V x n matrix, one column per file (R/minc_interface.R:1031-1077), Fortran order."""
    def real(item):
        v = reader(item)
        return v.real() if isinstance(v, StoredVolume) else v

    first = real(filenames[0])
    Y = np.empty((first.shape[0], len(filenames)), order="F")
    Y[:, 0] = first
    for j, f in enumerate(filenames[1:], start=1):
        v = real(f)
        if v.shape[0] != first.shape[0]:
            raise ValueError("all volumes must have the same number of voxels")
        Y[:, j] = v
    return Y


def _minc2_stage(filenames):
    """A study of MINC 2 files decoded in parallel by the library's reader (rmg_minc2_read_table) straight into the
    column-major matrix the entry points take: ("stored", U, scale, offset, voxel_min, slice_len) for uint16 / float32 files,
    ("real", Y) otherwise; None when the names are not all MINC 2 files."""
    from . import minc2
    if not all(isinstance(f, (str, os.PathLike)) for f in filenames):
        return None
    for f in filenames:
        if not os.path.exists(os.fspath(f)):
            raise FileNotFoundError(f"Error opening input file: {os.fspath(f)}.")          # minc2_model :809
    if not all(minc2.is_minc2(f) for f in filenames):
        return None
    U, lo, hi, vr, info = minc2.read_table(filenames)
    if U.dtype == np.float64:
        return ("real", U)
    if U.dtype == np.float32:
        return ("stored", U, None, None, 0.0, max(U.shape[0], 1))
    if not vr[1] > vr[0]:
        raise ValueError("the volumes' valid_range is empty")
    return ("stored", U, (hi - lo) / (vr[1] - vr[0]), lo, vr[0], info.slice_len if info.nslices > 1 else max(U.shape[0], 1))


def vertexTable(filenames: Sequence, column: int = 1) -> np.ndarray:
    """V x n matrix from one whitespace/comma separated text file per subject, `column` 1-based
    (R/minc_interface.R:1134-1150)."""
    cols = []
    for f in filenames:
        if isinstance(f, np.ndarray):
            a = np.asarray(f, dtype=np.float64)
            cols.append(a if a.ndim == 1 else a[:, column - 1])
            continue
        if not os.path.exists(f):
            raise FileNotFoundError(f"Error opening input file: {f}.")
        a = np.loadtxt(f, delimiter="," if str(f).endswith((".csv", ".csv.gz")) else None, ndmin=2)
        cols.append(a[:, column - 1])
    if len({len(c) for c in cols}) != 1:
        raise ValueError("all vertex files must have the same number of vertices")
    return np.asfortranarray(np.column_stack(cols))


def _stage_mask(mask, V, reader):
    if mask is None:
        return None
    m = reader(mask) if not isinstance(mask, np.ndarray) else np.asarray(mask, dtype=np.float64).reshape(-1)
    if isinstance(m, StoredVolume):
        m = m.real()
    if m.shape[0] != V:
        raise ValueError("Mask Volume input volume dimension mismatch.")          # :837
    return m


# ----------------------------------------------------------------------------- parseLmFormula
def _is_file_term(data, name, rows) -> bool:
    """A right-hand-side term whose entries are volumes (file names or arrays) rather than numbers:
    parseLmFormula's `file.info(term[1])` / '\\.mnc$' test (R/minc_interface.R:1266, :1295-1296)."""
    try:
        c = data[name]
    except Exception:  # noqa: BLE001
        return False
    first = c.iloc[rows[0]] if hasattr(c, "iloc") else c[rows[0]]
    if isinstance(first, np.ndarray):
        return first.ndim >= 1
    if isinstance(first, (str, os.PathLike)):
        path = os.fspath(first)
        return path.endswith(".mnc") or os.path.exists(path)
    return False


def parseLmFormula(rhs: str, data, rows):
    """Mirror of parseLmFormula (R/minc_interface.R:1234-1324): finds a voxel-wise (file) term on the right-hand
    side.  Returns dict(matrixFound, mmatrix (static part or None), rows (coefficient names), right (the file
    column's entries))."""
    intercept, terms = parse_rhs(rhs)
    file_terms = [t[0] for t in terms if len(t) == 1 and _is_file_term(data, t[0], rows)]
    if not file_terms:
        return dict(matrixFound=False, mmatrix=None, rows=None, right=None)
    name = file_terms[0]
    if len(terms) == 1:
        # one term: design [1 | z]; the reference labels the intercept 'Intercept' here (:1270)
        return dict(matrixFound=True, mmatrix=None, rows=["Intercept", name], right=_column(data, name, rows))
    if "*" in rhs or ":" in rhs:
        raise FormulaError("Only + sign allowed when using filenames")           # :1298
    if len(terms) > 2 or len(file_terms) > 1:
        raise FormulaError("Only 2 terms allowed when using filenames on the RHS")   # :1301
    other = [t[0] for t in terms if t[0] != name][0]
    X, names, _ = model_matrix(other if intercept else other + " - 1", data, rows)   # :1304-1311
    return dict(matrixFound=True, mmatrix=X, rows=names + [name], right=_column(data, name, rows))


def _cov_attrs(n, plm):
    """model / df attributes in the file-covariate case: the reference derives them from the STATIC model
    matrix alone (R/minc_voxel_statistics.R:908, :929-935), so Fdf1 = ncol(mmatrix) - 1, Fdf2 = n - ncol(mmatrix)
    (0, 0 for the 1 x 1 NA matrix of the no-static case); mirrored as is."""
    X = plm["mmatrix"]
    p = len(plm["rows"])
    if X is None:
        model, f1, f2 = np.array([[np.nan]]), 0, 0
    else:
        model, f1, f2 = X, X.shape[1] - 1, n - X.shape[1]
    return model, [[f1, f2]] + [f2] * p


# ----------------------------------------------------------------------------- mincLm
def mincLm(formula: str, data, subset=None, mask=None, maskval=None, parallel=None,
           reader: Callable = default_reader, device: int = 0) -> MincMatrix:
    """Linear model at every voxel.  `data[lhs]` holds one volume (file name or array) per subject.

    parallel: None (one GPU) or ("local", N) -> N GPUs of this box, contiguous voxel shards
    (replaces create_parallel_mask + mclapply, R/minc_voxel_statistics.R:848-906).  Unlike the
    reference's parallel path, maskval is honoured on both paths."""
    lhs, rhs = split_formula(formula)
    rows = _subset_rows(data, subset)
    plm = parseLmFormula(rhs, data, rows)                                        # :796
    filenames = _column(data, lhs, rows)
    lo, hi = core.mask_range(maskval)                                            # :781-787
    staged = _minc2_stage(filenames) if (reader is default_reader and not plm["matrixFound"]) else None
    Y_staged = staged[1] if staged is not None and staged[0] == "real" else None
    if Y_staged is not None:
        staged = None
    first = reader(filenames[0]) if staged is None and Y_staged is None else None
    if staged is not None or (isinstance(first, StoredVolume) and not plm["matrixFound"]):
        # the volumes travel in their stored type; the device forms the doubles
        if staged is not None:
            U, scale, offset, vmin, slice_len = staged[1:]
        else:
            vols = [first] + [reader(f) for f in filenames[1:]]
            U, scale, offset, vmin, slice_len = mincTableStored(vols)
        X, names, _ = model_matrix(rhs, data, rows)
        m = _stage_mask(mask, U.shape[0], reader)
        out = core.lm_fit_stored(U, X, scale, offset, voxel_min=vmin, slice_len=slice_len, mask=m, mask_lo=lo, mask_hi=hi,
                                 device=device, n_gpus=int(parallel[1]) if parallel is not None else None)
        n, p = X.shape
        attrs = {
            "likeVolume": filenames[0], "filenames": list(filenames), "model": X, "data": data,
            "call": dict(fn="mincLm", formula=formula, subset=subset, mask=mask, maskval=maskval, parallel=parallel),
            "stat-type": _stat_type(p), "df": _df_list(n, p), "model-colnames": names,
            "_stored": dict(U=U, scale=scale, offset=offset, voxel_min=vmin, slice_len=slice_len),
            "_mask": m, "_mask_range": (lo, hi), "_rows": rows, "_rhs": rhs, "_lhs": lhs,
        }
        return MincMatrix(out, colnames=_colnames(names), attrs=attrs, r_class=("mincLm", "mincMultiDim", "matrix"))
    if isinstance(first, StoredVolume):
        # stored volumes combined with a voxel-wise covariate: expand on the host like the CPU reader would
        stored_reader = reader

        def reader(item):
            v = stored_reader(item)
            return v.real() if isinstance(v, StoredVolume) else v

    Y = Y_staged if Y_staged is not None else mincTable(filenames, reader=reader)
    V, n = Y.shape
    m = _stage_mask(mask, V, reader)
    if plm["matrixFound"]:
        # a volume on the right-hand side: per-voxel design [mmatrix | z_v] (minc2_model with filenames_right)
        Z = mincTable(plm["right"], reader=reader)
        if Z.shape != Y.shape:
            raise ValueError("the right-hand-side volumes must match the left-hand-side volumes")
        out = core.lm_fit_cov(Y, Z, plm["mmatrix"], mask=m, mask_lo=lo, mask_hi=hi, device=device,
                              n_gpus=int(parallel[1]) if parallel is not None else None)
        model, dflist = _cov_attrs(n, plm)
        p = len(plm["rows"])
        attrs = {
            "likeVolume": filenames[0], "filenames": list(filenames), "model": model, "data": data,
            "call": dict(fn="mincLm", formula=formula, subset=subset, mask=mask, maskval=maskval, parallel=parallel),
            "stat-type": _stat_type(p), "df": dflist, "model-colnames": plm["rows"],
            "_Y": Y, "_Z": Z, "_Xs": plm["mmatrix"], "_mask": m, "_mask_range": (lo, hi), "_rows": rows, "_rhs": rhs, "_lhs": lhs,
        }
        return MincMatrix(out, colnames=_colnames(plm["rows"]), attrs=attrs, r_class=("mincLm", "mincMultiDim", "matrix"))
    X, names, _ = model_matrix(rhs, data, rows)
    n_gpus = None
    if parallel is not None:
        n_gpus = int(parallel[1])
    out = core.lm_fit(Y, X, mask=m, mask_lo=lo, mask_hi=hi, device=device, n_gpus=n_gpus)
    p = X.shape[1]
    attrs = {
        "likeVolume": filenames[0], "filenames": list(filenames), "model": X, "data": data,
        "call": dict(fn="mincLm", formula=formula, subset=subset, mask=mask, maskval=maskval, parallel=parallel),
        "stat-type": _stat_type(p), "df": _df_list(n, p), "model-colnames": names,
        "_Y": Y, "_mask": m, "_mask_range": (lo, hi), "_rows": rows, "_rhs": rhs, "_lhs": lhs,
    }
    return MincMatrix(out, colnames=_colnames(names), attrs=attrs, r_class=("mincLm", "mincMultiDim", "matrix"))


# ----------------------------------------------------------------------------- vertexLm / anatLm
def vertexLm(formula: str, data, subset=None, column: int = 1, device: int = 0) -> MincMatrix:
    lhs, rhs = split_formula(formula)
    if "$" in rhs:
        raise FormulaError("$ Not Permitted in Formula")                        # minc_vertex_statistics.R:380
    rows = _subset_rows(data, subset)
    plm = parseLmFormula(rhs, data, rows)                                        # minc_vertex_statistics.R:391
    filenames = _column(data, lhs, rows)
    Y = vertexTable(filenames, column=column)
    if plm["matrixFound"]:
        Z = vertexTable(plm["right"], column=column)                             # :398-401
        if Z.shape != Y.shape:
            raise ValueError("the right-hand-side vertex files must match the left-hand-side files")
        out = core.lm_fit_cov(Y, Z, plm["mmatrix"], device=device)               # vertex_lm_loop with data_right
        model, dflist = _cov_attrs(Y.shape[1], plm)
        p = len(plm["rows"])
        attrs = {"likeVolume": filenames[0], "model": model, "filenames": list(filenames), "stat-type": _stat_type(p),
                 "data": data, "call": dict(fn="vertexLm", formula=formula, subset=subset, column=column),
                 "df": dflist, "model-colnames": plm["rows"]}
        return MincMatrix(out, colnames=_colnames(plm["rows"]), attrs=attrs,
                          r_class=("vertexLm", "vertexMultiDim", "matrix"))
    X, names, _ = model_matrix(rhs, data, rows)
    out = core.vertex_lm(Y, X, device=device)
    n, p = X.shape
    attrs = {"likeVolume": filenames[0], "model": X, "filenames": list(filenames), "stat-type": _stat_type(p),
             "data": data, "call": dict(fn="vertexLm", formula=formula, subset=subset, column=column),
             "df": _df_list(n, p), "model-colnames": names, "_Y": Y, "_mask": None,
             "_mask_range": (core.DEFAULT_MASK_LO, core.DEFAULT_MASK_HI), "_rows": rows, "_rhs": rhs, "_lhs": lhs}
    return MincMatrix(out, colnames=_colnames(names), attrs=attrs, r_class=("vertexLm", "vertexMultiDim", "matrix"))


def anatLm(formula: str, data, anat, subset=None, weights=None, device: int = 0) -> MincMatrix:
    """`anat`: subjects x structures matrix (anatGetAll output); formula has no left-hand side
    ("~ Sex").  Rows of the result are structures (R/minc_anatomy.R:1125-1128, :1153)."""
    lhs, rhs = split_formula(formula)
    if "$" in rhs:
        raise FormulaError("$ Not Permitted in Formula")
    rows = _subset_rows(data, subset)
    for term in rhs.replace("*", "+").replace(":", "+").split("+"):
        t = term.strip()
        if t and t not in ("0", "1", "-1"):
            c = data[t]
            if np.asarray(c.iloc[0] if hasattr(c, "iloc") else c[0]).ndim > 0:
                raise ValueError("A term in your model is a matrix, this use of `anatLm` is deprecated, Please merge "
                                 "the two anatomy matrices using rowbind or similar and use a group add a group "
                                 "indicator to your `data.frame`")                # minc_anatomy.R:1115-1121
    X, names, _ = model_matrix(rhs, data, rows)
    A = np.asarray(anat, dtype=np.float64)
    if A.ndim != 2 or A.shape[0] != len(data):
        raise ValueError("anat must be a subjects x structures matrix with one row per row of data")
    Y = np.asfortranarray(A[rows, :].T)                                          # structures x subjects
    if weights is not None:
        if len(weights) != X.shape[0]:
            raise ValueError("Incorrect number of weights supplied, need one per observation")   # :1132-1134
        out = core.vertex_wlm(Y, X, weights, device=device)                      # vertex_wlm_loop, :1135-1142
    else:
        out = core.vertex_lm(Y, X, device=device)
    n, p = X.shape
    rownames = list(getattr(anat, "columns", [])) or None
    attrs = {"atlas": getattr(anat, "atlas", None), "definitions": getattr(anat, "definitions", None), "model": X,
             "data": data, "call": dict(fn="anatLm", formula=formula, subset=subset), "stat-type": _stat_type(p),
             "df": _df_list(n, p), "model-colnames": names, "_Y": Y, "_mask": None,
             "_mask_range": (core.DEFAULT_MASK_LO, core.DEFAULT_MASK_HI), "_rows": rows, "_rhs": rhs, "_lhs": lhs}
    return MincMatrix(out, colnames=_colnames(names), rownames=rownames, attrs=attrs, r_class=("anatModel", "matrix"))


# ----------------------------------------------------------------------------- anova
def _anova_result(out, X, names, assign, data, call, filenames, r_class, rownames=None):
    n, p = X.shape
    nterms = max(assign) if assign else 0
    labels = []
    for t in range(1, nterms + 1):                       # term labels as attr(terms(formula), "term.labels")
        cols = [names[i] for i in range(p) if assign[i] == t]
        labels.append(call["_term_labels"][t - 1] if call.get("_term_labels") else cols[0])
    dfr = n - p
    dflist = [[sum(1 for a in assign if a == t), dfr] for t in range(1, nterms + 1)]   # :329-338
    attrs = {"model": X, "filenames": filenames, "stat-type": ["F"] * nterms, "df": dflist, "data": data,
             "call": {k: v for k, v in call.items() if not k.startswith("_")}, "model-colnames": names,
             "assign": list(assign)}
    return MincMatrix(out, colnames=labels, rownames=rownames, attrs=attrs, r_class=r_class)


def _term_labels(rhs):
    from .formula import parse_rhs
    _, terms = parse_rhs(rhs)
    return [":".join(t) for t in terms]


def mincAnova(formula: str, data, subset=None, mask=None, maskval=None, reader: Callable = default_reader,
              device: int = 0) -> MincMatrix:
    """Sequential (type I) F-statistic of every model term at every voxel
    (R/minc_voxel_statistics.R:495-693 -> per_voxel_anova, src/minc_anova.c:142-304)."""
    lhs, rhs = split_formula(formula)
    rows = _subset_rows(data, subset)
    X, names, assign = model_matrix(rhs, data, rows)
    filenames = _column(data, lhs, rows)
    Y = mincTable(filenames, reader=reader)
    m = _stage_mask(mask, Y.shape[0], reader)
    lo, hi = core.mask_range(maskval)
    out = core.anova_fit(Y, X, assign, mask=m, mask_lo=lo, mask_hi=hi, device=device)
    call = dict(fn="mincAnova", formula=formula, subset=subset, mask=mask, maskval=maskval, _term_labels=_term_labels(rhs))
    res = _anova_result(out, X, names, assign, data, call, list(filenames), ("mincMultiDim", "matrix"))
    res.attrs["likeVolume"] = filenames[0]
    return res


def vertexAnova(formula: str, data, subset=None, column: int = 1, device: int = 0) -> MincMatrix:
    """R/minc_vertex_statistics.R:301-344 -> vertex_anova_loop (src/vertex_modelling.c:178)."""
    lhs, rhs = split_formula(formula)
    rows = _subset_rows(data, subset)
    X, names, assign = model_matrix(rhs, data, rows)
    filenames = _column(data, lhs, rows)
    Y = vertexTable(filenames, column=column)
    out = core.anova_fit(Y, X, assign, device=device)
    call = dict(fn="vertexAnova", formula=formula, subset=subset, column=column, _term_labels=_term_labels(rhs))
    return _anova_result(out, X, names, assign, data, call, list(filenames), ("vertexMultiDim", "matrix"))


def anatAnova(formula: str, data, anat, subset=None, device: int = 0) -> MincMatrix:
    """anatAnova (R/minc_anatomy.R) -> vertex_anova_loop on the structures x subjects matrix."""
    lhs, rhs = split_formula(formula)
    rows = _subset_rows(data, subset)
    X, names, assign = model_matrix(rhs, data, rows)
    A = np.asarray(anat, dtype=np.float64)
    if A.ndim != 2 or A.shape[0] != len(data):
        raise ValueError("anat must be a subjects x structures matrix with one row per row of data")
    Y = np.asfortranarray(A[rows, :].T)
    out = core.anova_fit(Y, X, assign, device=device)
    call = dict(fn="anatAnova", formula=formula, subset=subset, _term_labels=_term_labels(rhs))
    return _anova_result(out, X, names, assign, data, call, None, ("anatModel", "matrix"),
                         rownames=list(getattr(anat, "columns", [])) or None)


# ----------------------------------------------------------------------------- mincFDR
def smooth_spline_at_last_knot(x, y, df: float = 3.0) -> float:
    """predict(smooth.spline(x, y, df = df), x = max(x))$y for distinct knots x (all of them are knots when there are
    fewer than 50, R's rule): the natural cubic smoothing spline g minimising sum (y - g(x))^2 + lam * int g''^2 with lam
    chosen so that the trace of the smoother matrix equals `df`.  stats::smooth.spline is R's (third party, not in the
    reference tree); this is the textbook Reinsch form (Green & Silverman 1994, section 2.3).  R matches `df` only to the
    tolerance of its spar search (control.spar tol 1e-4), so agreement with R is to ~1e-4, not to the bit."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n = x.shape[0]
    if n < 4 or np.any(np.diff(x) <= 0):
        raise ValueError("smooth.spline needs at least four increasing, distinct x values")
    if not 1.0 < df <= n:
        raise ValueError("you must supply 1 < df <= n")
    h = np.diff(x)
    Q = np.zeros((n, n - 2))
    Rm = np.zeros((n - 2, n - 2))
    for i in range(n - 2):
        Q[i, i] = 1.0 / h[i]
        Q[i + 1, i] = -1.0 / h[i] - 1.0 / h[i + 1]
        Q[i + 2, i] = 1.0 / h[i + 1]
        Rm[i, i] = (h[i] + h[i + 1]) / 3.0
        if i + 1 < n - 2:
            Rm[i, i + 1] = Rm[i + 1, i] = h[i + 1] / 6.0
    K = Q @ np.linalg.solve(Rm, Q.T)
    d, U = np.linalg.eigh(0.5 * (K + K.T))
    d = np.maximum(d, 0.0)                      # two exact zeros: constants and lines are not penalised
    lo, hi = -60.0, 60.0                        # bisection on log(lam): trace falls from n to 2
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if np.sum(1.0 / (1.0 + np.exp(mid) * d)) > df:
            lo = mid
        else:
            hi = mid
    lam = np.exp(0.5 * (lo + hi))
    g = U @ ((U.T @ y) / (1.0 + lam * d))
    return float(g[-1])


def _estimate_pi0(stat, stat_type, df1, df2, mask, method, device):
    """pi0 of fast.qvalue (lambda = seq(0, 0.9, 0.05), pi0.method = "smoother", smooth.df = 3; R/minc_misc.R:367-374,
    437-451) or of qvalue::pi0est (lambda = seq(0.05, 0.95, 0.05)): mean(p >= lambda) / (1 - lambda) on the GPU,
    smoothed on the host, evaluated at max(lambda), capped at 1."""
    lam = np.arange(0, 19) * 0.05 if method in ("pFDR", "fastqvalue") else np.arange(1, 20) * 0.05
    frac = core.fdr_pi0_fractions(stat, stat_type, df1, df2, mask=mask, lambdas=lam, device=device)
    pi0 = min(smooth_spline_at_last_knot(lam, frac / (1.0 - lam), 3.0), 1.0)
    if pi0 <= 0:
        raise ValueError("ERROR: The estimated pi0 <= 0. Check that you have valid p-values or use another lambda method.")
    return pi0


def mincFDR(buffer: MincMatrix, columns: Optional[Sequence[str]] = None, mask=None, df=None, method: str = "FDR",
            statType=None, reader: Callable = default_reader, device: int = 0) -> MincMatrix:
    """False discovery rate q-values of a mincLm / mincAnova / vertexLm / anatLm result
    (mincFDR.mincMultiDim, R/minc_FDR.R:254-520).

    The beta, R-squared and logLik columns are dropped (:267-292); every remaining t / F column is converted to
    p-values (pt2 / pf) and adjusted over the voxels with mask > 0.5 on the GPU: method "FDR" / "p.adjust"
    (Benjamini-Hochberg), "BY", "pFDR" / "fastqvalue" (fast.qvalue, R/minc_misc.R:366-560) or "qvalue" (:429-438).
    Returns a V x ncols matrix of class c("mincQvals","mincMultiDim","matrix") with colnames "qvalue-<column>", 1
    outside the mask, and attributes `thresholds` (5 x ncols, rows 0.01 .. 0.2), `likeVolume`, `DF`."""
    if method not in core.FDR_METHODS:
        raise ValueError(f"unknown method {method!r}: one of {sorted(core.FDR_METHODS)}")
    if not isinstance(buffer, MincMatrix):
        raise TypeError("buffer must be the result of mincLm / mincAnova / vertexLm / anatLm")
    stattype = list(statType) if statType is not None else buffer.attrs.get("stat-type")
    if stattype is None:
        raise ValueError("Error: need to specify the type of statistic.")                       # :301
    dfl = df if df is not None else buffer.attrs.get("df")
    A = np.asarray(buffer)
    names = list(buffer.colnames)
    if len(stattype) == 1 and A.shape[1] != 1:
        stattype = stattype * A.shape[1]
    keep = [i for i, st in enumerate(stattype) if st not in ("beta", "R-squared", "logLik")]
    stattype = [stattype[i] for i in keep]
    names = [names[i] for i in keep]
    A = A[:, keep]
    if any(st not in ("t", "F") for st in stattype):
        raise ValueError("Error: not all the stat types are recognized. Currently allowed are: t F")   # :318-323 (GPU: t, F)
    if dfl is None:
        raise ValueError("Error: need to specify the degrees of freedom")                       # :344
    dfl = list(dfl)
    if len(dfl) == 1 and A.shape[1] != 1:
        dfl = dfl * A.shape[1]
    if len(dfl) != A.shape[1]:
        raise ValueError("Error: df needs to be of either length 1 or the same length as number of columns in the buffer")
    if columns is None:
        columns = names
    m = None
    if mask is not None:
        m = _stage_mask(mask, A.shape[0], reader)
    out = np.ones((A.shape[0], len(columns)), order="F")
    thresholds = np.full((5, len(columns)), np.nan)
    new_dfs = []
    for i, cname in enumerate(columns):
        c = names.index(cname)
        d = dfl[c]
        new_dfs.append(d)
        if stattype[c] == "t":
            st, d1, d2 = core.STAT_T, float(d[0] if isinstance(d, (list, tuple, np.ndarray)) else d), 0.0
        else:
            st, d1, d2 = core.STAT_F, float(d[0]), float(d[1])
        pi0 = 1.0
        if core.FDR_METHODS[method] >= 2:
            pi0 = _estimate_pi0(A[:, c], st, d1, d2, m, method, device)
        q, th = core.fdr(A[:, c], st, d1, d2, mask=m, device=device, method=method, pi0=pi0)
        out[:, i] = q
        thresholds[:, i] = th
    th = MincMatrix(thresholds, colnames=list(columns), rownames=[f"{x:g}" for x in core.FDR_LEVELS])
    return MincMatrix(out, colnames=[f"qvalue-{c}" for c in columns], rownames=buffer.rownames,
                      attrs={"thresholds": th, "likeVolume": buffer.attrs.get("likeVolume"), "DF": new_dfs},
                      r_class=("mincQvals", "mincMultiDim", "matrix"))


def vertexFDR(buffer: MincMatrix, **kw) -> MincMatrix:
    """vertexFDR.vertexMultiDim -> mincFDR.mincMultiDim (R/minc_FDR.R:528-536)."""
    return mincFDR(buffer, **kw)


def anatFDR(buffer: MincMatrix, **kw) -> MincMatrix:
    """anatFDR.anatModel -> vertexFDR.vertexMultiDim (R/minc_FDR.R:571-573)."""
    return mincFDR(buffer, **kw)


# ----------------------------------------------------------------------------- vertexTFCE
class BicObj:
    """A triangle mesh as read_obj() returns it (class "bic_obj"): `vertex_matrix` 3 x nvertices coordinates and
    `triangle_matrix` 3 x ntriangles 1-based vertex numbers.  Reading .obj files is I/O outside this package; the two
    matrices are what the statistics need."""

    def __init__(self, vertex_matrix, triangle_matrix):
        self.vertex_matrix = np.asarray(vertex_matrix, dtype=np.float64)
        self.triangle_matrix = np.asarray(triangle_matrix)
        if self.vertex_matrix.ndim != 2 or self.vertex_matrix.shape[0] != 3 or self.triangle_matrix.ndim != 2 \
                or self.triangle_matrix.shape[0] != 3:
            raise ValueError("vertex_matrix must be 3 x nvertices and triangle_matrix 3 x ntriangles")


def obj_to_adjacency(obj: BicObj):
    """The adjacency list vertexTFCE derives from a bic_obj: obj_to_graph (R/minc_graphs.R:8-24) builds an undirected
    igraph from the edge list cbind(triangle_matrix[1:2, ], triangle_matrix[2:3], triangle_matrix[c(1, 3), ]) -- as
    written, the middle term has no comma, so it is the single pair (triangle_matrix[2], triangle_matrix[3]), the second
    and third vertex of the FIRST triangle, and the 2-3 edges of the other triangles enter only where a neighbouring
    triangle lists them as its 1-2 or 1-3 edge -- and vertexTFCE takes as_adj_list(graph) - 1
    (R/minc_vertex_statistics.R:776-781).  Restated literally; neighbours are listed once (igraph repeats a neighbour per
    parallel edge, which changes no cluster)."""
    tm = np.asarray(obj.triangle_matrix).astype(np.int64)
    nv = obj.vertex_matrix.shape[1]
    edges = np.concatenate([tm[0:2, :], tm.reshape(-1, order="F")[1:3].reshape(2, 1), tm[[0, 2], :]], axis=1) - 1
    if edges.size and (edges.min() < 0 or edges.max() >= nv):
        raise ValueError("triangle_matrix refers to a vertex that vertex_matrix does not hold")
    nbrs = [set() for _ in range(nv)]
    for a, b in edges.T:
        nbrs[int(a)].add(int(b))
        nbrs[int(b)].add(int(a))
    return [np.array(sorted(s), dtype=np.int32) for s in nbrs]


def _surface_and_weights(surface, weights):
    """vertexTFCE.numeric's preamble (R/minc_vertex_statistics.R:751-781): a bic_obj gives its mesh areas as default
    weights and its graph as the adjacency."""
    if isinstance(surface, BicObj):
        if weights is None:
            weights = core.mesh_area(surface.vertex_matrix, surface.triangle_matrix)
        surface = obj_to_adjacency(surface)
    return surface, weights


def vertexTFCE(x, surface, R: int = 500, alternative: str = "two.sided", E: float = 0.5, H: float = 2.0, nsteps: int = 100,
               weights=None, side: str = "both", replace: bool = False, parallel=None, seed=None, idx=None,
               device: int = 0, column: int = 1):
    """Threshold-free cluster enhancement on a surface graph (R/minc_vertex_statistics.R:731-990).

    `surface`: the 0-based adjacency list form vertexTFCE accepts (one neighbour array per vertex), a CSR pair
    (ptr, idx), or a BicObj mesh (then the default weights are its vertex areas, mesh_area); igraph objects and .obj
    files are outside this package.  Dispatch as in the reference:
      * numeric vector -> enhanced vector (vertexTFCE.numeric)
      * plain matrix   -> every column enhanced (vertexTFCE.matrix)
      * file name(s)   -> read with vertexTable(column=) then as above (vertexTFCE.character)
      * vertexLm result -> permutation test on every tvalue- column: list(tfce, randomization_dist, args) of class
        c("vertexTFCE_randomization","vertex_randomization") (vertexTFCE.vertexLm)."""
    if side not in core.TFCE_SIDES:
        raise ValueError("'arg' should be one of 'both', 'positive', 'negative'")
    surface, weights = _surface_and_weights(surface, weights)
    if isinstance(x, MincMatrix) and "vertexLm" in x.r_class:
        if alternative not in ("two.sided", "greater"):
            raise ValueError("'arg' should be one of 'two.sided', 'greater'")
        if "_Y" not in x.attrs:
            raise TypeError("x must be the result of vertexLm")
        columns = [i for i, c in enumerate(x.colnames) if c.startswith("tvalue-")]
        names = [x.colnames[c] for c in columns]
        X = x.attrs["model"]
        n = X.shape[0]
        original = core.graph_tfce(np.asarray(x)[:, columns], surface, E=E, H=H, nsteps=nsteps, side=side, weights=weights,
                                   device=device)
        if idx is None:
            idx = draw_indices(n, R, replace=replace, seed=seed)
        idx = np.asarray(idx, dtype=np.int32)
        if idx.shape != (n, R):
            raise ValueError("idx must be n x R")
        dist = core.randomize_tfce(x.attrs["_Y"], X, idx, columns, surface, two_sided=(alternative == "two.sided"), E=E,
                                   H=H, nsteps=nsteps, side=side, weights=weights, device=device,
                                   n_gpus=int(parallel[1]) if parallel is not None else None)
        out = MincRandomization(tfce=MincMatrix(original, colnames=names),
                                randomization_dist=MincMatrix(dist, colnames=names),
                                args=dict(call=x.attrs.get("call"), side=side, alternative=alternative))
        out.r_class = ("vertexTFCE_randomization", "vertex_randomization")
        return out
    if isinstance(x, (str, os.PathLike)):                       # vertexTFCE.character: one file -> a vector
        x = vertexTable([x], column=column)[:, 0]
    elif isinstance(x, (list, tuple)) and len(x) and all(isinstance(f, (str, os.PathLike)) for f in x):
        x = vertexTable(list(x), column=column)[:, 0] if len(x) == 1 else vertexTable(list(x), column=column)
    A = np.asarray(x, dtype=np.float64)
    res = core.graph_tfce(A, surface, E=E, H=H, nsteps=nsteps, side=side, weights=weights, device=device)
    if A.ndim == 2:
        return MincMatrix(res, colnames=getattr(x, "colnames", None))
    return res


# ----------------------------------------------------------------------------- mincRandomize
# ----------------------------------------------------------------------------- mincTFCE
def volume_dims(like) -> tuple:
    """Dimensions (file order) of a like-volume: a 3-tuple, a 3-D array, or a '.npy' path (likeVolume(x) in R is a MINC
    file whose header carries the sizes)."""
    if isinstance(like, (tuple, list)) and len(like) == 3 and all(isinstance(v, (int, np.integer)) for v in like):
        return tuple(int(v) for v in like)
    if isinstance(like, StoredVolume):
        like = like.voxels
    if isinstance(like, np.ndarray) and like.ndim == 3:
        return tuple(int(v) for v in like.shape)
    if isinstance(like, (str, os.PathLike)) and os.fspath(like).endswith(".npy") and os.path.exists(os.fspath(like)):
        shp = np.load(os.fspath(like), mmap_mode="r").shape
        if len(shp) == 3:
            return tuple(int(v) for v in shp)
    if isinstance(like, (str, os.PathLike)) and os.path.exists(os.fspath(like)):
        from . import minc2
        if minc2.is_minc2(like):
            with minc2.Minc2Volume(like) as v:
                if len(v.sizes) == 3:
                    return v.sizes
    raise ValueError("like_volume must give three dimensions: a 3-tuple, a 3-D array, a .npy volume or a MINC 2 file")


def mincTFCE(x, d=0.1, E: float = 0.5, H: float = 2.0, side: str = "both", like_volume=None, R: int = 500,
             alternative: str = "two.sided", replace: bool = False, parallel=None, seed=None, idx=None,
             connectivity: int = 6, voxel_volume: float = 1.0, device: int = 0):
    """Threshold-free cluster enhancement of volumes (R/minc_voxel_statistics.R:1186-1439).  Dispatch as in the
    reference:
      * numeric vector (mincSingleDim) -> enhanced vector
      * matrix / mincMultiDim -> every column enhanced (NA / NaN -> 0 first; `d` may hold one step per column)
      * mincLm result -> permutation test on every tvalue- column: list(tfce, randomization_dist, args) of class
        c("mincTFCE_randomization","minc_randomization") (mincTFCE.mincLm).
    The reference hands each volume to the external `TFCE` program of minc-stuffs; here the tree's own graph TFCE runs on
    the voxel grid (see rmg_volume_tfce): `connectivity` 6 and `voxel_volume` = the product of the step sizes reproduce
    what vertexTFCE(x, neighbour_list(.., 6), weights = voxel volume) gives (tests/testthat/test_mincTFCE.R:153-170)."""
    if side not in core.TFCE_SIDES:
        raise ValueError("'arg' should be one of 'both', 'positive', 'negative'")
    if isinstance(x, MincMatrix) and "mincLm" in x.r_class:
        if alternative not in ("two.sided", "greater"):
            raise ValueError("'arg' should be one of 'two.sided', 'greater'")
        if "_Y" not in x.attrs and "_stored" not in x.attrs:
            raise TypeError("x must be the result of mincLm")
        if np.ndim(d) != 0:
            raise ValueError("the permutation test takes one height step d")
        dims = volume_dims(x.attrs["likeVolume"] if like_volume is None else like_volume)
        columns = [i for i, c in enumerate(x.colnames) if c.startswith("tvalue-")]
        names = [x.colnames[c] for c in columns]
        X = x.attrs["model"]
        n = X.shape[0]
        stats = np.nan_to_num(np.asarray(x)[:, columns], nan=0.0, posinf=np.inf, neginf=-np.inf)
        original = core.volume_tfce(stats, dims, d=d, E=E, H=H, side=side, connectivity=connectivity,
                                    voxel_volume=voxel_volume, device=device)
        if idx is None:
            idx = draw_indices(n, R, replace=replace, seed=seed)
        idx = np.asarray(idx, dtype=np.int32)
        if idx.shape != (n, R):
            raise ValueError("idx must be n x R")
        lo, hi = x.attrs["_mask_range"]
        Yd = x.attrs["_Y"] if "_Y" in x.attrs else _expand_stored(**x.attrs["_stored"])
        dist = core.randomize_volume_tfce(Yd, X, idx, columns, dims, two_sided=(alternative == "two.sided"), d=d,
                                          E=E, H=H, side=side, connectivity=connectivity, voxel_volume=voxel_volume,
                                          mask=x.attrs["_mask"], mask_lo=lo, mask_hi=hi,
                                          n_gpus=int(parallel[1]) if parallel is not None else None)
        out = MincRandomization(tfce=MincMatrix(original, colnames=names, attrs={"likeVolume": x.attrs.get("likeVolume")}),
                                randomization_dist=MincMatrix(dist, colnames=names),
                                args=dict(call=x.attrs.get("call"), side=side, alternative=alternative))
        out.r_class = ("mincTFCE_randomization", "minc_randomization")
        return out
    if like_volume is None and isinstance(x, MincMatrix):
        like_volume = x.attrs.get("likeVolume")
    A = np.asarray(x, dtype=np.float64)
    if A.ndim == 3 and like_volume is None:
        like_volume, A = A.shape, A.reshape(-1)
    dims = volume_dims(like_volume)
    A = np.nan_to_num(A, nan=0.0, posinf=np.inf, neginf=-np.inf)                    # setNA(0) %>% setNaN(0)
    kw = dict(E=E, H=H, side=side, connectivity=connectivity, voxel_volume=voxel_volume, device=device)
    if A.ndim == 2 and np.ndim(d) != 0:
        steps = np.resize(np.asarray(d, dtype=np.float64), A.shape[1])              # mapply recycles d over the columns
        res = np.column_stack([core.volume_tfce(A[:, j], dims, d=float(steps[j]), **kw) for j in range(A.shape[1])])
    else:
        res = core.volume_tfce(A, dims, d=float(np.asarray(d).reshape(-1)[0]), **kw)
    if A.ndim == 2:
        return MincMatrix(res, colnames=getattr(x, "colnames", None), attrs={"likeVolume": like_volume})
    return res


class MincRandomization(dict):
    """list(stats, randomization_dist, args) of class c("mincLm_randomization","minc_randomization")."""
    r_class = ("mincLm_randomization", "minc_randomization")


def draw_indices(n: int, R: int, replace: bool = False, seed=None) -> np.ndarray:
    """n x R index matrix: column r is sample(seq_len(n), replace = replace) (:1472).  In R the
    indices come from R's own RNG; numpy's generator stands in for it here."""
    rng = np.random.default_rng(seed)
    if replace:
        return rng.integers(0, n, size=(n, R)).astype(np.int32)
    return np.column_stack([rng.permutation(n) for _ in range(R)]).astype(np.int32)


def mincRandomize(x: MincMatrix, R: int = 500, alternative: str = "two.sided", replace: bool = False,
                  parallel=None, columns: Optional[Sequence[int]] = None, seed=None, idx=None,
                  device: int = 0) -> MincRandomization:
    """Permutation null of the extreme statistics of a mincLm / vertexLm / anatLm fit.

    All predictor columns of the data frame move together (:1457-1474), so the permuted design is
    the model matrix with its ROWS permuted; the whole voxel-wise refit per permutation happens on
    the GPU with Y resident (the reference re-runs mincLm, re-reading every file, :1477).
    `columns`: 0-based result columns (default: the tvalue- columns, :1333)."""
    if alternative not in ("two.sided", "greater"):
        raise ValueError("'arg' should be one of 'two.sided', 'greater'")        # match.arg
    if not isinstance(x, MincMatrix) or ("_Y" not in x.attrs and "_stored" not in x.attrs):
        raise TypeError("x must be the result of mincLm / vertexLm / anatLm")
    X = x.attrs["model"]
    n = x.attrs["_Y"].shape[1] if "_Y" in x.attrs else X.shape[0]
    if columns is None:
        columns = [i for i, c in enumerate(x.colnames) if c.startswith("tvalue-")]
    columns = [int(c) for c in columns]
    if idx is None:
        idx = draw_indices(n, R, replace=replace, seed=seed)
    idx = np.asarray(idx, dtype=np.int32)
    if idx.shape != (n, R):
        raise ValueError("idx must be n x R")
    lo, hi = x.attrs["_mask_range"]
    n_gpus = int(parallel[1]) if parallel is not None else None
    if "_Z" in x.attrs:
        # a voxel-wise covariate: its files are re-sampled with the static predictors and the per-voxel design is
        # refitted (the loop re-evaluates the whole call, R/minc_voxel_statistics.R:1468-1477)
        dist = core.randomize_cov(x.attrs["_Y"], x.attrs["_Z"], x.attrs["_Xs"], idx, columns,
                                  two_sided=(alternative == "two.sided"), mask=x.attrs["_mask"], mask_lo=lo, mask_hi=hi,
                                  device=device, n_gpus=n_gpus)
    elif "_stored" in x.attrs:    # the fit ran on volumes in their stored type: so does the permutation test
        sv = x.attrs["_stored"]
        dist = core.randomize_stored(sv["U"], X, idx, columns, scale=sv["scale"], offset=sv["offset"],
                                     voxel_min=sv["voxel_min"], slice_len=sv["slice_len"],
                                     two_sided=(alternative == "two.sided"), mask=x.attrs["_mask"], mask_lo=lo, mask_hi=hi,
                                     device=device, n_gpus=n_gpus)
    else:
        dist = core.randomize(x.attrs["_Y"], X, idx, columns, two_sided=(alternative == "two.sided"),
                              mask=x.attrs["_mask"], mask_lo=lo, mask_hi=hi, device=device, n_gpus=n_gpus)
    out = MincRandomization(
        stats=MincMatrix(np.asarray(x)[:, columns], colnames=[x.colnames[c] for c in columns]),
        randomization_dist=MincMatrix(dist, colnames=[x.colnames[c] for c in columns]),
        args=dict(call=x.attrs.get("call"), alternative=alternative),
    )
    return out


def vertexRandomize(x: MincMatrix, R: int = 500, alternative: str = "two.sided", replace: bool = False, parallel=None,
                    columns: Optional[Sequence[int]] = None, seed=None, idx=None, device: int = 0) -> MincRandomization:
    """vertexRandomize.vertexLm (R/minc_vertex_statistics.R:992-1024): the same permutation loop on a vertexLm result,
    returned with class c("vertexLm_randomization","vertex_randomization")."""
    if not isinstance(x, MincMatrix) or "vertexLm" not in x.r_class:
        raise TypeError("x must be the result of vertexLm")
    out = mincRandomize(x, R=R, alternative=alternative, replace=replace, parallel=parallel, columns=columns, seed=seed,
                        idx=idx, device=device)
    out.r_class = ("vertexLm_randomization", "vertex_randomization")
    return out


def thresholds(x: MincRandomization, probs=(0.01, 0.05, 0.1, 0.2)) -> MincMatrix:
    """apply(dist, 2, quantile, probs = 1 - probs) with rownames "1%", "5%", ... (:1577-1581).
    R's default quantile type 7 == numpy's default 'linear' method.  For a mincFDR result (thresholds.mincQvals,
    R/minc_FDR.R:659-664) it returns the stored thresholds attribute."""
    if isinstance(x, MincMatrix) and "mincQvals" in x.r_class:
        return x.attrs["thresholds"]
    d = np.asarray(x["randomization_dist"])
    q = np.quantile(d, [1 - p for p in probs], axis=0)
    return MincMatrix(q, colnames=x["randomization_dist"].colnames, rownames=[f"{p * 100:g}%" for p in probs])


# ----------------------------------------------------------------------------- model selection on the logLik column
def _model_size(x: MincMatrix):
    if not isinstance(x, MincMatrix) or "model" not in x.attrs or "logLik" not in (x.colnames or []):
        raise TypeError("expected a mincLm / vertexLm / anatLm result")
    X = np.asarray(x.attrs["model"])
    return X.shape[0], X.shape[1] + 1                  # n; parameters + 1 for the error scale


def AIC(x: MincMatrix, k: float = 2.0) -> np.ndarray:
    """AIC.mincLm / .vertexLm / .anatModel (R/minc_model_selection.R:2-17): k * (dfs - logLik), one value per row --
    the reference's own expression, including its use of k as the overall factor."""
    _, dfs = _model_size(x)
    return k * (dfs - x.col("logLik"))


def AICc(x: MincMatrix) -> np.ndarray:
    """AICc.mincLm (R/minc_model_selection.R:33-39): AIC + 2 (k + 1)(k + 2) / (n - k - 2) with k = ncol(model) + 1."""
    n, k = _model_size(x)
    return AIC(x) + 2.0 * (k + 1) * (k + 2) / (n - k - 2)


def BIC(x: MincMatrix) -> np.ndarray:
    """BIC.mincLm (R/minc_model_selection.R:52-58): -2 logLik + k log(n)."""
    n, k = _model_size(x)
    return -2.0 * x.col("logLik") + k * np.log(n)


def compare_models(*models: MincMatrix, metric: Callable = AICc) -> MincMatrix:
    """compare_models (R/minc_model_selection.R:104-127): rows x models matrix of the metric (lower is better), class
    c("model_comparison","matrix"), attributes formulae and metric."""
    if len(models) < 2:
        raise ValueError("more than one model must be supplied")                      # check_model_list, :66-68
    if any(not isinstance(m, MincMatrix) or m.r_class != models[0].r_class for m in models):
        raise ValueError("all models must be the same type")
    if any(m.shape[0] != models[0].shape[0] for m in models):
        raise ValueError("all models must have the same number of rows")
    table = np.column_stack([np.asarray(metric(m)) for m in models])
    return MincMatrix(table, attrs={"formulae": [(m.attrs.get("call") or {}).get("formula") for m in models],
                                    "metric": getattr(metric, "__name__", str(metric))},
                      r_class=("model_comparison", "matrix"))


def summary_model_comparison(comp: MincMatrix):
    """summary.model_comparison (R/minc_model_selection.R:168-190): wins per model (ties count for every minimiser),
    as a list of dict(model, formula, wins) sorted by wins."""
    t = np.asarray(comp)
    wins = (t == t.min(axis=1, keepdims=True)).sum(axis=0)
    rows = [dict(model=str(j + 1), formula=f, wins=int(w)) for j, (f, w) in enumerate(zip(comp.attrs["formulae"], wins))]
    return sorted(rows, key=lambda r: r["wins"])

