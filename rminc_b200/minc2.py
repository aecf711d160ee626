"""MINC 2 volumes through the library's own reader (rmg_minc2_*, rminc_b200/csrc/minc2_reader.cpp): what
mincGetVolume / mincTable / minc2_model's miopen_volume + miget_real_value_hyperslab do in the reference
(R/minc_interface.R:556-574, 1031-1077; src/modelling_functions.c:806-829, 1037-1055), without libminc or libhdf5.

`read_volume(path)` yields a StoredVolume for uint16 / float32 files (so mincLm ships 2 or 4 bytes per sample to the GPU)
and the real doubles for every other stored type; `read_table(paths)` decodes a whole study in parallel straight into the
V x n matrix the stored-type entry points take."""
from __future__ import annotations

import ctypes as C
import os
from typing import Sequence

import numpy as np

from . import _lib

MAX_DIMS = 5
U8, I8, U16, I16, U32, I32, F32, F64 = range(1, 9)
_NP = {U8: np.uint8, I8: np.int8, U16: np.uint16, I16: np.int16, U32: np.uint32, I32: np.int32, F32: np.float32, F64: np.float64}


class _Info(C.Structure):
    _fields_ = [("ndims", C.c_int32), ("dtype", C.c_int32), ("elem_size", C.c_int32), ("has_real_range", C.c_int32),
                ("sizes", C.c_int64 * MAX_DIMS), ("nvoxels", C.c_int64), ("nslices", C.c_int64), ("slice_len", C.c_int64),
                ("valid_range", C.c_double * 2), ("starts", C.c_double * MAX_DIMS), ("steps", C.c_double * MAX_DIMS),
                ("dircos", (C.c_double * 3) * MAX_DIMS), ("dimnames", (C.c_char * 16) * MAX_DIMS)]


def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


class Minc2Volume:
    """An open MINC 2 file.  Attributes mirror what the reference asks libminc for: `sizes` (miget_dimension_sizes),
    `dimnames`, `starts`, `steps` (miget_dimension_starts / _separations), `valid_range`, `dtype`."""

    def __init__(self, path):
        self.path = os.fspath(path)
        lib = _lib.load()
        h = C.c_void_p()
        err = _lib.errbuf()
        rc = lib.rmg_minc2_open(self.path.encode(), C.byref(h), err, len(err))
        if rc == _lib.RMG_ERR_IO and not os.path.exists(self.path):
            raise FileNotFoundError(err.value.decode(errors="replace"))
        _lib.check(rc, err)
        self._h = h
        info = _Info()
        lib.rmg_minc2_get_info(self._h, C.byref(info))
        nd = info.ndims
        self.sizes = tuple(int(info.sizes[d]) for d in range(nd))
        self.dimnames = tuple(bytes(info.dimnames[d]).split(b"\0", 1)[0].decode() for d in range(nd))
        self.starts = tuple(float(info.starts[d]) for d in range(nd))
        self.steps = tuple(float(info.steps[d]) for d in range(nd))
        self.direction_cosines = tuple(tuple(float(info.dircos[d][k]) for k in range(3)) for d in range(nd))
        self.dtype = np.dtype(_NP[info.dtype])
        self.dtype_code = int(info.dtype)
        self.nvoxels = int(info.nvoxels)
        self.nslices = int(info.nslices)
        self.slice_len = int(info.slice_len)
        self.valid_range = (float(info.valid_range[0]), float(info.valid_range[1]))
        self.has_real_range = bool(info.has_real_range)

    def close(self):
        if getattr(self, "_h", None):
            _lib.load().rmg_minc2_close(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def ranges(self):
        """(image-min, image-max), `nslices` entries each."""
        lo, hi = np.empty(self.nslices), np.empty(self.nslices)
        err = _lib.errbuf()
        _lib.check(_lib.load().rmg_minc2_read_ranges(self._h, lo.ctypes.data, hi.ctypes.data, err, len(err)), err)
        return lo, hi

    def stored(self, n_threads=None) -> np.ndarray:
        """The voxels as the file stores them, shaped like the image (slowest dimension first)."""
        out = np.empty(self.sizes, dtype=self.dtype)
        err = _lib.errbuf()
        _lib.check(_lib.load().rmg_minc2_read_stored(self._h, out.ctypes.data, out.nbytes, n_threads or host_threads(), err, len(err)), err)
        return out

    def real(self, n_threads=None) -> np.ndarray:
        """The doubles miget_real_value_hyperslab(MI_TYPE_DOUBLE) yields, flattened in file order (mincGetVolume)."""
        out = np.empty(self.nvoxels, dtype=np.float64)
        err = _lib.errbuf()
        _lib.check(_lib.load().rmg_minc2_read_real(self._h, out.ctypes.data, out.size, n_threads or host_threads(), err, len(err)), err)
        return out

    def voxel_volume(self) -> float:
        """|product of the spatial steps| (what mincTFCE needs from the likeVolume)."""
        v = 1.0
        for name, step in zip(self.dimnames, self.steps):
            if name in ("xspace", "yspace", "zspace"):
                v *= abs(step)
        return v


def is_minc2(path) -> bool:
    """True for an HDF5 file (signature at 0 or behind a user block); MINC 1 (NetCDF) files are not."""
    try:
        with open(path, "rb") as f:
            head = f.read(8)
            if head == b"\x89HDF\r\n\x1a\n":
                return True
            for off in (512, 1024, 2048):
                f.seek(off)
                if f.read(8) == b"\x89HDF\r\n\x1a\n":
                    return True
    except OSError:
        pass
    return False


def read_volume(path):
    """StoredVolume for uint16 (with its real ranges) and float32 files whose ranges follow the slowest dimension; the
    real doubles otherwise."""
    from .api import StoredVolume
    with Minc2Volume(path) as v:
        if v.dtype_code == F32:
            return StoredVolume(v.stored())
        if v.dtype_code == U16 and v.nslices in (1, v.sizes[0]) and v.valid_range[1] > v.valid_range[0]:
            lo, hi = v.ranges()
            return StoredVolume(v.stored(), image_min=lo, image_max=hi, valid_range=v.valid_range)
        return v.real()


def read_table(paths: Sequence, dtype=None, n_threads=None, pinned: bool = False, out=None):
    """mincTable for MINC 2 files, decoded in parallel into one Fortran-ordered V x n matrix.

    dtype None: the files' own stored type if it is uint16 / float32, float64 otherwise.  Returns (U, image_min, image_max,
    valid_range, info) with image_min / image_max [nslices x n] for uint16 tables and None otherwise; `info` is the first
    file's Minc2Volume (closed; its attributes stay readable).  pinned=True stages U in page-locked memory
    (rmg_host_alloc) so the upload runs at the link rate -- "auto": only up to 2 GiB, because page-locking fresh memory (~1.4 GB/s)
    costs a one-off study more than the library's pinned ring costs a pageable matrix; out= reuses the matrix of an earlier call of the same shape and type
    (page-locking half a gigabyte costs more than decoding it)."""
    paths = [os.fspath(p) for p in paths]
    with Minc2Volume(paths[0]) as first:
        pass
    code = {None: first.dtype_code if first.dtype_code in (U16, F32) else F64, np.uint16: U16, np.float32: F32,
            np.float64: F64}.get(dtype if dtype is None else np.dtype(dtype).type)
    if code is None:
        raise TypeError("read_table: dtype must be uint16, float32 or float64")
    V, n = first.nvoxels, len(paths)
    lib = _lib.load()
    npdt = np.dtype(_NP[code])
    if out is not None:
        if out.shape != (V, n) or out.dtype != npdt or not out.flags.f_contiguous:
            raise ValueError("out= must be a Fortran-ordered V x n matrix of the table's type")
        U = out
    elif pinned is True or (pinned == "auto" and V * n * npdt.itemsize <= (2 << 30) and lib.rmg_device_count() > 0):
        nbytes = V * n * npdt.itemsize
        p = lib.rmg_host_alloc(nbytes)
        if not p:
            raise MemoryError("rmg_host_alloc failed (no CUDA device?)")
        buf = (C.c_char * nbytes).from_address(p)
        U = np.frombuffer(buf, dtype=npdt).reshape((V, n), order="F")
        U = _Pinned(U, p)
    else:
        U = np.empty((V, n), dtype=npdt, order="F")
    lo = hi = None
    nsl = first.nslices
    if code == U16:
        lo, hi = np.empty((nsl, n), order="F"), np.empty((nsl, n), order="F")
    vr = np.zeros(2)
    arr = (C.c_char_p * n)(*[p.encode() for p in paths])
    err = _lib.errbuf()
    rc = lib.rmg_minc2_read_table(arr, n, U.ctypes.data, V, code, lo.ctypes.data if lo is not None else None,
                                  hi.ctypes.data if hi is not None else None, nsl, vr.ctypes.data, n_threads or host_threads(),
                                  err, len(err))
    if rc == _lib.RMG_ERR_IO and "Error opening input file" in err.value.decode(errors="replace"):
        raise FileNotFoundError(err.value.decode(errors="replace"))
    _lib.check(rc, err)
    return U, lo, hi, (float(vr[0]), float(vr[1])), first


class _Pinned(np.ndarray):
    """ndarray view of a rmg_host_alloc block that frees it when the last view dies."""

    def __new__(cls, arr, ptr):
        obj = arr.view(cls)
        obj._owner = _PinnedBlock(ptr)
        return obj

    def __array_finalize__(self, obj):
        self._owner = getattr(obj, "_owner", None)


class _PinnedBlock:
    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        if self.ptr:
            _lib.load().rmg_host_free(self.ptr)
            self.ptr = None
