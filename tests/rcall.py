"""`.Call` from Python: drives the routine tables of a library built from rminc_b200/csrc/rshim.c (UNCHANGED) plus the R
stand-ins of oracle/rstub -- and, where /root/reference was available at build time, the reference's own routines
(vertex_lm_loop, vertex_wlm_loop, vertex_anova_loop, per_voxel_anova, minc2_model) in the SAME library, registered as
src/RMINC_init.c:57-88 registers them.  Test infrastructure only.

    R = rcall.load()                       # oracle/_ref/librshim_ref.so, else oracle/librshim_test.so
    out = R.call("rmincgpu_vertex_lm_loop", Y, rcall.NA_MATRIX, X)     # numpy in, numpy out; RError on Rf_error
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "librshim_ref.so")
SHIM_LIB = os.path.join(ROOT, "oracle", "librshim_test.so")
NILSXP, LGLSXP, INTSXP, REALSXP, STRSXP, VECSXP, EXTPTRSXP, RAWSXP = 0, 10, 13, 14, 16, 19, 22, 24
NA_INTEGER = -2147483648


class RError(RuntimeError):
    """The routine called Rf_error (longjmp back to the dispatcher), or .Call itself refused the call."""

    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


class NaMatrix:
    """`matrix()`: the 1 x 1 logical NA that vertexLm / mincLm pass for "no right-hand-side data" / "no static part"."""


class Strings(list):
    """A character vector argument."""


class Raw(bytes):
    """A raw vector argument."""


class Logical:
    def __init__(self, value):
        self.value = NA_INTEGER if value is None else int(bool(value))


class Handle:
    """An SEXP kept as is (external pointers returned by one call and passed to the next)."""

    def __init__(self, ptr):
        self.ptr = ptr


NA_MATRIX = NaMatrix()


class RStub:
    def __init__(self, path):
        self.path = path
        lib = self.lib = C.CDLL(path)
        vp = C.c_void_p
        lib.rstub_dyn_load.restype = C.c_int
        lib.rstub_routine.restype = C.c_char_p
        lib.rstub_routine.argtypes = [C.c_int, C.POINTER(C.c_int)]
        lib.rstub_dotcall.restype = C.c_int
        lib.rstub_dotcall.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp), C.POINTER(vp), C.c_char_p, C.c_size_t]
        lib.rstub_nil.restype = vp
        lib.rstub_new_real.restype = vp
        lib.rstub_new_real.argtypes = [vp, C.c_ssize_t, C.c_int, C.c_int]
        lib.rstub_new_int.restype = vp
        lib.rstub_new_int.argtypes = [vp, C.c_ssize_t, C.c_int, C.c_int, C.c_int]
        lib.rstub_new_raw.restype = vp
        lib.rstub_new_raw.argtypes = [vp, C.c_ssize_t]
        lib.rstub_new_strings.restype = vp
        lib.rstub_new_strings.argtypes = [C.POINTER(C.c_char_p), C.c_int]
        lib.rstub_new_list.restype = vp
        lib.rstub_new_list.argtypes = [C.c_ssize_t]
        lib.rstub_list_set.argtypes = [vp, C.c_ssize_t, vp]
        lib.rstub_list_get.restype = vp
        lib.rstub_list_get.argtypes = [vp, C.c_ssize_t]
        for f in ("rstub_type", "rstub_nrow", "rstub_ncol"):
            getattr(lib, f).restype = C.c_int
            getattr(lib, f).argtypes = [vp]
        lib.rstub_length.restype = C.c_ssize_t
        lib.rstub_length.argtypes = [vp]
        lib.rstub_data.restype = vp
        lib.rstub_data.argtypes = [vp]
        lib.rstub_release.argtypes = [vp]
        lib.rstub_set_option_int.argtypes = [C.c_char_p, C.c_int]
        lib.rstub_register_volume.argtypes = [C.c_char_p, vp, C.POINTER(C.c_int64)]
        self.n_routines = lib.rstub_dyn_load()
        self.has_reference = any(name == "vertex_lm_loop" for name, _ in self.routines())

    def routines(self):
        out, na = [], C.c_int(0)
        for i in range(self.n_routines):
            out.append((self.lib.rstub_routine(i, C.byref(na)).decode(), na.value))
        return out

    def options(self, **kw):
        """options(name = int) ; None removes the option."""
        for k, v in kw.items():
            self.lib.rstub_set_option_int(k.encode(), NA_INTEGER if v is None else int(v))

    # ---- numpy / python -> SEXP
    def sexp(self, x):
        lib = self.lib
        if x is None:
            return lib.rstub_nil(), False
        if isinstance(x, Handle):
            return x.ptr, False
        if isinstance(x, NaMatrix):
            v = np.array([NA_INTEGER], dtype=np.int32)
            return lib.rstub_new_int(v.ctypes.data, 1, 1, 1, 1), True
        if isinstance(x, Logical):
            v = np.array([x.value], dtype=np.int32)
            return lib.rstub_new_int(v.ctypes.data, 1, 1, -1, 1), True
        if isinstance(x, Raw):
            return lib.rstub_new_raw(C.cast(C.c_char_p(bytes(x)), C.c_void_p), len(x)), True
        if isinstance(x, Strings):
            arr = (C.c_char_p * len(x))(*[s.encode() for s in x])
            return lib.rstub_new_strings(arr, len(x)), True
        if isinstance(x, (list, tuple)):
            lst = lib.rstub_new_list(len(x))
            for i, e in enumerate(x):
                p, _ = self.sexp(e)
                lib.rstub_list_set(lst, i, p)
            return lst, True
        if isinstance(x, (bool, np.bool_)):
            return self.sexp(Logical(x))
        a = np.asarray(x)
        if a.dtype.kind in "iu" or a.dtype.kind == "b":
            v = np.asfortranarray(a, dtype=np.int32)
            nrow, ncol = (v.shape[0], v.shape[1]) if v.ndim == 2 else (v.size, -1)
            return lib.rstub_new_int(v.ctypes.data, v.size, nrow, ncol, int(a.dtype.kind == "b")), True
        v = np.asfortranarray(a, dtype=np.float64)
        nrow, ncol = (v.shape[0], v.shape[1]) if v.ndim == 2 else (v.size, -1)
        return lib.rstub_new_real(v.ctypes.data, v.size, nrow, ncol), True

    # ---- SEXP -> numpy / python (copies)
    def value(self, p):
        lib = self.lib
        t = lib.rstub_type(p)
        if t == NILSXP:
            return None
        if t == EXTPTRSXP:
            return Handle(p)
        n, nrow, ncol = lib.rstub_length(p), lib.rstub_nrow(p), lib.rstub_ncol(p)
        if t == VECSXP:
            return [self.value(lib.rstub_list_get(p, i)) for i in range(n)]
        ctype, dtype = {REALSXP: (C.c_double, np.float64), INTSXP: (C.c_int, np.int32), LGLSXP: (C.c_int, np.int32)}[t]
        flat = np.ctypeslib.as_array(C.cast(lib.rstub_data(p), C.POINTER(ctype)), shape=(max(n, 1),))[:n].astype(dtype)
        return flat.reshape((nrow, ncol), order="F") if ncol >= 0 else flat

    def call(self, name, *args):
        """.Call(name, ...): returns the value as numpy (a Handle for an external pointer); raises RError."""
        ptrs, owned = [], []
        for a in args:
            p, mine = self.sexp(a)
            ptrs.append(p)
            if mine:
                owned.append(p)
        arr = (C.c_void_p * max(len(ptrs), 1))(*ptrs)
        res = C.c_void_p()
        err = C.create_string_buffer(512)
        rc = self.lib.rstub_dotcall(name.encode(), len(ptrs), arr, C.byref(res), err, len(err))
        try:
            if rc != 0:
                raise RError(rc, err.value.decode(errors="replace"))
            out = self.value(res.value)
            if not isinstance(out, Handle):
                self.lib.rstub_release(res)
            return out
        finally:
            for p in owned:
                self.lib.rstub_release(p)

    def release(self, handle):
        self.lib.rstub_release(handle.ptr)

    def register_volumes(self, prefix, Y, dims):
        """In-memory "MINC files" for minc2_model / per_voxel_anova: column j of Y (V x n) becomes file
        f"{prefix}{j}.mnc" of dimensions `dims` (file order).  Returns the names; keep Y alive while they are used."""
        sizes = (C.c_int64 * 3)(*[int(d) for d in dims])
        names = []
        for j in range(Y.shape[1]):
            nm = f"{prefix}{j:05d}.mnc"
            self.lib.rstub_register_volume(nm.encode(), Y[:, j].ctypes.data, sizes)
            names.append(nm)
        return Strings(names)


_cached = None


def load():
    """The shim + reference library when it exists (built by oracle/Makefile `ref` where /root/reference is present),
    else the shim alone (`make rshim`, built on demand)."""
    global _cached
    if _cached is None:
        path = REF_LIB
        if not os.path.exists(path):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "rshim"], check=True)
            path = SHIM_LIB
        _cached = RStub(path)
    return _cached
