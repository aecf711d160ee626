"""ctypes binding of the C ABI in include/rminc_b200.h.

The shared library is the product: there is NO fallback.  If librminc_b200.so is
missing or a compute entry point cannot reach a CUDA device, the error is raised to
the caller (RmincGpuError), never papered over with a CPU path.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librminc_b200.so")

RMG_OK, RMG_ERR_INVALID, RMG_ERR_SINGULAR, RMG_ERR_CUDA, RMG_ERR_NO_DEVICE, RMG_ERR_IO = 0, 1, 2, 3, 4, 5

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)

# name -> (restype, argtypes); mirrors include/rminc_b200.h declaration by declaration
_FIT_HEAD = [C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_int]
_MASK = [C.c_void_p, C.c_double, C.c_double]
_ERR = [C.c_char_p, C.c_size_t]
_PERM = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
SIGNATURES = {
    "rmg_version": (C.c_int, []),
    "rmg_device_count": (C.c_int, []),
    "rmg_max_p": (C.c_int, []),
    "rmg_design_factor": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int), C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.POINTER(C.c_int)] + _ERR),
    "rmg_vertex_lm": (C.c_int, _FIT_HEAD + [C.c_void_p, C.c_int64, C.c_int] + _ERR),
    "rmg_vertex_wlm": (C.c_int, _FIT_HEAD + [C.c_void_p, C.c_void_p, C.c_int64, C.c_int] + _ERR),
    "rmg_lm_fit": (C.c_int, _FIT_HEAD + _MASK + [C.c_void_p, C.c_int64, C.c_int] + _ERR),
    "rmg_lm_fit_device": (C.c_int, _FIT_HEAD + _MASK + [C.c_void_p, C.c_int64, C.c_int, C.c_void_p] + _ERR),
    "rmg_lm_fit_cov": (C.c_int, [C.c_void_p] + _FIT_HEAD + _MASK + [C.c_void_p, C.c_int64, C.c_int] + _ERR),
    "rmg_vertex_lm_cov": (C.c_int, [C.c_void_p] + _FIT_HEAD + [C.c_void_p, C.c_int64, C.c_int] + _ERR),
    "rmg_lm_fit_cov_device": (C.c_int, [C.c_void_p] + _FIT_HEAD + _MASK + [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p] + _ERR),
    "rmg_lm_fit_stored": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_double,
                                    C.c_int64, C.c_void_p, C.c_int] + _MASK + [C.c_void_p, C.c_int64, C.c_int] + _ERR),
    "rmg_lm_fit_stored_device": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                           C.c_double, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int] + _MASK
                                 + [C.c_void_p, C.c_int64, C.c_int, C.c_void_p] + _ERR),
    "rmg_fdr": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int] + _ERR),
    "rmg_fdr_device": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_int, C.c_void_p] + _ERR),
    "rmg_fdr_method": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_int, C.c_double, C.c_int] + _ERR),
    "rmg_fdr_method_device": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int, C.c_double, C.c_int, C.c_void_p] + _ERR),
    "rmg_fdr_pi0_fractions": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_int,
                                        C.c_void_p, C.c_int] + _ERR),
    "rmg_fdr_pi0_fractions_device": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p,
                                               C.c_int, C.c_void_p, C.c_int, C.c_void_p] + _ERR),
    "rmg_pvalue": (C.c_double, [C.c_double, C.c_int, C.c_double, C.c_double]),
    "rmg_lm_fit_cov_multi": (C.c_int, [C.c_void_p] + _FIT_HEAD + _MASK + [C.c_void_p, C.c_int64, C.c_int] + _ERR),
    "rmg_lm_fit_stored_multi": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_double,
                                          C.c_int64, C.c_void_p, C.c_int] + _MASK + [C.c_void_p, C.c_int64, C.c_int] + _ERR),
    "rmg_graph_tfce": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                 C.c_int, C.c_int, C.c_void_p, C.c_int] + _ERR),
    "rmg_randomize_tfce": (C.c_int, _FIT_HEAD + _PERM + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int,
                                                          C.c_int, C.c_void_p, C.c_int] + _ERR),
    "rmg_randomize_tfce_multi": (C.c_int, _FIT_HEAD + _PERM + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                                                C.c_int, C.c_int, C.c_void_p, C.c_int] + _ERR),
    "rmg_neighbour_list": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64,
                                     C.POINTER(C.c_int64)] + _ERR),
    "rmg_mesh_area": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p] + _ERR),
    "rmg_volume_tfce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                  C.c_int, C.c_void_p, C.c_int] + _ERR),
    "rmg_randomize_volume_tfce": (C.c_int, _FIT_HEAD + _MASK + _PERM + [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double,
                                                                        C.c_double, C.c_int, C.c_void_p, C.c_int] + _ERR),
    "rmg_anova_fit": (C.c_int, _FIT_HEAD + [C.c_void_p] + _MASK + [C.c_void_p, C.c_int64, C.c_int] + _ERR),
    "rmg_anova_fit_device": (C.c_int, _FIT_HEAD + [C.c_void_p] + _MASK + [C.c_void_p, C.c_int64, C.c_int, C.c_void_p] + _ERR),
    "rmg_randomize": (C.c_int, _FIT_HEAD + _MASK + _PERM + [C.c_void_p, C.c_int] + _ERR),
    "rmg_randomize_device": (C.c_int, _FIT_HEAD + _MASK + _PERM + [C.c_void_p, C.c_int, C.c_void_p] + _ERR),
    "rmg_lm_fit_multi": (C.c_int, _FIT_HEAD + _MASK + [C.c_void_p, C.c_int64, C.c_int] + _ERR),
    "rmg_randomize_multi": (C.c_int, _FIT_HEAD + _MASK + _PERM + [C.c_void_p, C.c_int] + _ERR),
    "rmg_randomize_cov": (C.c_int, [C.c_void_p] + _FIT_HEAD + _MASK + _PERM + [C.c_void_p, C.c_int] + _ERR),
    "rmg_randomize_cov_multi": (C.c_int, [C.c_void_p] + _FIT_HEAD + _MASK + _PERM + [C.c_void_p, C.c_int] + _ERR),
    "rmg_randomize_stored": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_double,
                                       C.c_int64, C.c_void_p, C.c_int] + _MASK + _PERM + [C.c_void_p, C.c_int] + _ERR),
    "rmg_randomize_stored_multi": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_double, C.c_int64, C.c_void_p, C.c_int] + _MASK + _PERM
                                   + [C.c_void_p, C.c_int] + _ERR),
    "rmg_dataset_create": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int] + _MASK + [C.c_int, C.POINTER(C.c_void_p)] + _ERR),
    "rmg_dataset_create_stored": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_double,
                                            C.c_int64] + _MASK + [C.c_int, C.POINTER(C.c_void_p)] + _ERR),
    "rmg_dataset_free": (None, [C.c_void_p]),
    "rmg_dataset_voxels": (C.c_int64, [C.c_void_p]),
    "rmg_dataset_subjects": (C.c_int, [C.c_void_p]),
    "rmg_dataset_gpus": (C.c_int, [C.c_void_p]),
    "rmg_dataset_sync": (C.c_int, [C.c_void_p] + _ERR),
    "rmg_dataset_lm_fit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int64] + _ERR),
    "rmg_dataset_anova": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64] + _ERR),
    "rmg_dataset_randomize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int] + _PERM + [C.c_void_p] + _ERR),
    "rmg_dataset_fdr": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p] + _ERR),
    "rmg_dataset_fdr_method": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_double] + _ERR),
    "rmg_dataset_fdr_pi0_fractions": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_int,
                                                C.c_void_p] + _ERR),
    "rmg_release_host_buffers": (None, []),
    "rmg_host_alloc": (C.c_void_p, [C.c_size_t]),
    "rmg_host_free": (None, [C.c_void_p]),
    "rmg_h2d_probe": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_double)] + _ERR),
    "rmg_last_kernel_ms": (C.c_double, []),
    "rmg_last_fit_variant": (C.c_int, []),
    "rmg_launch_count": (C.c_int64, []),
    # MINC 2 / HDF5 reader (host only)
    "rmg_minc2_open": (C.c_int, [C.c_char_p, C.c_void_p] + _ERR),
    "rmg_minc2_get_info": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rmg_minc2_read_ranges": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p] + _ERR),
    "rmg_minc2_read_stored": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int] + _ERR),
    "rmg_minc2_read_real": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int] + _ERR),
    "rmg_minc2_close": (None, [C.c_void_p]),
    "rmg_minc2_read_table": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_int64,
                                       C.c_void_p, C.c_int] + _ERR),
    "rmg_zlib_uncompress": (C.c_int64, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int]),
    "rmg_h5_read_dataset": (C.c_int, [C.c_char_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int] + _ERR),
    "rmg_h5_read_attribute": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_char_p,
                                        C.c_size_t] + _ERR),
}


class RmincGpuError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code


class SingularDesignError(RmincGpuError):
    """dpotri info > 0 -- the reference raises an R error (src/modelling_functions.c:551-557)."""


_lib = None


def load() -> C.CDLL:
    """Load librminc_b200.so (built in-tree by rminc_b200/build.py).  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RmincGpuError(RMG_ERR_NO_DEVICE,
                            f"{LIB_PATH} is missing: run `python -m rminc_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == the .so does not match the header
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, err: C.Array) -> None:
    if rc == RMG_OK:
        return
    msg = err.value.decode(errors="replace")
    if rc == RMG_ERR_SINGULAR:
        raise SingularDesignError(rc, msg)
    raise RmincGpuError(rc, msg)


def errbuf():
    return C.create_string_buffer(512)


def as_f64_fortran(a) -> np.ndarray:
    return np.asfortranarray(a, dtype=np.float64)


def ptr(a: np.ndarray) -> int:
    return a.ctypes.data
