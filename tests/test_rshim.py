"""The .Call shim (rminc_b200/csrc/rshim.c) EXECUTED on the CPU box: compiled unchanged against the R stand-ins of
oracle/rstub, dispatched by name through the table it registers (the analogue of src/RMINC_init.c:57-88).  Without a GPU
every routine must unpack its arguments, allocate its result and then fail with an R error carrying the library's
"no CUDA device ... no CPU fallback" message -- through Rf_error's longjmp, with nothing leaked or collected early.
Argument validation (the advisor's findings on the stored-type routines) is checked here too: it happens before the
C ABI is reached, so it needs no device."""
import os
import re

import numpy as np
import pytest

import rcall
from conftest import have_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def R(lib):
    return rcall.load()


def design(n):
    rng = np.random.default_rng(3)
    return np.column_stack([np.ones(n), np.arange(n) % 2, rng.normal(size=n)])


def test_dispatch_table_is_the_shims_registration(R):
    src = open(os.path.join(ROOT, "rminc_b200", "csrc", "rshim.c")).read()
    table = re.findall(r'\{"(rmincgpu_[a-z_]+)", \(DL_FUNC\)&(rmincgpu_[a-z_]+), (\d+)\}', src)
    assert len(table) >= 19 and all(a == b for a, b, _ in table)
    registered = dict(R.routines())
    for name, _, nargs in table:
        assert registered[name] == int(nargs), name
    if R.has_reference:   # the reference's own registration lines for the routines compiled in (src/RMINC_init.c:57-88)
        for name, nargs in (("vertex_lm_loop", 3), ("vertex_wlm_loop", 4), ("vertex_anova_loop", 3), ("minc2_model", 11),
                            ("per_voxel_anova", 7)):
            assert registered[name] == nargs


def test_dotcall_checks_like_r(R):
    Y, X = np.zeros((4, 6)), design(6)
    with pytest.raises(rcall.RError, match="Incorrect number of arguments"):
        R.call("rmincgpu_vertex_lm_loop", Y, X)
    with pytest.raises(rcall.RError, match="not in load table"):
        R.call("rmincgpu_no_such_routine", Y)


def every_routine_call(n=12, V=40):
    rng = np.random.default_rng(0)
    X = design(n)
    Y = np.asfortranarray(rng.normal(size=(V, n)))
    idx = np.column_stack([rng.permutation(n) + 1 for _ in range(5)]).astype(np.int32)      # 1-based, as sample() gives
    cols = np.array([6, 7, 8], dtype=np.int32)                                               # 1-based column numbers
    mask = (np.arange(V) % 3 > 0).astype(float)
    U = np.asfortranarray(rng.integers(0, 60000, (V, n)).astype(np.uint16))
    scale, offset = np.full((4, n), 1e-3), np.zeros((4, n))
    adj = [[v - 1] * (v > 0) + [v + 1] * (v < V - 1) for v in range(V)]
    adj = [np.asarray(a, dtype=np.int32) for a in adj]
    na = rcall.NA_MATRIX
    return {
        "rmincgpu_vertex_lm_loop": (Y, na, X),
        "rmincgpu_vertex_wlm_loop": (Y, na, X, np.ones(n)),
        "rmincgpu_lm": (Y, X, mask, 1.0, 99999999.0, 1),
        "rmincgpu_lm_cov": (Y, Y[:, ::-1].copy(), X[:, :2], None, 1.0, 99999999.0),
        "rmincgpu_lm_stored": (rcall.Raw(U.tobytes(order="F")), 1, float(V), scale, offset, 0.0, 10.0, X, None, 1.0, 99999999.0),
        "rmincgpu_randomize": (Y, X, mask, 1.0, 99999999.0, idx, cols, True, 1),
        "rmincgpu_randomize_cov": (Y, Y[:, ::-1].copy(), X[:, :2], mask, 1.0, 99999999.0, idx, cols, True, 1),
        "rmincgpu_randomize_stored": (rcall.Raw(U.tobytes(order="F")), 1, float(V), scale, offset, 0.0, 10.0, X, None, 1.0, 99999999.0,
                                      idx, cols, True, 1),
        "rmincgpu_anova": (Y, X, np.array([0, 1, 2], dtype=np.int32), None, 1.0, 99999999.0),
        "rmincgpu_fdr": (Y[:, 0].copy(), 1, np.array([float(n - 3)]), None),
        "rmincgpu_fdr_method": (Y[:, 0].copy(), 1, np.array([float(n - 3)]), None, 2, 0.9),
        "rmincgpu_fdr_pi0": (Y[:, 0].copy(), 1, np.array([float(n - 3)]), None, np.arange(0, 19) * 0.05),
        "rmincgpu_graph_tfce": (Y[:, 0].copy(), adj, 0.5, 2.0, 20, np.ones(V), 0),
        "rmincgpu_randomize_tfce": (Y, X, idx, cols, True, adj, np.ones(V), 0.5, 2.0, 20, 0),
        "rmincgpu_volume_tfce": (Y[:, 0].copy(), np.array([2.0, 4.0, 5.0]), 6, 1.0, 0.25, 0.5, 2.0, 0),
        "rmincgpu_randomize_volume_tfce": (Y, X, None, 1.0, 99999999.0, idx, cols, True, np.array([2.0, 4.0, 5.0]), 6, 1.0, 0.25,
                                           0.5, 2.0, 0, 1),
        "rmincgpu_dataset_create": (Y, mask, 1.0, 99999999.0, 1),
        "rmincgpu_minc2_model": (rcall.Strings(minc2_study(U, scale, offset)), na, X, None, 0.0, rcall.Strings([]), 1.0, 99999999.0, None,
                                 None, rcall.Strings(["lm"])),
        "rmincgpu_per_voxel_anova": (rcall.Strings(minc2_study(U, scale, offset)), X, np.array([0, 1, 2], dtype=np.int32), 0.0,
                                     rcall.Strings([]), 1.0, 99999999.0),
    }


_study_dir = None


def minc2_study(U, scale, offset, shape=(4, 2, 5)):
    """One MINC 2 file per column of U (uint16, `shape` voxels, per-slice ranges from scale / offset), written once."""
    global _study_dir
    import tempfile
    import h5mini
    if _study_dir is None:
        _study_dir = tempfile.mkdtemp(prefix="rmg_minc2_")
    names = []
    for j in range(U.shape[1]):
        nm = os.path.join(_study_dir, f"s{j:03d}_{U.shape[0]}.mnc")
        if not os.path.exists(nm):
            h5mini.write_minc2(nm, U[:, j].reshape(shape), image_min=offset[:, j], image_max=offset[:, j] + 65535.0 * scale[:, j],
                               valid_range=(0, 65535), chunks=(1,) + shape[1:])
        names.append(nm)
    return names


@pytest.mark.skipif(have_gpu(), reason="only meaningful on a machine without a CUDA device")
def test_every_routine_runs_and_raises_the_no_device_error(R):
    calls = every_routine_call()
    registered = dict(R.routines())
    handle_routines = {"rmincgpu_dataset_free", "rmincgpu_dataset_lm", "rmincgpu_dataset_anova", "rmincgpu_dataset_randomize",
                       "rmincgpu_dataset_fdr", "rmincgpu_dataset_fdr_method"}
    assert R.call("rmincgpu_launch_count")[0] == 0          # host-only: works without a device, nothing was launched
    for name in registered:
        if name.startswith("rmincgpu_") and name not in handle_routines and name != "rmincgpu_launch_count":
            assert name in calls, f"no test call for {name}"
    for name, args in calls.items():
        assert registered[name] == len(args), name
        with pytest.raises(rcall.RError, match="no CUDA device") as e:
            R.call(name, *args)
        assert e.value.code == 1, name                      # raised by the routine (Rf_error), not by the dispatcher
    # routines that take a handle refuse anything that is not an external pointer
    for name in sorted(handle_routines):
        args = [np.zeros(3)] + [np.zeros((2, 2))] * (registered[name] - 1)
        with pytest.raises(rcall.RError, match="not a GPU dataset handle"):
            R.call(name, *args)


def test_minc2_model_routine_validates_and_reports_like_the_reference(R, tmp_path):
    """rmincgpu_minc2_model takes minc2_model's own eleven arguments (src/modelling_functions.c:743-745): a missing file is
    the reference's "Error opening input file" (:809), a mask of another size its dimension-mismatch error, methods other
    than "lm" are refused -- all before any device is needed -- and the staging memory is released on every path."""
    import h5mini
    c = every_routine_call()
    args = list(c["rmincgpu_minc2_model"])
    files = list(args[0])

    def with_(i, v):
        a = list(args)
        a[i] = v
        return a

    with pytest.raises(rcall.RError, match=r"Error opening input file: /nonexistent/a.mnc\."):
        R.call("rmincgpu_minc2_model", *with_(0, rcall.Strings(["/nonexistent/a.mnc"] + files[1:])))
    if not have_gpu():
        with pytest.raises(rcall.RError, match="Error opening input file: .*gone.mnc"):      # not the first file: found while staging
            R.call("rmincgpu_minc2_model", *with_(0, rcall.Strings(files[:-1] + [str(tmp_path / "gone.mnc")])))
    with pytest.raises(rcall.RError, match='runs method "lm"'):
        R.call("rmincgpu_minc2_model", *with_(10, rcall.Strings(["ttest"])))
    with pytest.raises(rcall.RError, match="model matrix does not match the number of files"):
        R.call("rmincgpu_minc2_model", *with_(0, rcall.Strings(files[:-1])))
    small = str(tmp_path / "smallmask.mnc")
    h5mini.write_minc2(small, np.ones((2, 2, 5), dtype=np.uint8), valid_range=(0, 255))
    a = with_(4, 1.0)
    a[5] = rcall.Strings([small])
    with pytest.raises(rcall.RError, match="Mask Volume input volume dimension mismatch"):
        R.call("rmincgpu_minc2_model", *a)
    R.options(RMINC_GPU_N=0)
    try:
        with pytest.raises(rcall.RError, match="RMINC_GPU_N"):
            R.call("rmincgpu_minc2_model", *args)
    finally:
        R.options(RMINC_GPU_N=None)


def test_stored_type_arguments_are_validated_before_the_abi(R):
    """A 1-row table, a wrong slice_len, slice_len <= 0 / NA would make the C ABI read scale / offset out of bounds
    (it derives ceil(V / slice_len) rows itself); NA for two_sided / ngpu / dtype must not pass silently either."""
    n, V = 12, 40
    c = every_routine_call(n, V)
    base = list(c["rmincgpu_lm_stored"])

    def with_(i, v, routine="rmincgpu_lm_stored"):
        a = list(c[routine])
        a[i] = v
        return a

    with pytest.raises(rcall.RError, match=r"ceiling\(V / slice_len\) = 4 rows \(got 1 and 4\)"):
        R.call("rmincgpu_lm_stored", *with_(3, np.full((1, n), 1e-3)))
    with pytest.raises(rcall.RError, match=r"ceiling\(V / slice_len\) = 8 rows"):
        R.call("rmincgpu_lm_stored", *with_(6, 5.0))
    for bad in (0.0, -3.0, float("nan"), 2.5):
        with pytest.raises(rcall.RError, match="slice_len must be a whole number >= 1"):
            R.call("rmincgpu_lm_stored", *with_(6, bad))
    with pytest.raises(rcall.RError, match="nslices x n double matrices"):
        R.call("rmincgpu_lm_stored", *with_(4, np.zeros((4, n - 1))))
    with pytest.raises(rcall.RError, match="dtype must be 1"):
        R.call("rmincgpu_lm_stored", *with_(1, 7))
    with pytest.raises(rcall.RError, match="U holds"):
        R.call("rmincgpu_lm_stored", *with_(2, float(V + 1)))
    with pytest.raises(rcall.RError, match=r"ceiling\(V / slice_len\)"):
        R.call("rmincgpu_randomize_stored", *with_(4, np.zeros((3, n)), "rmincgpu_randomize_stored"))
    with pytest.raises(rcall.RError, match="two_sided must be TRUE or FALSE"):
        R.call("rmincgpu_randomize_stored", *with_(13, rcall.Logical(None), "rmincgpu_randomize_stored"))
    with pytest.raises(rcall.RError, match="two_sided must be TRUE or FALSE"):
        R.call("rmincgpu_randomize", *with_(7, rcall.Logical(None), "rmincgpu_randomize"))
    with pytest.raises(rcall.RError, match="ngpu must not be NA"):
        R.call("rmincgpu_randomize", *with_(8, np.array([rcall.NA_INTEGER], dtype=np.int32), "rmincgpu_randomize"))
    with pytest.raises(rcall.RError, match="idx must not contain NA"):
        bad_idx = c["rmincgpu_randomize"][5].copy()
        bad_idx[2, 1] = rcall.NA_INTEGER
        R.call("rmincgpu_randomize", *with_(5, bad_idx, "rmincgpu_randomize"))
    with pytest.raises(rcall.RError, match="Mask Volume input volume dimension mismatch"):
        R.call("rmincgpu_lm", *with_(2, np.ones(V - 1), "rmincgpu_lm"))
    R.options(RMINC_GPU_DEVICE=-2)
    try:
        with pytest.raises(rcall.RError, match="RMINC_GPU_DEVICE"):
            R.call("rmincgpu_vertex_lm_loop", *c["rmincgpu_vertex_lm_loop"])
    finally:
        R.options(RMINC_GPU_DEVICE=None)
    assert base == list(c["rmincgpu_lm_stored"])
