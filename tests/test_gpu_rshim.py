"""Every routine rshim.c registers, EXECUTED on the GPU through `.Call`: mini-SEXPs in, SEXP out, compared with the
reference's own routines called through the SAME dispatcher with the same SEXPs (vertex_lm_loop, vertex_wlm_loop,
vertex_anova_loop, minc2_model, per_voxel_anova: src/RMINC_init.c:57-88, compiled from /root/reference where it lies)
and with the oracle for the routines the reference implements in R (mincRandomize's loop, mincFDR, mincTFCE).
The R error path (Rf_error -> longjmp) is exercised with the reference's own dpotri message."""
import numpy as np
import pytest

import rcall
from helpers import assert_rows_match, cfg1_data, design_cfg2, grid_adjacency

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.fixture(scope="module")
def R(lib):
    from rminc_b200 import core
    assert core.device_count() > 0
    return rcall.load()


def need_ref(R):
    if not R.has_reference:
        pytest.skip("oracle/_ref was not built on a machine with /root/reference")


def one_based(idx):
    return (np.asarray(idx) + 1).astype(np.int32)


def perm_indices(n, R_, seed=4):
    rng = np.random.default_rng(seed)
    return np.column_stack([rng.permutation(n) for _ in range(R_)])


def test_vertex_routines_against_the_reference_routines(R, oracle):
    need_ref(R)
    rng = np.random.default_rng(1)
    V, n = 900, 30
    X = design_cfg2(n)
    Y = np.asfortranarray(5 + rng.normal(size=(V, n)))
    Y[7] = 0.0
    na = rcall.NA_MATRIX
    assert_rows_match(R.call("rmincgpu_vertex_lm_loop", Y, na, X), R.call("vertex_lm_loop", Y, na, X), RTOL, "vertex_lm_loop")
    w = rng.uniform(0.5, 2.0, n)
    assert_rows_match(R.call("rmincgpu_vertex_wlm_loop", Y, na, X, w), R.call("vertex_wlm_loop", Y, na, X, w), RTOL, "wlm")
    asg = np.array([0, 1, 2, 3], dtype=np.int32)
    assert_rows_match(R.call("rmincgpu_anova", Y, X, asg, None, 1.0, 99999999.0), R.call("vertex_anova_loop", Y, X, asg), RTOL,
                      "vertex_anova_loop")
    # data_right a matrix: the voxel-wise covariate design, with and without a static part
    Z = np.asfortranarray(50 + 4 * rng.normal(size=(V, n)))
    Yc = np.asfortranarray(Y[8:])
    Zc = np.asfortranarray(Z[8:])
    assert_rows_match(R.call("rmincgpu_vertex_lm_loop", Yc, Zc, X[:, :3].copy()), R.call("vertex_lm_loop", Yc, Zc, X[:, :3].copy()),
                      RTOL, "vertex_lm_loop with data_right")
    assert_rows_match(R.call("rmincgpu_vertex_lm_loop", Yc, Zc, na), R.call("vertex_lm_loop", Yc, Zc, na), RTOL, "no static part")


def reference_minc2_model(R, Y, dims, X, mask, lo, hi):
    """.Call("minc2_model", filenames, matrix(), mmatrix, NULL, have_mask, mask, lo, hi, NULL, NULL, "lm") on in-memory files."""
    R.lib.rstub_clear_volumes()
    files = R.register_volumes("subject_", Y, dims)
    mcol = None if mask is None else np.asfortranarray(np.asarray(mask, dtype=np.float64).reshape(-1, 1))   # kept alive
    mfile = rcall.Strings([]) if mask is None else R.register_volumes("mask_", mcol, dims)
    try:
        return R.call("minc2_model", files, rcall.NA_MATRIX, X, None, float(mask is not None), mfile, float(lo), float(hi), None,
                      None, rcall.Strings(["lm"]))
    finally:
        R.lib.rstub_clear_volumes()


def test_minc2_model_and_per_voxel_anova_against_the_reference_routines(R, oracle):
    need_ref(R)
    dims = (12, 10, 9)
    Y, X, mask = cfg1_data(dims=dims)
    V, n = Y.shape
    for lo, hi, m in ((1.0, 99999999.0, mask), (2.0, 2.0, mask * 2), (1.0, 99999999.0, None)):
        ref = reference_minc2_model(R, Y, dims, X, m, lo, hi)
        got = R.call("rmincgpu_lm", Y, X, m, lo, hi, 1)
        inside = np.ones(V, bool) if m is None else (m > lo - 0.5) & (m < hi + 0.5)
        assert_rows_match(got[inside], ref[inside], RTOL, f"minc2_model lm lo={lo}")
        assert (got[~inside] == 0).all()                       # every column; the reference only clears column 0 (Q1)
        assert (ref[~inside, 0] == 0).all()
    asg = np.array([0, 1], dtype=np.int32)
    R.lib.rstub_clear_volumes()
    files = R.register_volumes("subject_", Y, dims)
    mcol = np.asfortranarray(mask.reshape(-1, 1))
    mfile = R.register_volumes("mask_", mcol, dims)
    try:
        ref = R.call("per_voxel_anova", files, X, asg, 1.0, mfile, 1.0, 99999999.0)
    finally:
        R.lib.rstub_clear_volumes()
    got = R.call("rmincgpu_anova", Y, X, asg, mask, 1.0, 99999999.0)
    assert_rows_match(got, ref, RTOL, "per_voxel_anova")


def test_minc2_model_routine_on_files_against_the_reference_routine(R, oracle, tmp_path):
    """.Call("rmincgpu_minc2_model", <minc2_model's eleven arguments>) on MINC 2 FILES (uint16, per-slice ranges, chunked +
    deflate; a uint8 label mask; then int16 volumes, then a covariate study) against .Call("minc2_model", ...) of the
    reference's object code on in-memory volumes that hold the doubles libminc's reader would have produced."""
    need_ref(R)
    import h5mini
    from rminc_b200 import api
    dims = (12, 10, 9)
    Y, X, mask = cfg1_data(dims=dims)
    V, n = Y.shape
    files, cols = [], []
    for j in range(n):
        real = Y[:, j].reshape(dims)
        lo, hi = real.min(axis=(1, 2)) - 0.01, real.max(axis=(1, 2)) + 0.01
        vox = np.round((real - lo[:, None, None]) / (hi - lo)[:, None, None] * 65535.0).astype(np.uint16)
        f = str(tmp_path / f"s{j:02d}.mnc")
        h5mini.write_minc2(f, vox, image_min=lo, image_max=hi, valid_range=(0, 65535), chunks=(1, 10, 9), compress=4)
        files.append(f)
        cols.append(api.StoredVolume(vox, image_min=lo, image_max=hi).real())
    Yq = np.asfortranarray(np.stack(cols, axis=1))
    mfile = str(tmp_path / "mask.mnc")
    h5mini.write_minc2(mfile, (mask.reshape(dims) * 2).astype(np.uint8), valid_range=(0, 255), chunks=dims)
    F = rcall.Strings(files)
    for lo, hi, m in ((1.0, 99999999.0, mask * 2), (2.0, 2.0, mask * 2), (1.0, 99999999.0, None)):
        ref = reference_minc2_model(R, Yq, dims, X, m, lo, hi)
        got = R.call("rmincgpu_minc2_model", F, rcall.NA_MATRIX, X, None, float(m is not None),
                     rcall.Strings([mfile] if m is not None else []), lo, hi, None, None, rcall.Strings(["lm"]))
        inside = np.ones(V, bool) if m is None else (m > lo - 0.5) & (m < hi + 0.5)
        assert_rows_match(got[inside], ref[inside], RTOL, f"minc2_model on files lo={lo}")
        assert (got[~inside] == 0).all()
    # other stored types take the real-valued route; two GPUs (when present) give the same matrix
    f16 = []
    for j in range(n):
        real = Y[:, j].reshape(dims)
        vox = np.round((real - real.min()) / (real.max() - real.min()) * 65535.0 - 32768.0).astype(np.int16)
        f = str(tmp_path / f"i{j:02d}.mnc")
        h5mini.write_minc2(f, vox, image_min=real.min(), image_max=real.max(), valid_range=(-32768, 32767), chunks=(3, 10, 9))
        f16.append(f)
    Y16 = np.asfortranarray(np.stack([api.default_reader(f) for f in f16], axis=1))
    ref = reference_minc2_model(R, Y16, dims, X, None, 1.0, 99999999.0)
    args = (rcall.Strings(f16), rcall.NA_MATRIX, X, None, 0.0, rcall.Strings([]), 1.0, 99999999.0, None, None, rcall.Strings(["lm"]))
    got = R.call("rmincgpu_minc2_model", *args)
    assert_rows_match(got, ref, RTOL, "minc2_model on int16 files")
    from rminc_b200 import core
    if core.device_count() >= 2:
        R.options(RMINC_GPU_N=2)
        try:
            np.testing.assert_array_equal(R.call("rmincgpu_minc2_model", *args), got)
            got2 = R.call("rmincgpu_minc2_model", F, rcall.NA_MATRIX, X, None, 0.0, rcall.Strings([]), 1.0, 99999999.0, None, None,
                          rcall.Strings(["lm"]))
        finally:
            R.options(RMINC_GPU_N=None)
        assert_rows_match(got2, reference_minc2_model(R, Yq, dims, X, None, 1.0, 99999999.0), RTOL, "2 GPUs, stored type")
    # per_voxel_anova with its own seven arguments on the same files, against the reference's routine on in-memory volumes
    asg = np.array([0, 1], dtype=np.int32)
    R.lib.rstub_clear_volumes()
    vfiles = R.register_volumes("subject_", Yq, dims)
    mcol = np.asfortranarray((mask * 2).reshape(-1, 1))
    mreg = R.register_volumes("mask_", mcol, dims)
    try:
        ref_an = R.call("per_voxel_anova", vfiles, X, asg, 1.0, mreg, 1.0, 99999999.0)
    finally:
        R.lib.rstub_clear_volumes()
    got_an = R.call("rmincgpu_per_voxel_anova", F, X, asg, 1.0, rcall.Strings([mfile]), 1.0, 99999999.0)
    assert_rows_match(got_an, ref_an, RTOL, "per_voxel_anova on files")
    # filenames_right: the per-voxel design [mmatrix | z_v] (src/modelling_functions.c:848-854, 1137-1143)
    gotc = R.call("rmincgpu_minc2_model", F, rcall.Strings(f16), X[:, :1].copy(), None, 0.0, rcall.Strings([]), 1.0, 99999999.0, None, None,
                  rcall.Strings(["lm"]))
    wantc = R.call("rmincgpu_lm_cov", Yq, Y16, X[:, :1].copy(), None, 1.0, 99999999.0)
    np.testing.assert_array_equal(gotc, wantc)


def test_error_path_is_an_r_error_with_the_references_message(R):
    """dpotri info > 0: the reference calls error("element (%d, %d) is zero, so the inverse cannot be computed")
    (src/modelling_functions.c:551-557); so does the shim, through Rf_error's longjmp, leaving nothing behind."""
    n = 20
    Y = np.asfortranarray(np.random.default_rng(2).normal(size=(64, n)))
    X = np.column_stack([np.ones(n), np.arange(n) % 2, np.zeros(n)])
    msg = r"element \(3, 3\) is zero, so the inverse cannot be computed"
    if R.has_reference:
        with pytest.raises(rcall.RError, match=msg):
            R.call("vertex_lm_loop", Y, rcall.NA_MATRIX, X)
    for _ in range(3):          # repeatedly: the unwinding must leave the protection stack and the library usable
        with pytest.raises(rcall.RError, match=msg):
            R.call("rmincgpu_vertex_lm_loop", Y, rcall.NA_MATRIX, X)
        with pytest.raises(rcall.RError, match=msg):
            R.call("rmincgpu_lm", Y, X, None, 1.0, 99999999.0, 1)
    ok = R.call("rmincgpu_vertex_lm_loop", Y, rcall.NA_MATRIX, X[:, :2].copy())
    assert ok.shape == (64, 7) and np.isfinite(ok).all()
    with pytest.raises(rcall.RError, match="model matrix has 19 rows, data has 20 subjects"):
        R.call("rmincgpu_vertex_lm_loop", Y, rcall.NA_MATRIX, X[:19, :2].copy())
    with pytest.raises(rcall.RError, match="device 99 out of range"):
        R.options(RMINC_GPU_DEVICE=99)
        try:
            R.call("rmincgpu_vertex_lm_loop", Y, rcall.NA_MATRIX, X[:, :2].copy())
        finally:
            R.options(RMINC_GPU_DEVICE=None)


def test_randomize_fdr_and_stored_routines(R, oracle):
    Y, X, mask = cfg1_data(dims=(10, 12, 9))
    V, n = Y.shape
    idx = perm_indices(n, 19)
    cols = np.array([5, 6], dtype=np.int32)                       # 1-based tvalue- columns of the 7-column result
    want = oracle.randomize(Y, X, idx, [4, 5], two_sided=True, mask=mask)
    from rminc_b200 import core
    for G in sorted({1, min(2, core.device_count())}):
        got = R.call("rmincgpu_randomize", Y, X, mask, 1.0, 99999999.0, one_based(idx), cols, True, G)
        np.testing.assert_allclose(got, want, rtol=RTOL, atol=0)
    got = R.call("rmincgpu_randomize", Y, X, None, 1.0, 99999999.0, one_based(idx), cols, False, 1)
    np.testing.assert_allclose(got, oracle.randomize(Y, X, idx, [4, 5], two_sided=False), rtol=RTOL, atol=0)
    Zc = np.asfortranarray(50 + 3 * np.random.default_rng(6).standard_normal(Y.shape))
    got = R.call("rmincgpu_randomize_cov", Y, Zc, X, mask, 1.0, 99999999.0, one_based(idx[:, :5]), np.array([6, 7], dtype=np.int32), True, 1)
    np.testing.assert_allclose(got, oracle.randomize_cov(Y, Zc, X, idx[:, :5], [5, 6], mask=mask), rtol=RTOL, atol=0)
    fit = oracle.minc2_model_lm(Y, (10, 12, 9), X, mask=mask)
    q, th = R.call("rmincgpu_fdr", fit[:, 5].copy(), 1, np.array([float(n - 2)]), mask)
    wq, wth, _ = oracle.fdr(fit[:, 5], "t", n - 2, mask=mask)
    np.testing.assert_allclose(q, wq, rtol=RTOL, atol=0)
    assert np.array_equal(np.isnan(th), np.isnan(wth))
    lam = np.arange(0, 19) * 0.05
    _, _, pv = oracle.fdr(fit[:, 5], "t", n - 2, mask=mask)
    frac = R.call("rmincgpu_fdr_pi0", fit[:, 5].copy(), 1, np.array([float(n - 2)]), mask, lam)
    np.testing.assert_allclose(frac, oracle.pi0_fractions(pv, lam), atol=1.5 / pv.size)
    for code, method in ((1, "BY"), (2, "pFDR"), (3, "qvalue")):
        q, th = R.call("rmincgpu_fdr_method", fit[:, 5].copy(), 1, np.array([float(n - 2)]), mask, code, 0.9)
        np.testing.assert_allclose(q, oracle.fdr(fit[:, 5], "t", n - 2, mask=mask, method=method, pi0=0.9)[0], rtol=RTOL, atol=0)
    # volumes in their stored type: a raw vector of uint16 samples + the real-range tables
    rng = np.random.default_rng(5)
    sl = 12 * 9
    U = np.asfortranarray(np.round(rng.normal(30000, 400, (V, n))).astype(np.uint16))
    scale, offset = rng.uniform(1.5, 2.5, (10, n)) / 65535.0, rng.normal(-0.8, 0.05, (10, n))
    Yx = oracle.expand_stored(U, scale, offset, slice_len=sl)
    raw = rcall.Raw(U.tobytes(order="F"))
    got = R.call("rmincgpu_lm_stored", raw, 1, float(V), scale, offset, 0.0, float(sl), X, mask, 1.0, 99999999.0)
    assert_rows_match(got, oracle.minc2_model_lm(Yx, (10, 12, 9), X, mask=mask), RTOL, "lm_stored")
    got = R.call("rmincgpu_randomize_stored", raw, 1, float(V), scale, offset, 0.0, float(sl), X, mask, 1.0, 99999999.0,
                 one_based(idx), cols, True, 1)
    np.testing.assert_allclose(got, oracle.randomize(Yx, X, idx, [4, 5], mask=mask), rtol=RTOL, atol=0)
    Z = np.asfortranarray(50 + 4 * rng.normal(size=(V, n)))
    got = R.call("rmincgpu_lm_cov", Y, Z, X, mask, 1.0, 99999999.0)
    assert_rows_match(got, oracle.lm_loop_cov(Y, Z, X, mask=mask), RTOL, "lm_cov")


def test_tfce_routines(R, oracle):
    from helpers import tfce_randomize_oracle
    rng = np.random.default_rng(6)
    a, b, n = 12, 14, 16
    V = a * b
    adj = grid_adjacency(a, b)
    adj_sexp = [np.asarray(x, dtype=np.int32) for x in adj]
    X = np.column_stack([np.ones(n), np.arange(n) % 2])
    Y = np.asfortranarray(rng.normal(size=(V, n)) + 0.8 * np.outer(rng.random(V) > 0.7, X[:, 1]))
    ptr, aidx = oracle.adjacency_csr(adj)
    m = rng.normal(size=V)
    got = R.call("rmincgpu_graph_tfce", m, adj_sexp, 0.5, 2.0, 30, np.ones(V), 0)
    wg = oracle.graph_tfce(m, ptr, aidx, nsteps=30)
    np.testing.assert_allclose(got, wg, rtol=1e-9, atol=1e-10 * np.abs(wg).max())
    idx = perm_indices(n, 5)
    cols = np.array([5, 6], dtype=np.int32)
    got = R.call("rmincgpu_randomize_tfce", Y, X, one_based(idx), cols, True, adj_sexp, np.ones(V), 0.5, 2.0, 30, 0)
    np.testing.assert_allclose(got, tfce_randomize_oracle(oracle, Y, X, idx, [4, 5], adj, nsteps=30), rtol=1e-9, atol=1e-12)
    dims = np.array([6.0, 7.0, 4.0])
    mv = rng.normal(size=168)
    got = R.call("rmincgpu_volume_tfce", mv, dims, 6, 1.0, 0.2, 0.5, 2.0, 0)
    wvv = oracle.volume_tfce(mv, (6, 7, 4), d=0.2)
    np.testing.assert_allclose(got, wvv, rtol=1e-9, atol=1e-10 * np.abs(wvv).max())
    Yv = np.asfortranarray(Y[:168])
    got = R.call("rmincgpu_randomize_volume_tfce", Yv, X, None, 1.0, 99999999.0, one_based(idx), cols, True, dims, 6, 1.0, 0.3,
                 0.5, 2.0, 0, 1)
    np.testing.assert_allclose(got, oracle.randomize_volume_tfce(Yv, X, idx, [4, 5], (6, 7, 4), d=0.3), rtol=1e-9, atol=1e-12)


def test_dataset_routines_and_the_finalizer(R, oracle):
    from rminc_b200 import core
    Y, X, mask = cfg1_data(dims=(10, 12, 9))
    V, n = Y.shape
    idx = perm_indices(n, 11)
    cols = np.array([5, 6], dtype=np.int32)
    want = oracle.minc2_model_lm(Y, (10, 12, 9), X, mask=mask)
    for G in sorted({1, min(2, core.device_count())}):
        h = R.call("rmincgpu_dataset_create", Y, mask, 1.0, 99999999.0, G)
        assert isinstance(h, rcall.Handle)
        assert_rows_match(R.call("rmincgpu_dataset_lm", h, X), want, RTOL, "dataset_lm")
        np.testing.assert_allclose(R.call("rmincgpu_dataset_randomize", h, X, one_based(idx), cols, True),
                                   oracle.randomize(Y, X, idx, [4, 5], mask=mask), rtol=RTOL, atol=0)
        q, th = R.call("rmincgpu_dataset_fdr", h, 6, 1, np.array([float(n - 2)]))
        np.testing.assert_allclose(q, oracle.fdr(want[:, 5], "t", n - 2, mask=mask)[0], rtol=RTOL, atol=0)
        qb, _ = R.call("rmincgpu_dataset_fdr_method", h, 6, 1, np.array([float(n - 2)]), 1, 1.0)          # "BY"
        np.testing.assert_allclose(qb, oracle.fdr(want[:, 5], "t", n - 2, mask=mask, method="BY")[0], rtol=RTOL, atol=0)
        a = R.call("rmincgpu_dataset_anova", h, X, np.array([0, 1], dtype=np.int32))
        assert_rows_match(a, core.anova_fit(Y, X, [0, 1], mask=mask), RTOL, "dataset_anova")
        if G == 1:
            assert R.call("rmincgpu_dataset_free", h) is None
            with pytest.raises(rcall.RError, match="has been freed"):
                R.call("rmincgpu_dataset_lm", h, X)
        R.release(h)            # the collector: runs the finalizer (a no-op after an explicit free)
