"""Parity of the CUDA path against the CPU oracle, through the C ABI, on a real B200.

Tolerance (BASELINE.json north_star): masked voxels and output layout bit-exact; betas, F, t,
R-squared and logLik within 1e-9 relative in FP64; rank-deficiency behaviour identical.
"""
import os

import numpy as np
import pytest

from helpers import assert_rows_match, cfg1_data, design_cfg1, design_cfg2, design_cov, ellipsoid_mask

pytestmark = pytest.mark.gpu
RTOL = 1e-9
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def core(lib):
    from rminc_b200 import core as c
    assert c.device_count() > 0, "GPU tests need a CUDA device"
    return c


@pytest.fixture(params=["tma", "ldg"])
def variant(request, monkeypatch):
    monkeypatch.setenv("RMG_FIT_VARIANT", request.param)
    return request.param


def test_golden_vectors_of_the_reference(core, variant):
    g = np.load(os.path.join(GOLD, "ref_golden.npz"))
    assert_rows_match(core.vertex_lm(g["A_Y"], g["A_X"]), g["A_out"], RTOL, "A")
    for key, kw in (("B_out_default", {}), ("B_out_maskval2", dict(mask_lo=2, mask_hi=2))):
        out = core.lm_fit(g["B_Y"], g["B_X"], mask=g["B_mask"], **kw)
        assert_rows_match(out, g[key], RTOL, key)
        zero = g[key][:, 0] == 0
        assert (out[zero] == 0).all() and not np.signbit(out[zero]).any()       # masked rows: +0.0 bit-exact
    assert_rows_match(core.lm_fit(g["B_Y"], g["B_X"]), g["B_out_nomask"], RTOL, "B nomask")
    # rank-deficient design: same pivoted layout, trailing zero beta, t ~ 0 in the coupled slots
    out, want = core.vertex_lm(g["C_Y"], g["C_X"]), g["C_out"]
    assert_rows_match(out[:, [0, 1, 2, 3, 4, 10]], want[:, [0, 1, 2, 3, 4, 10]], RTOL, "C")
    assert (out[:, 5] == 0).all() and (out[:, 9] == 0).all()
    assert np.all(np.abs(out[:, [6, 7, 9]]) < 1e-6) and np.all(np.abs(want[:, [6, 7, 9]]) < 1e-6)
    assert_rows_match(out[:, 8], want[:, 8], RTOL, "C t_age")


def test_config1_full_parity(core, oracle, variant):
    """BASELINE config 1: mincLm(y ~ group), 20 volumes of 60x80x50, masked."""
    Y, X, mask = cfg1_data()
    want = oracle.minc2_model_lm(Y, (60, 80, 50), X, mask=mask)
    got = core.lm_fit(Y, X, mask=mask)
    inside = mask > 0.5
    assert 0.3 < inside.mean() < 0.45
    assert (got[~inside] == 0).all() and not np.signbit(got[~inside]).any()
    assert_rows_match(got, want, RTOL, "cfg1 masked")
    # integer labels + maskval exercises the +-0.5 interval rule
    labels = (np.arange(Y.shape[0]) % 4).astype(float)
    assert_rows_match(core.lm_fit(Y, X, mask=labels, mask_lo=2, mask_hi=2),
                      oracle.minc2_model_lm(Y, (60, 80, 50), X, mask=labels, lo=2, hi=2), RTOL, "maskval")


def test_degenerate_voxels_match_by_class(core, oracle, variant):
    Y, X, _ = cfg1_data(dims=(20, 20, 20))
    Y = Y.copy(order="F")
    Y[::7] = 0.0                                  # all-zero voxels: logLik=+Inf, F/R2/t=NaN, beta=0
    got, want = core.lm_fit(Y, X), oracle.vertex_lm_loop(Y, X)
    assert np.isposinf(got[::7, -1]).all() and np.isnan(got[::7, 0]).all() and (got[::7, 2:4] == 0).all()
    assert_rows_match(got, want, RTOL, "degenerate")


@pytest.mark.parametrize("mean,sd", [(0.0, 0.2), (100.0, 10.0), (1000.0, 1.0), (30000.0, 50.0), (-5e4, 3.0)])
def test_large_mean_small_sd_keeps_1e9(core, oracle, variant, mean, sd):
    """Hard part H1: one-pass rss must not cancel when |mean| >> sd (ushort-scaled intensities)."""
    rng = np.random.default_rng(5)
    V, n = 4096, 500
    X = design_cfg2(n)
    Y = np.asfortranarray(mean + sd * rng.standard_normal((V, n)) + 0.5 * sd * np.outer(rng.random(V), X[:, 1]))
    assert_rows_match(core.lm_fit(Y, X), oracle.vertex_lm_loop(Y, X), RTOL, f"mean={mean} sd={sd}")


def test_near_perfect_fits_take_the_exact_path(core, oracle, variant):
    """rss/|y'|^2 < 1e-3: the kernel recomputes rss/mss from explicit residuals."""
    rng = np.random.default_rng(6)
    V, n = 2048, 120
    X = design_cov(n, 4)
    B = rng.normal(size=(V, 4)) * 50
    # residuals are computed as y - fitted with |fitted| ~ 100: their absolute accuracy is ~1e-14 in
    # the reference and here alike, so the attainable agreement on rss is ~1e-14/noise, not 1e-9
    for noise, tol in ((1e-1, RTOL), (1e-4, RTOL), (1e-7, 1e-5)):
        Y = np.asfortranarray(B @ X.T + noise * rng.standard_normal((V, n)))
        assert_rows_match(core.lm_fit(Y, X), oracle.vertex_lm_loop(Y, X), tol, f"noise={noise}")


@pytest.mark.parametrize("p", [1, 2, 3, 5, 6, 7, 8, 9, 12, 13, 16, 20, 32])
def test_every_padded_width(core, oracle, p):
    rng = np.random.default_rng(1000 + p)
    n, V = 64, 1026
    X = design_cov(n, p, seed=77) if p > 1 else np.ones((n, 1))
    Y = np.asfortranarray(rng.normal(3, 1, (V, n)))
    got, want = core.lm_fit(Y, X), oracle.vertex_lm_loop(Y, X)
    if p == 1:
        # intercept only: mss is exactly 0 up to rounding noise, so F = (mss/0)/resvar and R2 are
        # Inf/NaN/1e-33 noise in the reference; compare beta, t, logLik
        assert np.all(np.abs(got[:, 1]) < 1e-25) and np.all(np.abs(want[:, 1]) < 1e-25)
        got, want = got[:, 2:], want[:, 2:]
    assert_rows_match(got, want, RTOL, f"p={p}")


def test_no_intercept_design(core, oracle, variant):
    rng = np.random.default_rng(8)
    n, V = 80, 3000
    X = rng.normal(size=(n, 3))
    Y = np.asfortranarray(rng.normal(10, 1, (V, n)))
    assert_rows_match(core.lm_fit(Y, X), oracle.vertex_lm_loop(Y, X), RTOL, "no intercept")


def test_ragged_shapes_and_leading_dimensions(core, oracle):
    rng = np.random.default_rng(9)
    X = design_cfg2(33)
    for V in (1, 2, 3, 31, 255, 513, 1025, 4099):
        Y = np.asfortranarray(rng.normal(0, 1, (V, 33)))
        m = (rng.random(V) > 0.3).astype(float)
        assert_rows_match(core.lm_fit(Y, X, mask=m), oracle.minc2_model_lm(Y, (1, 1, V), X, mask=m), RTOL, f"V={V}")
    assert core.lm_fit(np.zeros((0, 33), order="F"), X).shape == (0, 11)
    # everything masked out
    Y = np.asfortranarray(rng.normal(size=(777, 33)))
    out = core.lm_fit(Y, X, mask=np.zeros(777))
    assert (out == 0).all() and not np.signbit(out).any()


def test_split_n_for_few_vertices_many_subjects(core, oracle, monkeypatch):
    """vertexLm shape (config 3 scaled down): small V, n = 2000, p = 7 -> subject axis is split."""
    rng = np.random.default_rng(10)
    V, n = 1500, 2000
    X = design_cov(n, 7)
    Y = np.asfortranarray(rng.normal(2.5, 0.4, (V, n)))
    want = oracle.vertex_lm_loop(Y, X)
    for split in ("0", "1", "2", "4", "8"):
        monkeypatch.setenv("RMG_FIT_VARIANT", "ldg")
        monkeypatch.setenv("RMG_FIT_SPLIT", split)
        assert_rows_match(core.vertex_lm(Y, X), want, RTOL, f"split={split}")


def test_chunked_host_pipeline(core, oracle, monkeypatch):
    """Host entry point with several voxel chunks in flight (ragged last chunk)."""
    Y, X, mask = cfg1_data(dims=(30, 40, 25))
    want = oracle.minc2_model_lm(Y, (30, 40, 25), X, mask=mask)
    monkeypatch.setenv("RMG_CHUNK_VOXELS", "7000")
    assert_rows_match(core.lm_fit(Y, X, mask=mask), want, RTOL, "chunked")
    assert core.last_kernel_ms() > 0
    assert_rows_match(core.lm_fit(Y, X, mask=mask, n_gpus=1), want, RTOL, "multi(1)")


def test_device_resident_entry_point(core, oracle, variant):
    import torch
    Y, X, mask = cfg1_data(dims=(24, 32, 20))
    want = oracle.minc2_model_lm(Y, (24, 32, 20), X, mask=mask)
    Yt = torch.from_numpy(np.ascontiguousarray(Y.T)).cuda()        # (n, V) row-major == V x n column-major
    mt = torch.from_numpy(mask).cuda()
    before = core.launch_count()
    out = core.lm_fit_torch(Yt, X, mask=mt)
    torch.cuda.synchronize()
    assert core.launch_count() > before
    assert_rows_match(out.cpu().numpy().T, want, RTOL, "device")
    out2 = core.lm_fit_torch(Yt, X)
    torch.cuda.synchronize()
    assert_rows_match(out2.cpu().numpy().T, oracle.vertex_lm_loop(Y, X), RTOL, "device nomask")


def test_singular_design_error_contract(core):
    n = 30
    X = np.column_stack([np.ones(n), np.arange(n) % 2, np.zeros(n)])
    with pytest.raises(core.SingularDesignError, match=r"element \(3, 3\) is zero, so the inverse cannot be computed"):
        core.lm_fit(np.ones((10, n)), X)


# ------------------------------------------------------------------ mincRandomize
@pytest.fixture(params=["refit", "gemm"])
def perm_path(request, monkeypatch):
    monkeypatch.setenv("RMG_PERM_PATH", request.param)
    return request.param


def perm_indices(n, R, seed=4, replace=False):
    rng = np.random.default_rng(seed)
    if replace:
        return np.column_stack([rng.integers(0, n, n) for _ in range(R)])
    return np.column_stack([rng.permutation(n) for _ in range(R)])


def test_randomize_matches_oracle(core, oracle, perm_path):
    Y, X, mask = cfg1_data(dims=(16, 20, 12))
    n = X.shape[0]
    idx = perm_indices(n, 40)
    tcols = [4, 5]
    for two_sided in (True, False):
        for m in (mask, None):
            want = oracle.randomize(Y, X, idx, tcols, two_sided=two_sided, mask=m)
            got = core.randomize(Y, X, idx, tcols, two_sided=two_sided, mask=m)
            assert got.shape == (40, 2)
            np.testing.assert_allclose(got, want, rtol=RTOL, atol=0, err_msg=f"{perm_path} two_sided={two_sided}")


def test_randomize_p4_high_mean_and_degenerate_rows(core, oracle, perm_path):
    rng = np.random.default_rng(12)
    V, n = 3000, 60
    X = design_cfg2(n)
    Y = np.asfortranarray(30000 + 50 * rng.standard_normal((V, n)))
    Y[::11] = 0.0            # non-finite statistics -> 0
    Y[5::13] = 123.0         # constant non-zero voxels: rounding-noise rss in the reference (H2): masked out
    mask = np.ones(V)
    mask[5::13] = 0
    idx = perm_indices(n, 70)
    cols = [6, 7, 8, 9]
    want = oracle.randomize(Y, X, idx, cols, two_sided=True, mask=mask)
    got = core.randomize(Y, X, idx, cols, two_sided=True, mask=mask)
    np.testing.assert_allclose(got, want, rtol=RTOL)
    got1 = core.randomize(Y, X, idx, cols, two_sided=False, mask=mask)
    np.testing.assert_allclose(got1, oracle.randomize(Y, X, idx, cols, two_sided=False, mask=mask), rtol=RTOL)
    assert (got1 >= 0).all()


def test_randomize_with_replacement_and_other_columns(core, oracle):
    """replace=TRUE index vectors; non-t columns go through the refit path."""
    Y, X, mask = cfg1_data(dims=(10, 12, 10))
    n = X.shape[0]
    idx = perm_indices(n, 25, replace=True)
    keep = [r for r in range(idx.shape[1]) if len(set(X[idx[:, r], 1])) == 2][:20]
    idx = idx[:, keep]
    cols = [0, 1, 3, 5, 6]          # F, R2, beta, t, logLik
    want = oracle.randomize(Y, X, idx, cols, two_sided=True, mask=mask)
    got = core.randomize(Y, X, idx, cols, two_sided=True, mask=mask)
    np.testing.assert_allclose(got, want, rtol=RTOL)
    # a resample that empties a factor level makes the refit singular -> the whole call fails
    bad = idx.copy()
    bad[:, 3] = np.flatnonzero(X[:, 1] == 0)[0]
    with pytest.raises(core.SingularDesignError, match="permutation 4"):
        core.randomize(Y, X, bad, [4, 5], mask=mask)


def test_randomize_with_replacement_default_t_columns_both_paths(core, oracle, perm_path):
    """replace = TRUE index vectors with the default tvalue- columns (R/minc_voxel_statistics.R:1470-1492): every such
    permutation changes the column space, so the GEMM path factors it from scratch and packs its own Q (and, when a
    resample drops the balance, its own constant column) -- the branch true permutations never take."""
    rng = np.random.default_rng(5)
    V, n = 2500, 40
    X = design_cfg2(n)
    p = X.shape[1]
    Y = np.asfortranarray(5 + 0.2 * rng.standard_normal((V, n)) + 0.1 * np.outer(rng.random(V), X[:, 1]))
    mask = (np.arange(V) % 7 > 0).astype(float)
    idx = np.column_stack([perm_indices(n, 30, seed=9, replace=True), perm_indices(n, 6, seed=10)])   # mixed
    keep = [r for r in range(idx.shape[1]) if np.linalg.matrix_rank(X[idx[:, r]]) == p]
    idx = idx[:, keep]
    assert idx.shape[1] >= 30
    cols = list(range(p + 2, 2 * p + 2))
    for two_sided in (True, False):
        want = oracle.randomize(Y, X, idx, cols, two_sided=two_sided, mask=mask)
        got = core.randomize(Y, X, idx, cols, two_sided=two_sided, mask=mask)
        np.testing.assert_allclose(got, want, rtol=RTOL, atol=0, err_msg=f"{perm_path} two_sided={two_sided}")


def test_randomize_rank_deficient_and_no_intercept_designs(core, oracle, perm_path):
    """a rank-deficient base design [1, g, 1-g, age] (pivot [0,1,3,2], rank 3) and a design without an
    intercept, through rmg_randomize on both paths."""
    rng = np.random.default_rng(6)
    V, n = 1500, 36
    g = (np.arange(n) % 2).astype(float)
    age = np.round(rng.normal(60, 10, n), 1)
    Y = np.asfortranarray(2 + 0.3 * rng.standard_normal((V, n)) + 0.01 * np.outer(rng.random(V), age))
    idx = perm_indices(n, 17, seed=11)
    Xd = np.column_stack([np.ones(n), g, 1 - g, age])
    cols = [6, 7, 8, 9]
    want = oracle.randomize(Y, Xd, idx, cols)
    got = core.randomize(Y, Xd, idx, cols)
    # beta stays in pivoted order while se is un-pivoted (Q2): the age coefficient (slot 2) is divided by the aliased
    # column's huge se, slot 3's zero coefficient by the finite one -- every t is rounding noise or an exact 0, in the
    # reference (|t| ~ 1e-14) as here; "identical" means the same ~0 pattern
    assert got.shape == want.shape
    assert (np.abs(want) < 1e-6).all() and (np.abs(got) < 1e-6).all(), perm_path
    assert (got[:, 3] == 0).all() and (want[:, 3] == 0).all()
    # no intercept: nothing is centred, nothing is skipped
    X0 = np.column_stack([g, age / 60.0])
    for two_sided in (True, False):
        np.testing.assert_allclose(core.randomize(Y, X0, idx, [4, 5], two_sided=two_sided),
                                   oracle.randomize(Y, X0, idx, [4, 5], two_sided=two_sided), rtol=RTOL, atol=0,
                                   err_msg=f"{perm_path} no intercept two_sided={two_sided}")
    # ... and with a mask, where the 0 floor of masked rows takes part in a one-sided maximum (Q6)
    mask = (np.arange(V) % 3 > 0).astype(float)
    np.testing.assert_allclose(core.randomize(Y, X0, idx, [4, 5], two_sided=False, mask=mask),
                               oracle.randomize(Y, X0, idx, [4, 5], two_sided=False, mask=mask), rtol=RTOL, atol=0)


def test_randomize_gemm_equals_refit_on_larger_case(core, monkeypatch):
    rng = np.random.default_rng(13)
    V, n = 20000, 100
    X = design_cfg2(n)
    Y = np.asfortranarray(rng.normal(0, 0.2, (V, n)))
    mask = ellipsoid_mask((20, 25, 40))
    idx = perm_indices(n, 130)
    cols = [6, 7, 8, 9]
    monkeypatch.setenv("RMG_PERM_PATH", "refit")
    a = core.randomize(Y, X, idx, cols, mask=mask)
    monkeypatch.setenv("RMG_PERM_PATH", "gemm")
    b = core.randomize(Y, X, idx, cols, mask=mask)
    np.testing.assert_allclose(b, a, rtol=1e-10)


def test_randomize_voxel_blocks_and_staging_ring(core, oracle, monkeypatch):
    """rmg_randomize uploads Y in voxel blocks and runs every block's permutations as it lands; pageable input goes
    through the pinned staging ring.  Block boundaries (incl. an odd tail that takes the refit path) change nothing."""
    Y, X, mask = cfg1_data(dims=(16, 20, 13))           # V = 4160
    Y = np.asfortranarray(Y[:4159])                     # odd V: the last block is odd
    mask = mask[:4159]
    n = X.shape[0]
    idx = perm_indices(n, 33)
    cols = [4, 5]
    want = oracle.randomize(Y, X, idx, cols, two_sided=True, mask=mask)
    for chunk, rows in (("1024", None), ("700", "3"), ("2048", "20"), ("100000", None)):
        monkeypatch.setenv("RMG_PERM_CHUNK_VOXELS", chunk)
        if rows:
            monkeypatch.setenv("RMG_STAGING_ROWS", rows)
        else:
            monkeypatch.delenv("RMG_STAGING_ROWS", raising=False)
        got = core.randomize(Y, X, idx, cols, two_sided=True, mask=mask)
        np.testing.assert_allclose(got, want, rtol=RTOL, atol=0, err_msg=f"chunk={chunk} rows={rows}")
    got = core.randomize(Y, X, idx, cols, two_sided=False, mask=None)
    np.testing.assert_allclose(got, oracle.randomize(Y, X, idx, cols, two_sided=False, mask=None), rtol=RTOL)
    # a sub-matrix (ldY > V) as input
    monkeypatch.setenv("RMG_PERM_CHUNK_VOXELS", "1500")
    import ctypes
    from rminc_b200 import _lib
    Yb = np.asfortranarray(np.vstack([Y, np.full((77, n), 9.0)]))
    dist = np.empty((33, 2), order="F")
    ix = np.asfortranarray(idx, dtype=np.int32)
    cc = np.asarray(cols, dtype=np.int32)
    Xf = np.asfortranarray(X)
    err = _lib.errbuf()
    rc = _lib.load().rmg_randomize(Yb.ctypes.data, 4159, Yb.shape[0], n, Xf.ctypes.data, X.shape[1], mask.ctypes.data, 0.5,
                                   99999999.5, ix.ctypes.data, 33, cc.ctypes.data, 2, 1, dist.ctypes.data, 0, err, len(err))
    assert rc == 0, err.value
    np.testing.assert_allclose(dist, want, rtol=RTOL)


def test_randomize_on_stored_volumes(core, oracle, monkeypatch):
    """rmg_randomize_stored: uint16 / float32 samples expanded once on the device == randomize on the reader's doubles."""
    U, X, scale, offset = _stored_case(5000, 48, 1000, seed=8)
    V, n = U.shape
    Y = oracle.expand_stored(U, scale, offset, voxel_min=100.0, slice_len=1000)
    mask = (np.arange(V) % 5 > 0).astype(float)
    idx = perm_indices(n, 21)
    p = X.shape[1]
    cols = list(range(p + 2, 2 * p + 2))
    want = oracle.randomize(Y, X, idx, cols, two_sided=True, mask=mask)
    for chunk, rows in ((None, None), ("1400", None), ("1024", "5")):
        for k, v in (("RMG_PERM_CHUNK_VOXELS", chunk), ("RMG_STAGING_ROWS", rows)):
            if v:
                monkeypatch.setenv(k, v)
            else:
                monkeypatch.delenv(k, raising=False)
        got = core.randomize_stored(U, X, idx, cols, scale=scale, offset=offset, voxel_min=100.0, slice_len=1000, mask=mask)
        np.testing.assert_allclose(got, want, rtol=RTOL, atol=0, err_msg=f"chunk={chunk} rows={rows}")
        # and exactly what the double path gives on the expanded matrix (same doubles, same kernels)
        np.testing.assert_array_equal(got, core.randomize(Y, X, idx, cols, mask=mask))
    got2 = core.randomize_stored(U, X, idx, cols, scale=scale, offset=offset, voxel_min=100.0, slice_len=1000, mask=mask,
                                 n_gpus=min(2, core.device_count()))      # voxel shards keep their global slice index
    np.testing.assert_allclose(got2, want, rtol=RTOL)
    F = np.asfortranarray(np.random.default_rng(4).normal(0, 0.2, (3000, 48)).astype(np.float32))
    monkeypatch.setenv("RMG_PERM_CHUNK_VOXELS", "1024")
    np.testing.assert_allclose(core.randomize_stored(F, X, idx, cols, two_sided=False),
                               oracle.randomize(F.astype(np.float64), X, idx, cols, two_sided=False), rtol=RTOL)
    with pytest.raises(core.RmincGpuError, match="voxel_min must be an integer"):
        core.randomize_stored(U, X, idx, cols, scale=scale, offset=offset, voxel_min=0.5, slice_len=1000)


def test_randomize_device_entry_point(core, oracle):
    import torch
    Y, X, mask = cfg1_data(dims=(12, 16, 10))
    n = X.shape[0]
    idx = perm_indices(n, 16)
    Yt = torch.from_numpy(np.ascontiguousarray(Y.T)).cuda()
    mt = torch.from_numpy(mask).cuda()
    d = core.randomize_torch(Yt, X, idx, [4, 5], two_sided=True, mask=mt)
    torch.cuda.synchronize()
    np.testing.assert_allclose(d.cpu().numpy().T, oracle.randomize(Y, X, idx, [4, 5], mask=mask), rtol=RTOL)


# ------------------------------------------------------------------ anova epilogue (scope row f-1)
def test_anova_matches_oracle(core, oracle, variant):
    g = np.load(os.path.join(GOLD, "ref_golden.npz"))
    out = core.anova_fit(g["D_Y"], g["D_X"], g["D_asgn"], mask=g["D_mask"])
    assert_rows_match(out, g["D_out"], RTOL, "anova golden")
    assert (out[g["D_mask"] < 0.5] == 0).all()
    rng = np.random.default_rng(31)
    n, V = 60, 5000
    grp = np.arange(n) % 3
    age = np.round(rng.normal(60, 10, n), 1)
    X = np.column_stack([np.ones(n), grp == 1, grp == 2, age, (grp == 1) * age, (grp == 2) * age]).astype(float)
    asgn = [0, 1, 1, 2, 3, 3]
    Y = np.asfortranarray(30000 + 50 * rng.standard_normal((V, n)) + 30 * np.outer(rng.random(V), X[:, 1]))
    mask = (rng.random(V) > 0.3).astype(float)
    assert_rows_match(core.anova_fit(Y, X, asgn, mask=mask), oracle.anova(Y, X, asgn, mask=mask), RTOL, "anova")
    assert_rows_match(core.anova_fit(Y, X, asgn), oracle.anova(Y, X, asgn), RTOL, "anova nomask")
    # single-term and intercept-free designs; one term spanning all columns
    X2 = np.column_stack([np.ones(n), age])
    assert_rows_match(core.anova_fit(Y, X2, [0, 1]), oracle.anova(Y, X2, [0, 1]), RTOL, "anova p=2")
    X3 = np.column_stack([grp == 0, grp == 1, grp == 2]).astype(float)
    assert_rows_match(core.anova_fit(Y, X3, [1, 1, 1]), oracle.anova(Y, X3, [1, 1, 1]), RTOL, "anova no intercept")
    with pytest.raises(core.RmincGpuError, match="full-rank"):
        core.anova_fit(Y, np.column_stack([np.ones(n), grp == 1, grp != 1, age]).astype(float), [0, 1, 1, 2])


def test_anova_device_entry_point(core, oracle):
    import torch
    rng = np.random.default_rng(32)
    n, V = 40, 4096
    X = design_cov(n, 5)
    asgn = [0, 1, 2, 2, 3]
    Y = rng.normal(1, 0.5, (V, n))
    Yt = torch.from_numpy(np.ascontiguousarray(Y.T)).cuda()
    out = core.anova_fit_torch(Yt, X, asgn)
    torch.cuda.synchronize()
    assert_rows_match(out.cpu().numpy().T, oracle.anova(Y, X, asgn), RTOL, "anova device")


# ------------------------------------------------------------------ weighted least squares (scope row f-2)
def test_weighted_fit_matches_oracle(core, oracle):
    rng = np.random.default_rng(51)
    for n, p, V in [(25, 3, 40), (120, 5, 3001), (500, 4, 2048)]:
        X = design_cov(n, p, seed=p)
        w = rng.uniform(0.2, 3.0, n)
        Y = np.asfortranarray(rng.normal(1000, 5, (V, n)) + 3 * np.outer(rng.random(V), X[:, 1]))
        Y[V // 2] = 0.0
        assert_rows_match(core.vertex_wlm(Y, X, w), oracle.vertex_wlm_loop(Y, X, w), RTOL, f"wlm n={n}")
    # unit weights reproduce the unweighted fit; no-intercept weighted design
    assert_rows_match(core.vertex_wlm(Y, X, np.ones(n)), oracle.vertex_lm_loop(Y, X), RTOL, "unit weights")
    Xn = rng.normal(size=(n, 3))
    assert_rows_match(core.vertex_wlm(Y, Xn, w), oracle.vertex_wlm_loop(Y, Xn, w), RTOL, "wlm no intercept")
    with pytest.raises(core.RmincGpuError, match="weights must be finite and > 0"):
        core.vertex_wlm(Y, X, np.r_[0.0, np.ones(n - 1)])
    with pytest.raises(ValueError, match="Incorrect number of weights"):
        core.vertex_wlm(Y, X, np.ones(n - 1))


# ------------------------------------------------------------------ voxel-wise covariate designs
def _cov_case(V, n, ps, seed, zmean=100.0, zsd=3.0):
    rng = np.random.default_rng(seed)
    Xs = design_cov(n, ps, seed=seed + 1000) if ps else None     # (a different stream than the data's)
    Y = np.asfortranarray(rng.normal(2.0, 0.5, (V, n)))
    Z = np.asfortranarray(rng.normal(zmean, zsd, (V, n)))
    Y += 0.02 * (Z - zmean) * rng.random((V, 1))
    return Y, Z, Xs


def test_covariate_golden_vectors(core):
    g = np.load(os.path.join(GOLD, "ref_golden.npz"))
    for key, Xs in (("E_out_static", g["E_Xs"]), ("E_out_nostatic", None)):
        got, want = core.lm_fit_cov(g["E_Y"], g["E_Z"], Xs), g[key]
        keep = np.ones(len(want), bool)
        keep[7] = False                      # vertex 7: constant covariate, checked by class below
        assert_rows_match(got[keep], want[keep], RTOL, key)
        p = (want.shape[1] - 3) // 2
        # rank p-1 at vertex 7: F, R2, betas (0 in the covariate's slot), logLik match; t = 0 in its slot
        cols = [0, 1] + list(range(2, 2 + p)) + [2 * p + 2]
        assert_rows_match(got[7, cols], want[7, cols], RTOL, key + " v7")
        assert got[7, 1 + p] == 0.0 and got[7, 1 + 2 * p] == 0.0 and want[7, 1 + 2 * p] == 0.0


@pytest.mark.parametrize("ps", [0, 1, 2, 3, 4, 6, 8, 11, 16])
def test_covariate_matches_oracle_every_width(core, oracle, ps):
    n = 64 if ps <= 8 else 96
    Y, Z, Xs = _cov_case(1538, n, ps, seed=100 + ps)
    assert_rows_match(core.lm_fit_cov(Y, Z, Xs), oracle.lm_loop_cov(Y, Z, Xs), RTOL, f"ps={ps}")


@pytest.mark.parametrize("zmean,zsd", [(0.0, 1.0), (1000.0, 1.0), (30000.0, 50.0), (-5e4, 3.0)])
def test_covariate_large_mean_small_sd(core, oracle, zmean, zsd):
    """Both streamed operands carry ushort-like offsets: the first-sample shift keeps d^2 and rss at 1e-9."""
    rng = np.random.default_rng(9)
    V, n = 2048, 500
    Xs = design_cfg2(n)
    Y = np.asfortranarray(30000 + 50 * rng.standard_normal((V, n)) + 20 * np.outer(rng.random(V), Xs[:, 1]))
    Z = np.asfortranarray(zmean + zsd * rng.standard_normal((V, n)))
    assert_rows_match(core.lm_fit_cov(Y, Z, Xs), oracle.lm_loop_cov(Y, Z, Xs), RTOL, f"z {zmean} {zsd}")


def test_covariate_near_perfect_fits_and_near_collinear_covariates(core, oracle):
    """d^2/|z'|^2 or rss/|y'|^2 below 1e-3: the kernel recomputes them from explicit residuals (exact_cov)."""
    rng = np.random.default_rng(16)
    V, n = 1024, 120
    Xs = design_cov(n, 3, seed=77)
    B = rng.normal(size=(V, 3)) * 20
    for noise, tol in ((1e-1, RTOL), (1e-4, RTOL), (1e-6, 1e-6)):
        Z = np.asfortranarray(rng.normal(5.0, 1.0, (V, n)))
        Y = np.asfortranarray(B @ Xs.T + 3.0 * Z + noise * rng.standard_normal((V, n)))
        assert_rows_match(core.lm_fit_cov(Y, Z, Xs), oracle.lm_loop_cov(Y, Z, Xs), tol, f"y noise={noise}")
    # covariate almost inside span(Xs), but above dqrdc2's 1e-7 cut (|z_perp|/|z| ~ noise/35): d^2 << |z|^2
    for noise, tol in ((1e-1, RTOL), (1e-3, RTOL), (1e-4, 1e-8)):
        Zc = np.asfortranarray(B @ Xs.T + noise * rng.standard_normal((V, n)))
        Yc = np.asfortranarray(rng.normal(2.0, 0.5, (V, n)))
        assert_rows_match(core.lm_fit_cov(Yc, Zc, Xs), oracle.lm_loop_cov(Yc, Zc, Xs), tol, f"z noise={noise}")


def test_covariate_mask_odd_shapes_and_chunks(core, oracle, monkeypatch):
    Y, Z, Xs = _cov_case(1001, 37, 3, seed=5)           # odd V -> one voxel per thread, unaligned rows
    mask = (np.arange(1001) % 5).astype(float)
    assert_rows_match(core.lm_fit_cov(Y, Z, Xs, mask=mask, mask_lo=2, mask_hi=3),
                      oracle.lm_loop_cov(Y, Z, Xs, mask=mask, lo=2, hi=3), RTOL, "mask")
    got = core.lm_fit_cov(Y, Z, Xs, mask=mask)
    assert (got[mask < 0.5] == 0).all() and not np.signbit(got[mask < 0.5]).any()
    monkeypatch.setenv("RMG_CHUNK_VOXELS", "300")      # several chunks through the host pipeline
    assert_rows_match(core.lm_fit_cov(Y, Z, Xs), oracle.lm_loop_cov(Y, Z, Xs), RTOL, "chunked")
    monkeypatch.delenv("RMG_CHUNK_VOXELS")
    Y, Z, Xs = _cov_case(64, 2000, 6, seed=6)           # few vertices, many subjects -> split-n
    assert_rows_match(core.lm_fit_cov(Y, Z, Xs), oracle.lm_loop_cov(Y, Z, Xs), RTOL, "split-n")
    Xn = design_cov(40, 3, seed=2)[:, 1:]               # no intercept in the static part: no shift
    Y, Z, _ = _cov_case(500, 40, 0, seed=7, zmean=1.0, zsd=1.0)
    assert_rows_match(core.lm_fit_cov(Y, Z, Xn), oracle.lm_loop_cov(Y, Z, Xn), RTOL, "no intercept")


def test_covariate_rank_and_error_contract(core, oracle):
    Y, Z, Xs = _cov_case(400, 30, 2, seed=8)
    Z[10, :] = 7.5                                      # constant: in span(Xs) -> rank p-1 at that voxel
    Z[20, :] = 3.0 + 2.0 * Xs[:, 1]                     # exactly a combination of the static columns
    Y[30, :] = 1.0 + 0.5 * Xs[:, 1] + 2.0 * Z[30, :]    # perfect fit through the covariate: rss is rounding noise
    got, want = core.lm_fit_cov(Y, Z, Xs), oracle.lm_loop_cov(Y, Z, Xs)
    keep = np.ones(400, bool)
    keep[[10, 20, 30]] = False
    assert_rows_match(got[keep], want[keep], RTOL, "regular voxels")
    for v in (10, 20):
        cols = [0, 1, 2, 3, 4, 8]                       # F, R2, betas, logLik are independent of R's noise diagonal
        assert_rows_match(got[v, cols], want[v, cols], RTOL, f"rank-deficient voxel {v}")
        assert got[v, 4] == 0.0 and got[v, 7] == 0.0 and want[v, 4] == 0.0 and want[v, 7] == 0.0
    assert got[30, 2:5] == pytest.approx([1.0, 0.5, 2.0], rel=1e-9) and got[30, 1] == pytest.approx(1.0, rel=1e-12)
    # all-zero covariate at a fitted voxel: the reference's dpotri error aborts the call ...
    Z0 = Z.copy(order="F")
    Z0[5, :] = 0.0
    with pytest.raises(core.SingularDesignError, match="element \\(3, 3\\) is zero, so the inverse cannot be computed"):
        core.lm_fit_cov(Y, Z0, Xs)
    with pytest.raises(oracle.OracleError):
        oracle.lm_loop_cov(Y, Z0, Xs)
    # ... but not when the mask excludes it
    mask = np.ones(400)
    mask[5] = 0
    assert_rows_match(core.lm_fit_cov(Y, Z0, Xs, mask=mask)[keep], oracle.lm_loop_cov(Y, Z0, Xs, mask=mask)[keep], RTOL, "masked zero")
    # rank-deficient static part is refused, not guessed
    with pytest.raises(core.RmincGpuError, match="rank-deficient static design|is zero, so the inverse"):
        core.lm_fit_cov(Y, Z, np.column_stack([Xs, Xs[:, 1]]))


def test_covariate_device_entry_point(core, oracle):
    import torch
    Y, Z, Xs = _cov_case(3000, 50, 3, seed=12)
    Yt = torch.from_numpy(np.ascontiguousarray(Y.T)).cuda()
    Zt = torch.from_numpy(np.ascontiguousarray(Z.T)).cuda()
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = core.lm_fit_cov_torch(Yt, Zt, Xs, status=status)
    torch.cuda.synchronize()
    assert int(status.item()) == 0
    assert_rows_match(out.cpu().numpy().T, oracle.lm_loop_cov(Y, Z, Xs), RTOL, "device")
    Zt[:, 17] = 0.0
    core.lm_fit_cov_torch(Yt, Zt, Xs, out=out, status=status)
    torch.cuda.synchronize()
    assert int(status.item()) == 4 and torch.isnan(out[:, 17]).all()


# ------------------------------------------------------------------ volumes in their stored type
def _stored_case(V, n, slice_len, seed, p=4, effect=300.0):
    rng = np.random.default_rng(seed)
    X = design_cfg2(n)[:, :p] if p <= 4 else design_cov(n, p, seed=seed + 1000)
    nsl = -(-V // slice_len)
    U = np.clip(rng.normal(30000, 400, (V, n)) + effect * np.outer(rng.random(V), X[:, min(1, p - 1)]), 0, 65535)
    U = np.asfortranarray(np.round(U).astype(np.uint16))
    smin = rng.normal(-0.8, 0.05, (nsl, n))
    smax = smin + rng.uniform(1.5, 2.5, (nsl, n))
    return U, X, (smax - smin) / 65535.0, smin


@pytest.mark.parametrize("V,slice_len", [(4096, 512), (4100, 410), (3001, 333), (2048, 2048)])
def test_stored_u16_matches_oracle_on_expanded_doubles(core, oracle, V, slice_len):
    """uint16 voxels + per-slice real ranges: the device forms the same doubles libminc would and fits them."""
    U, X, scale, offset = _stored_case(V, 60, slice_len, seed=V)
    Y = oracle.expand_stored(U, scale, offset, slice_len=slice_len)
    got = core.lm_fit_stored(U, X, scale, offset, slice_len=slice_len)
    assert_rows_match(got, oracle.vertex_lm_loop(Y, X), RTOL, "stored u16")
    # and it is the same fit as the double path on the expanded matrix, far inside the tolerance
    assert_rows_match(got, core.lm_fit(Y, X), 1e-11, "stored vs expanded on device")


def test_stored_variants(core, oracle, monkeypatch):
    U, X, scale, offset = _stored_case(5000, 48, 1000, seed=3)
    V = U.shape[0]
    # non-zero valid_min, mask with maskval, masked rows +0.0
    Y = oracle.expand_stored(U, scale, offset, voxel_min=100.0, slice_len=1000)
    mask = (np.arange(V) % 4).astype(float)
    got = core.lm_fit_stored(U, X, scale, offset, voxel_min=100.0, slice_len=1000, mask=mask, mask_lo=2, mask_hi=2)
    assert_rows_match(got, oracle.minc2_model_lm(Y, (5, 10, 100), X, mask=mask, lo=2, hi=2), RTOL, "mask")
    assert (got[mask != 2] == 0).all() and not np.signbit(got[mask != 2]).any()
    # several chunks through the host pipeline: slices keep their global index
    monkeypatch.setenv("RMG_CHUNK_VOXELS", "1400")
    assert_rows_match(core.lm_fit_stored(U, X, scale, offset, voxel_min=100.0, slice_len=1000),
                      oracle.vertex_lm_loop(Y, X), RTOL, "chunked")
    monkeypatch.delenv("RMG_CHUNK_VOXELS")
    # float32 volumes convert exactly
    F = np.asfortranarray(np.random.default_rng(4).normal(0, 0.2, (3000, 48)).astype(np.float32))
    assert_rows_match(core.lm_fit_stored(F, X), oracle.vertex_lm_loop(F.astype(np.float64), X), RTOL, "f32")
    F1 = np.asfortranarray(F[:2999])                                         # odd V: scalar path
    assert_rows_match(core.lm_fit_stored(F1, X), oracle.vertex_lm_loop(F1.astype(np.float64), X), RTOL, "f32 odd")
    # every padded design width
    for p in (1, 2, 3, 5, 7, 8, 11, 16):
        Up, Xp, sc, of = _stored_case(1024, 64, 256, seed=50 + p, p=p)
        Yp = oracle.expand_stored(Up, sc, of, slice_len=256)
        c0 = 2 if p == 1 else 0          # intercept only: F and R2 are 0/0 noise in the reference (see test_every_padded_width)
        assert_rows_match(core.lm_fit_stored(Up, Xp, sc, of, slice_len=256)[:, c0:], oracle.vertex_lm_loop(Yp, Xp)[:, c0:],
                          RTOL, f"p={p}")
    # argument contract
    with pytest.raises(core.RmincGpuError, match="voxel_min must be an integer"):
        core.lm_fit_stored(U, X, scale, offset, voxel_min=0.5, slice_len=1000)
    with pytest.raises(ValueError):
        core.lm_fit_stored(U.astype(np.int32), X, scale, offset)


@pytest.mark.parametrize("mode", ["fold", "exact", "plain"])
@pytest.mark.parametrize("slice_len", [1000, 333, 8000])
def test_stored_masked_fit_on_compacted_quads(core, oracle, monkeypatch, slice_len, mode):
    """uint16 files AND a mask: the mask is compacted into a list of voxel quads and only those are fitted.  Ragged masks (runs
    that start and end inside quads, empty stretches, a fully selected stretch), slice boundaries inside quads (333), both
    expansions, and the plain kernel (RMG_PACKED_NOCOMPACT=1) as the twin; masked rows are +0.0 in every column."""
    if mode == "exact":
        monkeypatch.setenv("RMG_STORED_EXACT", "1")
    if mode == "plain":
        monkeypatch.setenv("RMG_PACKED_NOCOMPACT", "1")
    V, n = 8000, 44
    U, X, scale, offset = _stored_case(V, n, slice_len, seed=77)
    U = U.copy(order="F")
    U[40:48] = 7                                                   # one stored value everywhere: the real values follow the ranges
    U[100] = (2000 + 30000 * X[:, 1]).astype(np.uint16)           # a large group effect: little is left for the residuals
    rng = np.random.default_rng(5)
    mask = np.zeros(V)
    mask[37:1203] = 1
    mask[1500:1502] = 1
    mask[2001] = 3
    mask[3000:5000] = (rng.random(2000) < 0.4) * 2.0
    mask[7996:] = 1
    Y = oracle.expand_stored(U, scale, offset, voxel_min=10.0, slice_len=slice_len)
    for lo, hi in ((1.0, 99999999.0), (2.0, 2.0)):
        got = core.lm_fit_stored(U, X, scale, offset, voxel_min=10.0, slice_len=slice_len, mask=mask, mask_lo=lo, mask_hi=hi)
        want = oracle.minc2_model_lm(Y, (1, 1, V), X, mask=mask, lo=lo, hi=hi)
        inside = (mask > lo - 0.5) & (mask < hi + 0.5)
        assert_rows_match(got[inside], want[inside], RTOL, f"compacted quads {mode} lo={lo}")
        assert (got[~inside] == 0).all() and not np.signbit(got[~inside]).any()
    F = np.asfortranarray(np.random.default_rng(4).normal(0, 0.2, (V, n)).astype(np.float32))
    gotf = core.lm_fit_stored(F, X, mask=mask)
    assert_rows_match(gotf, oracle.minc2_model_lm(F.astype(np.float64), (1, 1, V), X, mask=mask), RTOL, f"f32 quads {mode}")
    # an empty mask selects nothing
    assert (core.lm_fit_stored(U, X, scale, offset, voxel_min=10.0, slice_len=slice_len, mask=np.zeros(V)) == 0).all()


def test_stored_degenerate_and_near_perfect_voxels(core, oracle):
    """Quantised data: constant voxels (background) and voxels the design explains almost entirely."""
    rng = np.random.default_rng(8)
    V, n = 2048, 40
    X = design_cfg1(n)
    U = np.round(rng.normal(20000, 300, (V, n))).astype(np.uint16)
    U[::9] = 0                                                              # background: all subjects zero
    U[1::9] = (1000 + 20000 * X[:, 1] + rng.integers(0, 3, (len(U[1::9]), n))).astype(np.uint16)   # near-perfect fit
    U = np.asfortranarray(U)
    scale = np.full((1, n), 2.0 / 65535.0)
    offset = np.full((1, n), -1.0)
    Y = oracle.expand_stored(U, scale, offset)
    got, want = core.lm_fit_stored(U, X, scale, offset), oracle.vertex_lm_loop(Y, X)
    const = np.zeros(V, bool)
    const[::9] = True
    # constant voxels: rss is exactly 0 here (classes as SURVEY A.4: the reference's rounding noise is unmatchable)
    assert np.isposinf(got[const, -1]).all() and np.isnan(got[const, 0]).all()
    assert_rows_match(got[~const], want[~const], RTOL, "quantised")


@pytest.mark.parametrize("masked", [False, True])
def test_stored_fit_with_edges_inside_a_threads_voxels(core, oracle, masked):
    """An intensity image with tissue / background transitions: every 7th voxel is background (about 100 +- 30 counts next to
    30000 +- 400), every file has the same real range, so a voxel's spread is a fraction of a percent of the step to its
    neighbour.  The one-shift-per-thread arithmetic cannot serve such quads: the tile (TMA variant) or the warp (compacted
    quads, with a mask) votes itself into per-voxel shifts.  Either way the result equals the oracle's."""
    rng = np.random.default_rng(31)
    V, n = 80128, 40
    X = design_cfg2(n)
    U = np.clip(rng.normal(30000, 400, (V, n)) + 300.0 * np.outer(rng.random(V), X[:, 1]), 0, 65535)
    U[::7] = rng.normal(100, 30, (len(U[::7]), n)).clip(0, 65535)
    U[5::11] = 0                                                            # and constant voxels in between
    U = np.asfortranarray(np.round(U).astype(np.uint16))
    nsl, slice_len = 8, 10016
    scale = np.full((nsl, n), 2.0 / 65535.0)
    offset = np.full((nsl, n), -0.8)
    Y = oracle.expand_stored(U, scale, offset, slice_len=slice_len)
    mask = None
    if masked:
        mask = (rng.random(V) < 0.6).astype(float)
        mask[1000:3000] = 0
    got = core.lm_fit_stored(U, X, scale, offset, slice_len=slice_len, mask=mask)
    if not masked:
        assert core.last_fit_variant() == "lm_fit_packed_tma_kernel"
    want = oracle.minc2_model_lm(Y, (1, 1, V), X, mask=mask) if masked else oracle.vertex_lm_loop(Y, X)
    const = np.zeros(V, bool)
    const[5::11] = True
    inside = np.ones(V, bool) if mask is None else mask > 0.5
    ok = inside & ~const
    assert_rows_match(got[ok], want[ok], RTOL, f"edges, masked={masked}")
    assert np.isposinf(got[inside & const, -1]).all()                       # exact rss = 0 classes
    if masked:
        assert (got[~inside] == 0).all()


def test_stored_device_entry_point(core, oracle):
    import torch
    U, X, scale, offset = _stored_case(8192, 50, 1024, seed=21)
    Y = oracle.expand_stored(U, scale, offset, slice_len=1024)
    Ut = torch.from_numpy(np.ascontiguousarray(U.T).view(np.int16)).cuda().view(torch.uint16)
    sc = torch.from_numpy(np.ascontiguousarray(scale.T)).cuda()
    of = torch.from_numpy(np.ascontiguousarray(offset.T)).cuda()
    out = core.lm_fit_stored_torch(Ut, X, sc, of, slice_len=1024)
    torch.cuda.synchronize()
    assert_rows_match(out.cpu().numpy().T, oracle.vertex_lm_loop(Y, X), RTOL, "device u16")
    assert core.last_fit_variant() == "lm_fit_packed_kernel"


@pytest.mark.parametrize("slice_len", [1001, 4096, 80128])
@pytest.mark.parametrize("mode", ["fma16", "fma8", "exact"])
def test_stored_tma_ring_variant(core, oracle, monkeypatch, slice_len, mode):
    """Large V takes the TMA-staged kernel: stored samples + the tile's (scale, offset) rows through the shared-memory
    ring.  slice_len = 1001 puts slice boundaries inside tiles and inside a thread's four voxels.  Modes: one fused
    multiply-add per sample with 16 (default) or 8 subjects per ring stage (the fallback when the larger ring does not
    fit), and the reader's three roundings (RMG_STORED_EXACT=1)."""
    if mode == "fma8":
        monkeypatch.setenv("RMG_TMA_CFG", "1")
    if mode == "exact":
        monkeypatch.setenv("RMG_STORED_EXACT", "1")
    V, n = 80128, 37
    U, X, scale, offset = _stored_case(V, n, slice_len, seed=slice_len)
    Y = oracle.expand_stored(U, scale, offset, voxel_min=3.0, slice_len=slice_len)
    got = core.lm_fit_stored(U, X, scale, offset, voxel_min=3.0, slice_len=slice_len)
    assert core.last_fit_variant() == "lm_fit_packed_tma_kernel"
    assert_rows_match(got, oracle.vertex_lm_loop(Y, X), RTOL, "tma u16")
    # the one-FMA expansion differs from the reader's doubles by at most a rounding of the real range: the fits agree
    # far inside the tolerance
    assert_rows_match(got, core.lm_fit(Y, X), 1e-11, "tma u16 vs the double path on the reader's doubles")
    if mode != "fma16":
        return
    if slice_len == 4096:
        F = np.asfortranarray(Y.astype(np.float32))
        got = core.lm_fit_stored(F, X[:, :3])
        assert core.last_fit_variant() == "lm_fit_packed_tma_kernel"
        assert_rows_match(got, oracle.vertex_lm_loop(F.astype(np.float64), X[:, :3]), RTOL, "tma f32")
        Fr = np.asfortranarray(F[:79999])                                     # ragged tail: TMA zero-fill, scalar stores
        assert_rows_match(core.lm_fit_stored(Fr, X), oracle.vertex_lm_loop(Fr.astype(np.float64), X), RTOL, "tma f32 ragged")


def test_randomize_with_a_voxelwise_covariate(core, oracle):
    """mincRandomize on y ~ static + file covariate: the reference re-samples the covariate files with the static
    predictor rows and refits the per-voxel design (R/minc_voxel_statistics.R:1468-1477).  Against the oracle's literal
    loop over the restated covariate fit (itself bit-equal to the reference's vertex_lm_loop object code): true
    permutations, replace = TRUE index vectors, a mask, one-sided, the default intercept-only static part, two GPUs when
    there are two."""
    Y, Z, Xs = _cov_case(2501, 24, 3, seed=51)
    n, ps = Xs.shape
    p = ps + 1
    rng = np.random.default_rng(52)
    idx = np.column_stack([rng.permutation(n) for _ in range(7)] + [rng.integers(0, n, n) for _ in range(4)])
    cols = list(range(p + 2, 2 * p + 2))                                   # every tvalue- column, the covariate's last
    mask = (np.arange(2501) % 4 != 1).astype(float)
    for m in (None, mask):
        got = core.randomize_cov(Y, Z, Xs, idx, cols, mask=m)
        np.testing.assert_allclose(got, oracle.randomize_cov(Y, Z, Xs, idx, cols, mask=m), rtol=RTOL, atol=0)
    got = core.randomize_cov(Y, Z, Xs, idx, [0, 1, 2 * p + 1], two_sided=False)     # F, R-squared, the covariate's t, signed
    np.testing.assert_allclose(got, oracle.randomize_cov(Y, Z, Xs, idx, [0, 1, 2 * p + 1], two_sided=False), rtol=RTOL, atol=0)
    got = core.randomize_cov(Y, Z, None, idx[:, :5], [4, 5])                # y ~ z: static part = the intercept
    np.testing.assert_allclose(got, oracle.randomize_cov(Y, Z, None, idx[:, :5], [4, 5]), rtol=RTOL, atol=0)
    if oracle.have_ref():                                                   # the literal loop over the reference's object code
        np.testing.assert_allclose(core.randomize_cov(Y[:300], Z[:300], Xs, idx[:, :6], cols),
                                   oracle.randomize_cov(Y[:300], Z[:300], Xs, idx[:, :6], cols, use_ref=True), rtol=RTOL, atol=0)
    G = min(2, core.device_count())
    np.testing.assert_allclose(core.randomize_cov(Y, Z, Xs, idx, cols, mask=mask, n_gpus=G),
                               oracle.randomize_cov(Y, Z, Xs, idx, cols, mask=mask), rtol=RTOL, atol=0)
    # the identity permutation reproduces the column maxima of the plain covariate fit
    fit = core.lm_fit_cov(Y, Z, Xs)
    ident = np.arange(n)[:, None]
    np.testing.assert_allclose(core.randomize_cov(Y, Z, Xs, ident, cols)[0], np.abs(fit[:, cols]).max(axis=0), rtol=1e-12)
    # error contract: a re-sampled static part that loses rank; an all-zero covariate at a fitted voxel
    bad = np.zeros((n, 1), dtype=np.int64)
    with pytest.raises(core.RmincGpuError, match="rank-deficient static design|is zero, so the inverse"):
        core.randomize_cov(Y, Z, Xs, bad, cols)
    Z0 = Z.copy()
    Z0[17] = 0.0
    with pytest.raises(core.RmincGpuError, match="is zero, so the inverse cannot be computed"):
        core.randomize_cov(Y, Z0, Xs, idx[:, :2], cols)


# ------------------------------------------------------------------ mincFDR
@pytest.mark.parametrize("stat_type,df1,df2", [("t", 496, 0), ("t", 8, 0), ("F", 3, 496), ("F", 1, 18)])
def test_fdr_matches_oracle(core, oracle, stat_type, df1, df2):
    """p-values (incomplete beta on the device), Benjamini-Hochberg q-values and thresholds against p.adjust's
    algorithm with scipy's distributions; NaN statistics stay NaN and do not count; 1 outside the mask."""
    rng = np.random.default_rng(df1)
    V = 200_003
    if stat_type == "t":
        s = rng.standard_t(df1, V)
        s[: V // 20] += rng.normal(5, 1, V // 20)
    else:
        s = rng.f(df1, df2, V)
        s[: V // 20] *= 8
    s[::97] = np.nan
    s[5] = np.inf
    s[6] = 0.0
    mask = (rng.random(V) > 0.3).astype(float)
    code = core.STAT_T if stat_type == "t" else core.STAT_F
    for m in (None, mask):
        q, th = core.fdr(s, code, df1, df2, mask=m)
        wq, wth, wp = oracle.fdr(s, stat_type, df1, df2, mask=m)
        assert np.array_equal(np.isnan(q), np.isnan(wq))
        ok = ~np.isnan(wq)
        assert q[ok] == pytest.approx(wq[ok], rel=1e-9, abs=1e-300)
        if m is not None:
            assert (q[m < 0.5] == 1).all()
        assert np.array_equal(np.isnan(th), np.isnan(wth))
        assert th[~np.isnan(wth)] == pytest.approx(wth[~np.isnan(wth)], rel=1e-8)


@pytest.mark.parametrize("method,pi0", [("BY", 1.0), ("pFDR", 0.83), ("fastqvalue", 1.0), ("qvalue", 0.83)])
def test_fdr_other_methods_match_oracle(core, oracle, method, pi0):
    """mincFDR's methods beyond Benjamini-Hochberg (R/minc_FDR.R:429-438) against the restated p.adjust(., "BY"),
    fast.qvalue (whose qvalue_min step is pinned to the reference's object code in test_oracle.py) and qvalue forms."""
    rng = np.random.default_rng(11)
    V = 120_011
    s = rng.standard_t(40, V)
    s[: V // 15] += rng.normal(4.5, 1, V // 15)
    s[rng.integers(0, V, 500)] = np.round(s[rng.integers(0, V, 500)], 1)        # ties
    mask = (rng.random(V) > 0.25).astype(float)
    for m in (None, mask):
        q, th = core.fdr(s, core.STAT_T, 40, mask=m, method=method, pi0=pi0)
        wq, wth, _ = oracle.fdr(s, "t", 40, mask=m, method=method, pi0=pi0)
        assert q == pytest.approx(wq, rel=1e-9, abs=1e-300)
        if m is not None:
            assert (q[m < 0.5] == 1).all()
        assert np.array_equal(np.isnan(th), np.isnan(wth))
        assert th[~np.isnan(wth)] == pytest.approx(wth[~np.isnan(wth)], rel=1e-8)
    sf = rng.f(3, 60, 50_000)
    sf[:2000] *= 9
    q, th = core.fdr(sf, core.STAT_F, 3, 60, method=method, pi0=pi0)
    wq, wth, _ = oracle.fdr(sf, "F", 3, 60, method=method, pi0=pi0)
    assert q == pytest.approx(wq, rel=1e-9, abs=1e-300) and th[~np.isnan(wth)] == pytest.approx(wth[~np.isnan(wth)], rel=1e-8)


def test_fdr_storey_edge_cases_and_pi0_grid(core, oracle):
    """The largest p as the LAST element (the reference reads past the end there: kept raw), as a middle element with a
    successor above 1 (becomes 1); the pi0 grid; the reference's error cases."""
    from scipy import stats as S
    t_of_p = lambda p: S.t.isf(np.asarray(p) / 2.0, 25)
    for p in ([0.2, 0.9, 0.01, 0.95], [0.2, 0.95, 0.9, 0.01], [0.97, 0.99, 0.98, 0.96, 0.5], [0.3]):
        s = t_of_p(p)
        for pi0 in (1.0, 0.6):
            q, _ = core.fdr(s, core.STAT_T, 25, method="pFDR", pi0=pi0)
            assert q == pytest.approx(oracle.fast_qvalue(np.asarray(p, float), pi0), rel=1e-8), (p, pi0)
    rng = np.random.default_rng(2)
    s = rng.standard_t(25, 30_001)
    mask = (rng.random(30_001) > 0.5).astype(float)
    lam = np.arange(0, 19) * 0.05
    _, _, pv = oracle.fdr(s, "t", 25, mask=mask)
    assert core.fdr_pi0_fractions(s, core.STAT_T, 25, mask=mask, lambdas=lam) == pytest.approx(oracle.pi0_fractions(pv, lam), abs=2.0 / pv.size)
    with pytest.raises(core.RmincGpuError, match="pi0 <= 0"):
        core.fdr(s, core.STAT_T, 25, method="pFDR", pi0=0.0)
    s[3] = np.nan
    with pytest.raises(core.RmincGpuError, match="missing value"):
        core.fdr(s, core.STAT_T, 25, method="qvalue", pi0=0.9)
    with pytest.raises(core.RmincGpuError, match="missing value"):
        core.fdr_pi0_fractions(s, core.STAT_T, 25)
    with pytest.raises(core.RmincGpuError, match=r"\[0, 1\)"):
        core.fdr_pi0_fractions(s[4:], core.STAT_T, 25, lambdas=[0.5, 1.0])
    q, th = core.fdr(s, core.STAT_T, 25, method="BY")                            # p.adjust drops the NA, as for "fdr"
    assert np.isnan(q[3]) and q[np.arange(30_001) != 3] == pytest.approx(oracle.fdr(s, "t", 25, method="BY")[0][np.arange(30_001) != 3], rel=1e-9)


def test_fdr_edge_cases_and_device_form(core, oracle):
    import torch
    # nothing significant: every threshold NA; everything masked: all ones
    s = np.random.default_rng(1).standard_t(30, 5000)
    q, th = core.fdr(s, core.STAT_T, 30)
    wq, wth, _ = oracle.fdr(s, "t", 30)
    assert q == pytest.approx(wq, rel=1e-9) and np.isnan(th).all() and np.isnan(wth).all()
    q, th = core.fdr(s, core.STAT_T, 30, mask=np.zeros(5000))
    assert (q == 1).all() and np.isnan(th).all()
    q, th = core.fdr(np.zeros(0), core.STAT_T, 30)
    assert q.shape == (0,)
    with pytest.raises(core.RmincGpuError, match="degrees of freedom"):
        core.fdr(s, core.STAT_F, 3, 0)
    st = torch.from_numpy(np.r_[s, 9.0, -11.0]).cuda()
    qd, thd = core.fdr_torch(st, core.STAT_T, 30)
    wq, wth, _ = oracle.fdr(np.r_[s, 9.0, -11.0], "t", 30)
    assert qd.cpu().numpy() == pytest.approx(wq, rel=1e-9)
    assert thd[~np.isnan(wth)] == pytest.approx(wth[~np.isnan(wth)], rel=1e-8)


def test_sharded_forms_of_the_next_row_paths(core, oracle):
    """rmg_lm_fit_cov_multi / rmg_lm_fit_stored_multi: contiguous voxel shards over the visible GPUs (1 or more),
    slices keep their global index across shards."""
    G = min(2, core.device_count())
    Y, Z, Xs = _cov_case(3001, 40, 3, seed=31)
    mask = (np.arange(3001) % 3 != 0).astype(float)
    assert_rows_match(core.lm_fit_cov(Y, Z, Xs, mask=mask, n_gpus=G), oracle.lm_loop_cov(Y, Z, Xs, mask=mask), RTOL, "cov multi")
    U, X, scale, offset = _stored_case(6002, 36, 700, seed=32)
    Yx = oracle.expand_stored(U, scale, offset, voxel_min=1.0, slice_len=700)
    assert_rows_match(core.lm_fit_stored(U, X, scale, offset, voxel_min=1.0, slice_len=700, n_gpus=G),
                      oracle.vertex_lm_loop(Yx, X), RTOL, "stored multi")
    with pytest.raises(core.RmincGpuError, match="GPUs requested"):
        core.lm_fit_stored(U, X, scale, offset, slice_len=700, n_gpus=core.device_count() + 1)


def test_pageable_input_through_the_pinned_staging_ring(core, oracle, monkeypatch):
    """Pageable host input (R's heap, plain numpy) is copied by the host threads into a ring of pinned slots that the
    DMA engine drains.  RMG_STAGING_ROWS forces that path at test sizes: doubles, a second operand (covariate) and
    2-byte samples, several chunks, more subject groups than ring slots."""
    monkeypatch.setenv("RMG_STAGING_ROWS", "3")
    monkeypatch.setenv("RMG_CHUNK_VOXELS", "2000")
    Y, X, mask = cfg1_data(dims=(11, 23, 29))
    assert_rows_match(core.lm_fit(Y, X, mask=mask), oracle.minc2_model_lm(Y, (11, 23, 29), X, mask=mask), RTOL, "staged f64")
    Yc, Zc, Xs = _cov_case(5001, 31, 3, seed=41)
    assert_rows_match(core.lm_fit_cov(Yc, Zc, Xs), oracle.lm_loop_cov(Yc, Zc, Xs), RTOL, "staged covariate")
    U, Xu, scale, offset = _stored_case(7003, 29, 500, seed=42)
    Yu = oracle.expand_stored(U, scale, offset, slice_len=500)
    assert_rows_match(core.lm_fit_stored(U, Xu, scale, offset, slice_len=500), oracle.vertex_lm_loop(Yu, Xu), RTOL, "staged u16")
    F = np.asfortranarray(Yu.astype(np.float32))
    assert_rows_match(core.lm_fit_stored(F, Xu), oracle.vertex_lm_loop(F.astype(np.float64), Xu), RTOL, "staged f32")


def test_stored_device_shards_keep_global_slices(core, oracle):
    """Two voxel shards of one volume through the device entry point with v_offset: the rows equal the unsharded fit."""
    import torch
    V, n, slice_len = 90_000, 20, 7001
    U, X, scale, offset = _stored_case(V, n, slice_len, seed=77)
    Ut = torch.from_numpy(np.ascontiguousarray(U.T).view(np.int16)).cuda().view(torch.uint16)
    sc = torch.from_numpy(np.ascontiguousarray(scale.T)).cuda()
    of = torch.from_numpy(np.ascontiguousarray(offset.T)).cuda()
    whole = core.lm_fit_stored_torch(Ut, X, sc, of, slice_len=slice_len)
    a = 44_000                                                       # 8-voxel aligned cut inside slice 6
    lo = core.lm_fit_stored_torch(Ut[:, :a], X, sc, of, slice_len=slice_len)
    hi = core.lm_fit_stored_torch(Ut[:, a:], X, sc, of, slice_len=slice_len, v_offset=a)
    torch.cuda.synchronize()
    # (the shards are small enough for the direct-load kernel, the whole volume takes the TMA ring)
    assert_rows_match(torch.cat([lo, hi], dim=1).cpu().numpy().T, whole.cpu().numpy().T, 1e-12, "shards vs whole")
    Y = oracle.expand_stored(U, scale, offset, slice_len=slice_len)
    assert_rows_match(whole.cpu().numpy().T, oracle.vertex_lm_loop(Y, X), RTOL, "whole")


# ------------------------------------------------------------------ vertexTFCE
def zeros_agree(got, want, scale):
    """The lazy sweep reconstructs a vertex's value as a difference of cluster-sized sums: vertices that were never active
    are exactly 0, but a contribution below an ulp of its cluster's accumulated value (the reference's last threshold is a
    rounding residue ~1e-16) may come out as 0 or as rounding noise.  Zero patterns therefore agree up to 1e-12 of the
    map maximum; the level sweep (RMG_TFCE_SWEEP=1) reproduces them exactly."""
    tol = 1e-12 * scale
    return bool((np.abs(got[want == 0]) <= tol).all() and (np.abs(want[got == 0]) <= tol).all())


def test_graph_tfce_matches_oracle(core, oracle):
    """Batched graph TFCE against the restated (and reference-pinned) weighted-quick-union sweep: sides, weights,
    exponents, step counts; maps with ties, zeros and an all-negative map."""
    from helpers import grid_adjacency
    rng = np.random.default_rng(5)
    a, b = 60, 70
    adj = grid_adjacency(a, b)
    ptr, idx = oracle.adjacency_csr(adj)
    V = a * b
    M = rng.standard_normal((V, 7))
    M[1000:1200, 0] += 3.0
    M[:, 1] = np.round(M[:, 1], 1)                # many exact ties
    M[::5, 2] = 0.0
    M[:, 3] = -np.abs(M[:, 3]) - 0.5              # nothing positive
    M[:, 4] = np.abs(M[:, 4])                     # nothing negative
    w = rng.uniform(0.5, 2.0, V)
    for side in ("both", "positive", "negative"):
        for weights in (None, w):
            for E, H, ns in ((0.5, 2.0, 100), (1.0, 1.0, 37), (0.66, 2.0, 253)):
                got = core.graph_tfce(M, adj, E=E, H=H, nsteps=ns, side=side, weights=weights)
                want = np.column_stack([oracle.graph_tfce(M[:, j], ptr, idx, E=E, H=H, nsteps=ns, side=side, weights=weights)
                                        for j in range(M.shape[1])])
                scale = np.abs(want).max(axis=0, keepdims=True) + 1e-300
                assert np.abs(got - want).max() <= 1e-10 * scale.max(), (side, weights is None, E, H, ns)
                assert zeros_agree(got, want, scale.max())
    # a single vector, CSR input, argument contract
    v = core.graph_tfce(M[:, 0], (ptr, idx))
    w0 = oracle.graph_tfce(M[:, 0], ptr, idx)
    assert v.shape == (V,) and np.abs(v - w0).max() <= 1e-10 * np.abs(w0).max()
    with pytest.raises(ValueError, match="Mismatch between adjacency list and data"):
        core.graph_tfce(M[:-1, 0], adj)
    with pytest.raises(core.RmincGpuError, match="nsteps"):
        core.graph_tfce(M[:, 0], adj, nsteps=1000)
    # a one-directional edge (a directed igraph, a hand-built list) would give other clusters than the reference, which
    # follows an edge from its higher-valued end only: refused
    one_way = [list(x) for x in adj]
    one_way[17].remove(18)
    with pytest.raises(core.RmincGpuError, match="not symmetric: vertex 18"):
        core.graph_tfce(M[:, 0], one_way)


def test_tfce_lazy_sweep_equals_level_sweep(core, oracle, monkeypatch):
    """The two device formulations of the enhancement: the lazy sweep (default; work = newly active vertices + merges per
    level, values resolved by pointer jumping) and the level sweep over every active vertex (RMG_TFCE_SWEEP=1), which
    reproduces the reference's zero pattern exactly.  Graphs and voxel grids, ties, weights, maps with one giant cluster and
    with thousands of tiny ones, long chains (deep merge histories)."""
    from helpers import grid_adjacency
    rng = np.random.default_rng(9)
    a, b = 90, 80
    adj = grid_adjacency(a, b)
    ptr, idx = oracle.adjacency_csr(adj)
    V = a * b
    M = rng.standard_normal((V, 6))
    M[:, 1] = np.round(M[:, 1], 1)
    M[:, 2] = 3.0 + 0.01 * rng.standard_normal(V)              # everything joins one cluster early
    M[:, 3] = np.linspace(4.0, 0.1, V)                         # a monotone ramp: one long chain of merges in index order
    M[:, 4] = np.linspace(0.1, 4.0, V)                         # ... and against it
    w = rng.uniform(0.5, 2.0, V)
    for weights in (None, w):
        for side in ("both", "positive"):
            lazy = core.graph_tfce(M, adj, side=side, weights=weights)
            monkeypatch.setenv("RMG_TFCE_SWEEP", "1")
            sweep = core.graph_tfce(M, adj, side=side, weights=weights)
            monkeypatch.delenv("RMG_TFCE_SWEEP")
            want = np.column_stack([oracle.graph_tfce(M[:, j], ptr, idx, side=side, weights=weights) for j in range(M.shape[1])])
            scale = np.abs(want).max(axis=0, keepdims=True)
            assert (np.abs(sweep - want) <= 1e-10 * scale).all() and np.array_equal(sweep == 0, want == 0)
            assert (np.abs(lazy - want) <= 1e-10 * scale).all() and zeros_agree(lazy, want, scale.max())
            assert (np.abs(lazy - sweep) <= 1e-12 * scale).all()
    dims = (14, 17, 12)
    Mv = rng.standard_normal((int(np.prod(dims)), 3))
    for conn in (6, 26):
        lazy = core.volume_tfce(Mv, dims, d=0.1, connectivity=conn, voxel_volume=0.2)
        monkeypatch.setenv("RMG_TFCE_SWEEP", "1")
        sweep = core.volume_tfce(Mv, dims, d=0.1, connectivity=conn, voxel_volume=0.2)
        monkeypatch.delenv("RMG_TFCE_SWEEP")
        assert np.abs(lazy - sweep).max() <= 1e-12 * np.abs(sweep).max() and np.array_equal(lazy == 0, sweep == 0)


def test_randomize_tfce_matches_literal_loop(core, oracle, monkeypatch):
    """vertexTFCE.vertexLm's permutation test: per permutation refit -> non-finite to 0 -> enhance -> abs -> maximum."""
    from helpers import grid_adjacency, tfce_randomize_oracle
    rng = np.random.default_rng(9)
    a, b, n = 24, 31, 16
    adj = grid_adjacency(a, b)
    V = a * b
    X = np.column_stack([np.ones(n), np.arange(n) % 2, rng.normal(size=n)])
    Y = np.asfortranarray(rng.normal(3, 0.4, (V, n)))
    Y[200:260] += 0.8 * X[:, 1]
    Y[7] = 0.0                                        # a degenerate vertex: NaN statistics -> 0
    idx = np.column_stack([rng.permutation(n) for _ in range(11)]).astype(np.int32)
    cols = [5, 6, 7]
    for two_sided, side in ((True, "both"), (False, "positive")):
        got = core.randomize_tfce(Y, X, idx, cols, adj, two_sided=two_sided, side=side, nsteps=50)
        want = tfce_randomize_oracle(oracle, Y, X, idx, cols, adj, two_sided=two_sided, side=side, nsteps=50)
        np.testing.assert_allclose(got, want, rtol=1e-8)
    # permutations sharded over the visible GPUs: same rows
    G = min(2, core.device_count())
    np.testing.assert_allclose(core.randomize_tfce(Y, X, idx, cols, adj, two_sided=False, side="positive", nsteps=50, n_gpus=G),
                               want, rtol=1e-8)
    # pageable Y through the pinned staging ring (forced at this size)
    monkeypatch.setenv("RMG_STAGING_ROWS", "3")
    np.testing.assert_allclose(core.randomize_tfce(Y, X, idx, cols, adj, two_sided=False, side="positive", nsteps=50), want, rtol=1e-8)


# ------------------------------------------------------------------ mincTFCE on volumes
def test_volume_tfce_matches_restatement_and_the_graph_path(core, oracle):
    """rmg_volume_tfce against the scipy-labelling restatement (every connectivity, side, exponents, several maps at
    once), and against the graph path on the reference's neighbour_list with the thresholds made to coincide."""
    rng = np.random.default_rng(21)
    dims = (12, 17, 9)
    V = int(np.prod(dims))
    M = rng.standard_normal((V, 5))
    blob = np.zeros(dims)
    blob[3:8, 4:11, 2:6] = 2.5
    M[:, 0] += blob.reshape(-1)
    M[:, 1] = np.round(M[:, 1], 1)                 # exact ties, values sitting on thresholds
    M[::3, 2] = 0.0                                # masked-out voxels are zeros
    M[:, 3] = -np.abs(M[:, 3]) - 0.2               # nothing positive
    M[:, 4] *= 0.01                                # below the first threshold everywhere -> all zero
    for conn in (6, 18, 26):
        for side in ("both", "positive", "negative"):
            for d, E, H, vv in ((0.1, 0.5, 2.0, 1.0), (0.25, 1.0, 1.0, 0.001), (0.05, 0.66, 2.0, 0.056 ** 3)):
                got = core.volume_tfce(M, dims, d=d, E=E, H=H, side=side, connectivity=conn, voxel_volume=vv)
                want = np.column_stack([oracle.volume_tfce(M[:, j], dims, d=d, E=E, H=H, side=side, connectivity=conn,
                                                           voxel_volume=vv) for j in range(M.shape[1])])
                scale = np.abs(want).max() + 1e-300
                assert np.abs(got - want).max() <= 1e-10 * scale, (conn, side, d)
                assert zeros_agree(got, want, scale)
                assert (got[:, 4] == 0).all()
    # test_mincTFCE.R:153-170: vertexTFCE on neighbour_list(.., 6) with weights = voxel volume is the same enhancement
    z, y, x = dims
    v = M[:, 0].copy()
    v[v > 3.0] = 3.0
    v[100] = 3.0                                    # hmax = 30 d
    nl = core.neighbour_list(x, y, z, 6)
    a = core.volume_tfce(v, dims, d=0.1, voxel_volume=0.001)
    b = core.graph_tfce(v, nl, nsteps=30, weights=0.001)
    # (graph_tfce's negative side has its own hmax, so compare where the positive side decides)
    pos = core.volume_tfce(v, dims, d=0.1, voxel_volume=0.001, side="positive")
    np.testing.assert_allclose(pos, core.graph_tfce(v, nl, nsteps=30, weights=0.001, side="positive"), rtol=1e-9, atol=1e-10 * np.abs(pos).max())
    assert a.shape == b.shape
    # argument contract
    with pytest.raises(core.RmincGpuError, match="increase d"):
        core.volume_tfce(v * 100, dims, d=0.1)
    with pytest.raises(core.RmincGpuError, match="connectivity must be 6, 18 or 26"):
        core.volume_tfce(v, dims, connectivity=8)
    with pytest.raises(ValueError):
        core.volume_tfce(v, (12, 17, 8))


def test_randomize_volume_tfce_matches_literal_loop(core, oracle):
    """mincTFCE.mincLm's permutation test with a mask: refit -> non-finite to 0 -> enhance -> abs -> maximum."""
    rng = np.random.default_rng(23)
    dims, n = (9, 10, 8), 14
    V = int(np.prod(dims))
    X = np.column_stack([np.ones(n), np.arange(n) % 2, rng.normal(size=n)])
    Y = np.asfortranarray(rng.normal(0, 0.3, (V, n)))
    eff = np.zeros(dims)
    eff[2:6, 3:7, 2:5] = 0.5
    Y += np.outer(eff.reshape(-1), X[:, 1])
    Y[11] = 0.0
    mask = ellipsoid_mask(dims)
    idx = np.column_stack([rng.permutation(n) for _ in range(9)]).astype(np.int32)
    cols = [5, 6, 7]
    for two_sided, side, m in ((True, "both", mask), (False, "positive", None)):
        got = core.randomize_volume_tfce(Y, X, idx, cols, dims, two_sided=two_sided, side=side, d=0.2, mask=m, voxel_volume=0.5)
        want = oracle.randomize_volume_tfce(Y, X, idx, cols, dims, two_sided=two_sided, side=side, d=0.2, mask=m, voxel_volume=0.5)
        np.testing.assert_allclose(got, want, rtol=1e-8)
    G = min(2, core.device_count())
    np.testing.assert_allclose(core.randomize_volume_tfce(Y, X, idx, cols, dims, two_sided=False, side="positive", d=0.2,
                                                          voxel_volume=0.5, n_gpus=G), want, rtol=1e-8)


def test_volume_tfce_golden_vectors_of_the_reference_object_code(core):
    """Golden case F: the reference's compiled neighbour_list + graph_tfce on a voxel grid (weights = voxel volume, 8 steps
    of 0.25 up to hmax = 2) against rmg_volume_tfce with d = 0.25, and against rmg_graph_tfce on the golden adjacency."""
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_golden.npz"))
    dims = tuple(int(v) for v in G["F_dims"])
    vol = G["F_vol"]
    for conn in (6, 18, 26):
        for side in ("positive", "negative", "both"):
            want = G[f"F_tfce{conn}_{side}"]
            tol = 1e-10 * np.abs(want).max()
            got = core.volume_tfce(vol, dims, d=0.25, side=side, connectivity=conn, voxel_volume=0.001)
            assert np.abs(got - want).max() <= tol and zeros_agree(got, want, np.abs(want).max()), (conn, side)
            g2 = core.graph_tfce(vol, (G[f"F_ptr{conn}"], G[f"F_idx{conn}"]), nsteps=8, side=side, weights=0.001)
            assert np.abs(g2 - want).max() <= tol and zeros_agree(g2, want, np.abs(want).max()), (conn, side)

