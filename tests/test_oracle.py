"""The CPU oracle is pinned before anything is compared with it.

(1) the reference's own object code (oracle/_ref, built from /root/reference/src when present)
    and the committed golden vectors it produced (tests/golden/ref_golden.npz);
(2) published base-R lm() outputs (tests/golden/r_lm_known_answers.json) -- what the
    reference's testthat files compare with (tests/testthat/test_mincLm.R:34-56);
(3) an independent numpy closed-form OLS and scipy's LAPACK dpotri;
(4) SURVEY.md Appendix D known-answer vector, A.3 rank-deficiency and A.4 degenerate voxels.
"""
import json
import os

import numpy as np
import pytest

from helpers import assert_rows_match, cfg1_data, design_cfg2

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def numpy_ols(y, X):
    """Closed-form summary.lm identities (full-rank designs)."""
    n, p = X.shape
    beta, *_ = np.linalg.lstsq(X, y, rcond=None)
    fitted = X @ beta
    resid = y - fitted
    rss = resid @ resid
    mss = ((fitted - fitted.mean()) ** 2).sum()
    resvar = rss / (n - p)
    se = np.sqrt(np.diag(np.linalg.inv(X.T @ X)) * resvar)
    return dict(coef=beta, t=beta / se, F=(mss / (p - 1)) / resvar, R2=mss / (mss + rss),
                logLik=-0.5 * n * (np.log(2 * np.pi) + 1 - np.log(n) + np.log(rss)))


APPX_D_X = np.array([[1, 0, 2], [1, 0, 3], [1, 0, 5], [1, 1, 1], [1, 1, 4], [1, 1, 6], [1, 0, 7], [1, 1, 8]], float)
APPX_D_Y = np.array([1.0, 2.5, 2.0, 4.0, 3.5, 6.0, 3.0, 7.5])


def test_appendix_d_known_answer(oracle):
    r = oracle.voxel_lm(APPX_D_Y, APPX_D_X)
    assert r["F"] == pytest.approx(17.0949875380781, rel=1e-13)
    assert r["R2"] == pytest.approx(0.87241635162350284, rel=1e-14)
    assert r["coef"] == pytest.approx([0.217620481927707, 2.90060240963855, 0.448795180722892], rel=1e-12)
    assert r["t"] == pytest.approx([0.291073571681093, 4.51468309113494, 3.20108273476179], rel=1e-12)
    assert r["logLik"] == pytest.approx(-8.6568452963150087, rel=1e-14)
    assert list(r["pivot"]) == [0, 1, 2] and r["rank"] == 3
    # permutation check (mincRandomize semantics: X_pi = X[idx, ], same y)
    idx = [3, 0, 7, 1, 5, 2, 6, 4]
    rp = oracle.voxel_lm(APPX_D_Y, APPX_D_X[idx])
    assert rp["t"] == pytest.approx([1.68979927547816, -0.234970271257997, 0.224649658769916], rel=1e-12)


def test_published_r_lm_outputs(oracle):
    cases = json.load(open(os.path.join(GOLD, "r_lm_known_answers.json")))["cases"]
    assert len(cases) >= 5
    for c in cases:
        y = np.array(c["y"], float)
        X = np.column_stack([np.ones(len(y))] + [np.array(v, float) for v in c["X"].values()])
        r = oracle.voxel_lm(y, X)
        assert np.allclose(r["coef"], c["beta"], rtol=0, atol=c["beta_tol"]), c["name"]
        assert np.allclose(r["t"], c["t"], rtol=0, atol=c["t_tol"]), c["name"]
        assert abs(r["F"] - c["F"]) <= c["F_tol"], c["name"]
        assert abs(r["R2"] - c["R2"]) <= c["R2_tol"], c["name"]
        if "logLik" in c:
            assert abs(r["logLik"] - c["logLik"]) <= c["logLik_tol"], c["name"]


def test_golden_vectors_from_reference_object_code(oracle):
    """Committed outputs of the reference's own vertex_lm_loop / minc2_model (make_golden.py)."""
    g = np.load(os.path.join(GOLD, "ref_golden.npz"))
    assert np.array_equal(oracle.vertex_lm_loop(g["A_Y"], g["A_X"]), g["A_out"], equal_nan=True)
    dims = tuple(int(d) for d in g["B_dims"])
    assert np.array_equal(oracle.minc2_model_lm(g["B_Y"], dims, g["B_X"], mask=g["B_mask"]), g["B_out_default"])
    assert np.array_equal(oracle.minc2_model_lm(g["B_Y"], dims, g["B_X"], mask=g["B_mask"], lo=2, hi=2),
                          g["B_out_maskval2"])
    assert np.array_equal(oracle.minc2_model_lm(g["B_Y"], dims, g["B_X"]), g["B_out_nomask"])
    assert np.array_equal(oracle.vertex_lm_loop(g["C_Y"], g["C_X"]), g["C_out"])
    # the fixture itself exercises what it claims to
    assert np.isposinf(g["A_out"][5, -1]) and np.isnan(g["A_out"][5, 0])       # all-zero vertex
    assert (g["B_out_maskval2"][g["B_mask"] != 2] == 0).all()
    assert (g["C_out"][:, 5] == 0).all()                                        # aliased slot is beta 0


def test_restatement_equals_reference_object_code(oracle):
    """Bit-for-bit against oracle/_ref (the reference's compiled voxel_lm etc.), when present."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    rng = np.random.default_rng(7)
    for n, p, V in [(20, 2, 300), (37, 4, 200), (12, 6, 100), (9, 1, 50)]:
        X = np.column_stack([np.ones(n)] + [rng.normal(size=n) for _ in range(p - 1)])
        Y = rng.normal(5, 2, (V, n))
        a = oracle.vertex_lm_loop(Y, X)
        b = oracle.vertex_lm_loop(Y, X, use_ref=True)
        assert np.array_equal(a, b, equal_nan=True), (n, p)
    # slice loop + mask through the reference's minc2_model with in-memory volumes
    dims = (5, 7, 6)
    V = int(np.prod(dims))
    X = design_cfg2(30)
    Y = rng.normal(0, 0.2, (V, 30))
    mask = rng.integers(0, 3, V).astype(float)
    for kw in (dict(), dict(lo=2, hi=2), dict(lo=1, hi=1)):
        a = oracle.minc2_model_lm(Y, dims, X, mask=mask, **kw)
        b = oracle.minc2_model_lm(Y, dims, X, mask=mask, use_ref=True, fill=0.0, **kw)
        assert np.array_equal(a, b), kw
    # Quirk Q1: the reference writes only column 0 of masked rows; the rest is never touched
    a = oracle.minc2_model_lm(Y, dims, X, mask=mask)
    c = oracle.minc2_model_lm(Y, dims, X, mask=mask, use_ref=True, fill=np.nan)
    inside = mask > 0.5
    assert (c[~inside, 0] == 0).all() and np.isnan(c[~inside, 1:]).all()
    assert np.array_equal(c[inside], a[inside])
    # a single voxel through voxel_lm, and the singular-design error contract
    r1 = oracle.voxel_lm(Y[0], X)
    r2 = oracle.voxel_lm(Y[0], X, use_ref=True)
    for k in ("coef", "t"):
        assert np.array_equal(r1[k], r2[k])
    # an all-zero column (e.g. a factor level emptied by resampling) is cycled to the back and
    # leaves R[4,4] == 0 exactly -> dpotri info 4 -> Rf_error (src/modelling_functions.c:551-557)
    Xz = np.column_stack([X[:, 0], X[:, 1], np.zeros(30), X[:, 2]])
    with pytest.raises(oracle.OracleError, match=r"element \(4, 4\) is zero"):
        oracle.voxel_lm(Y[0], Xz, use_ref=True)
    with pytest.raises(oracle.OracleError, match=r"element \(4, 4\) is zero"):
        oracle.voxel_lm(Y[0], Xz)


def test_against_independent_numpy_ols(oracle):
    rng = np.random.default_rng(11)
    for n, p in [(20, 2), (50, 4), (200, 7), (500, 4)]:
        X = np.column_stack([np.ones(n)] + [rng.normal(size=n) for _ in range(p - 1)])
        for mean, sd in [(0, 1), (100, 10), (30000, 50)]:
            y = rng.normal(mean, sd, n) + X[:, -1] * sd
            r, w = oracle.voxel_lm(y, X), numpy_ols(y, X)
            assert r["coef"] == pytest.approx(w["coef"], rel=1e-9, abs=1e-9 * sd)
            assert r["t"] == pytest.approx(w["t"], rel=1e-9, abs=1e-10)
            assert r["F"] == pytest.approx(w["F"], rel=1e-9)
            assert r["R2"] == pytest.approx(w["R2"], rel=1e-9)
            assert r["logLik"] == pytest.approx(w["logLik"], rel=1e-11)


def test_dpotri_restatement_matches_lapack(oracle):
    from scipy.linalg import lapack
    rng = np.random.default_rng(5)
    for n, p in [(30, 3), (100, 6), (500, 4)]:
        X = np.column_stack([np.ones(n)] + [rng.normal(60, 10, size=n) for _ in range(p - 1)])
        d = oracle.design_probe(X)
        assert d["info"] == 0 and d["rank"] == p
        inv, info = lapack.dpotri(d["R"], lower=0)
        assert info == 0
        assert d["c"] == pytest.approx(np.diag(inv), rel=1e-12)
        assert d["c"] == pytest.approx(np.diag(np.linalg.inv(X.T @ X)), rel=1e-8)


def test_rank_deficient_behaviour(oracle):
    """pivoted-order betas with a trailing zero, rdf = n - p, near-zero t."""
    rng = np.random.default_rng(3)
    n = 40
    g = (np.arange(n) % 2).astype(float)
    age = np.round(rng.normal(60, 10, n), 1)
    X = np.column_stack([np.ones(n), g, 1 - g, age])
    d = oracle.design_probe(X)
    assert list(d["pivot"]) == [0, 1, 3, 2] and d["rank"] == 3 and d["info"] == 0
    y = rng.normal(size=n)
    r = oracle.voxel_lm(y, X)
    w = numpy_ols(y, X[:, [0, 1, 3]])
    assert r["coef"][:3] == pytest.approx(w["coef"], rel=1e-9)     # [b_1, b_g, b_age] ...
    assert r["coef"][3] == 0.0                                      # ... and the aliased slot is 0
    assert np.all(np.abs(r["t"][[0, 1, 3]]) < 1e-6)                # huge c from the ~1e-16 diagonal
    # an exact duplicate column behaves the same way (its residual is rounding noise, not 0) ...
    Xd = np.column_stack([np.ones(n), g, g, age])
    dd = oracle.design_probe(Xd)
    assert list(dd["pivot"]) == [0, 1, 3, 2] and dd["rank"] == 3 and dd["info"] == 0
    # ... while an all-zero column leaves R[4,4] == 0 exactly -> dpotri info 4 -> the reference raises
    Xz = np.column_stack([np.ones(n), g, np.zeros(n), age])
    assert oracle.design_probe(Xz)["info"] == 4
    with pytest.raises(oracle.OracleError):
        oracle.voxel_lm(y, Xz)


def test_degenerate_voxels(oracle):
    """y == 0 gives logLik=+Inf, F/R2/t = NaN, beta = 0."""
    X = design_cfg2(30)
    r = oracle.voxel_lm(np.zeros(30), X)
    assert np.isposinf(r["logLik"]) and np.isnan(r["F"]) and np.isnan(r["R2"])
    assert np.isnan(r["t"]).all() and (r["coef"] == 0).all()


def test_layout_and_mask_rule(oracle):
    """Columns [F, R2, beta.., t.., logLik]; mask interval rule lo-0.5 < m < hi+0.5; masked rows all 0."""
    Y, X, mask = cfg1_data(dims=(6, 8, 5))
    V = Y.shape[0]
    out = oracle.minc2_model_lm(Y, (6, 8, 5), X, mask=mask)
    inside = mask > 0.5
    assert (out[~inside] == 0).all() and not np.signbit(out[~inside]).any()
    v = int(np.flatnonzero(inside)[0])
    r = oracle.voxel_lm(Y[v], X)
    assert out[v, 0] == r["F"] and out[v, 1] == r["R2"] and out[v, 6] == r["logLik"]
    assert np.array_equal(out[v, 2:4], r["coef"]) and np.array_equal(out[v, 4:6], r["t"])
    labels = (np.arange(V) % 4).astype(float)
    o2 = oracle.minc2_model_lm(Y, (6, 8, 5), X, mask=labels, lo=2, hi=2)
    assert np.array_equal(o2[:, 0] != 0, labels == 2)
    o3 = oracle.minc2_model_lm(Y, (6, 8, 5), X, mask=labels + 0.4, lo=2, hi=2)   # 2.4 in, 1.4/3.4 out
    assert np.array_equal(o3[:, 0] != 0, labels == 2)
    assert_rows_match(oracle.vertex_lm_loop(Y, X)[inside], out[inside], rtol=0)


def test_randomize_semantics(oracle):
    """Column max over ALL rows (masked zeros included), non-finite -> 0, abs() when two-sided."""
    Y, X, mask = cfg1_data(dims=(4, 5, 5))
    V, n = Y.shape
    rng = np.random.default_rng(4)
    idx = np.column_stack([rng.permutation(n) for _ in range(6)])
    cols = [4, 5]
    d2 = oracle.randomize(Y, X, idx, cols, two_sided=True, mask=mask)
    d1 = oracle.randomize(Y, X, idx, cols, two_sided=False, mask=mask)
    for r in range(idx.shape[1]):
        out = oracle.minc2_model_lm(Y, (4, 5, 5), X[idx[:, r]], mask=mask)
        out[~np.isfinite(out)] = 0
        assert np.array_equal(d2[r], np.abs(out[:, cols]).max(axis=0))
        assert np.array_equal(d1[r], out[:, cols].max(axis=0))
    assert (d1 >= 0).all()   # Q6: masked zeros keep every maximum >= 0 even for "greater"
    ident = np.arange(n)[:, None]
    d0 = oracle.randomize(Y, X, ident, cols, two_sided=True, mask=mask)
    base = oracle.minc2_model_lm(Y, (4, 5, 5), X, mask=mask)
    assert np.array_equal(d0[0], np.abs(base[:, cols]).max(axis=0))


def test_anova_restatement(oracle):
    """voxel_anova (src/minc_anova.c:4-80): sequential F per term; equals the reference's object
    code bit for bit and the nested-model (type I) identity F_t = ((RSS_{t-1}-RSS_t)/df_t)/(RSS/(n-p))."""
    rng = np.random.default_rng(21)
    n, V = 30, 60
    g = np.array(list("abc" * 10))
    age = np.round(rng.normal(60, 10, n), 1)
    X = np.column_stack([np.ones(n), g == "b", g == "c", age, (g == "b") * age, (g == "c") * age]).astype(float)
    asgn = [0, 1, 1, 2, 3, 3]
    Y = rng.normal(2, 1, (V, n))
    f = oracle.anova(Y, X, asgn)
    assert f.shape == (V, 3)

    def rss(Xs, y):
        b = np.linalg.lstsq(Xs, y, rcond=None)[0]
        r = y - Xs @ b
        return r @ r

    for v in (0, 17, 59):
        r = [rss(X[:, :c], Y[v]) for c in (1, 3, 4, 6)]
        want = [((r[i] - r[i + 1]) / d) / (r[3] / (n - 6)) for i, d in enumerate((2, 1, 2))]
        assert f[v] == pytest.approx(want, rel=1e-9)
    mask = (rng.random(V) > 0.4) * rng.integers(1, 3, V).astype(float)
    fm = oracle.anova(Y, X, asgn, mask=mask, lo=2, hi=2)
    assert np.array_equal(fm[mask == 2], f[mask == 2]) and (fm[mask != 2] == 0).all()
    if oracle.have_ref():
        assert np.array_equal(oracle.anova(Y, X, asgn, use_ref=True), f)                          # vertex_anova_loop
        assert np.array_equal(oracle.anova(Y, X, asgn, mask=mask, lo=2, hi=2, use_ref=True, sizes=(3, 4, 5)), fm)  # per_voxel_anova
    g = np.load(os.path.join(GOLD, "ref_golden.npz"))
    assert np.array_equal(oracle.anova(g["D_Y"], g["D_X"], g["D_asgn"], mask=g["D_mask"]), g["D_out"])


def test_weighted_least_squares_restatement(oracle):
    """voxel_wlm (src/weighted_least_squares.c:14-166): equals the reference's object code bit for bit,
    an independent WLS (tests/testthat/test_anatLm.R:121-148 compares with lm(weights=)), and reduces
    to voxel_lm for unit weights."""
    rng = np.random.default_rng(41)
    n, V = 25, 40
    X = np.column_stack([np.ones(n), np.arange(n) % 2, np.round(rng.normal(60, 10, n), 1)])
    w = rng.uniform(0.5, 2.0, n)
    Y = rng.normal(5, 1, (V, n))
    out = oracle.vertex_wlm_loop(Y, X, w)
    if oracle.have_ref():
        assert np.array_equal(out, oracle.vertex_wlm_loop(Y, X, w, use_ref=True))
    W = np.diag(w)
    for v in (0, 13, 39):
        y = Y[v]
        beta = np.linalg.solve(X.T @ W @ X, X.T @ W @ y)
        fitted = X @ beta
        res = y - fitted
        wrss = (w * res ** 2).sum()
        m = (w * fitted).sum() / w.sum()
        mss = (w * (fitted - m) ** 2).sum()
        se = np.sqrt(np.diag(np.linalg.inv(X.T @ W @ X)) * wrss / (n - 3))
        ll = 0.5 * (np.log(w).sum() - n * (np.log(2 * np.pi) + 1 - np.log(n) + np.log(wrss)))
        want = np.r_[(mss / 2) / (wrss / (n - 3)), mss / (mss + wrss), beta, beta / se, ll]
        assert out[v] == pytest.approx(want, rel=1e-9)
    assert np.allclose(oracle.vertex_wlm_loop(Y, X, np.ones(n)), oracle.vertex_lm_loop(Y, X), rtol=1e-12, atol=0)


def test_voxelwise_covariate_restatement(oracle):
    """A file term on the right-hand side (src/vertex_modelling.c:112-140, src/modelling_functions.c:1118-1144):
    the design of voxel v is [Xs | z_v], or [1 | z_v] without a static part.  Equals the reference's object
    code bit for bit, the committed golden vectors, and an independent per-voxel OLS."""
    g = np.load(os.path.join(GOLD, "ref_golden.npz"))
    Y, Z, Xs = g["E_Y"], g["E_Z"], g["E_Xs"]
    a = oracle.lm_loop_cov(Y, Z, Xs)
    b = oracle.lm_loop_cov(Y, Z, None)
    assert np.array_equal(a, g["E_out_static"]) and np.array_equal(b, g["E_out_nostatic"])
    if oracle.have_ref():
        assert np.array_equal(a, oracle.lm_loop_cov(Y, Z, Xs, use_ref=True))
        assert np.array_equal(b, oracle.lm_loop_cov(Y, Z, None, use_ref=True))
    n = Y.shape[1]
    for v in (0, 11, 39):
        for out, X in ((a, np.column_stack([Xs, Z[v]])), (b, np.column_stack([np.ones(n), Z[v]]))):
            w = numpy_ols(Y[v], X)
            p = X.shape[1]
            want = np.r_[w["F"], w["R2"], w["coef"], w["t"], w["logLik"]]
            assert out[v] == pytest.approx(want, rel=1e-8), (v, p)
    # vertex 7 has a constant covariate: negligible column -> rank p-1, beta = 0 and t = 0 in its slot, the other
    # betas are the fit without it (dqrdc2's limited pivoting)
    assert a[7, 4] == 0.0 and a[7, 7] == 0.0
    w = numpy_ols(Y[7], Xs)
    assert a[7, 2:4] == pytest.approx(w["coef"], rel=1e-9)
    # masked rows are 0 in every column
    mask = (np.arange(Y.shape[0]) % 3 != 0).astype(float)
    am = oracle.lm_loop_cov(Y, Z, Xs, mask=mask)
    assert np.array_equal(am[mask > 0.5], a[mask > 0.5]) and (am[mask < 0.5] == 0).all()
    # an all-zero covariate at an unmasked voxel is the reference's dpotri error (exact zero on R's diagonal)
    Z0 = Z.copy()
    Z0[3, :] = 0.0
    with pytest.raises(oracle.OracleError, match="is zero, so the inverse cannot be computed"):
        oracle.lm_loop_cov(Y, Z0, Xs)


def test_stored_type_expansion(oracle):
    """The doubles the reference's read loop produces from stored voxels (libminc's real-range conversion,
    restated: orc_expand_u16): (u - valid_min) * scale[slice, subject] + offset[slice, subject], each operation
    rounded once; float32 converts exactly."""
    rng = np.random.default_rng(31)
    V, n, slice_len = 210, 7, 40                        # 6 slices, the last one short
    U = rng.integers(0, 65536, (V, n)).astype(np.uint16)
    nsl = -(-V // slice_len)
    smin = rng.normal(-1, 0.1, (nsl, n))
    smax = smin + rng.uniform(1, 3, (nsl, n))
    scale, offset = (smax - smin) / 65535.0, smin
    got = oracle.expand_stored(U, scale, offset, voxel_min=0.0, slice_len=slice_len)
    sl = np.arange(V) // slice_len
    want = (U.astype(np.float64) - 0.0) * scale[sl, :] + offset[sl, :]
    assert np.array_equal(got, want)
    assert got.min() >= smin.min() - 1e-12 and got.max() <= smax.max() + 1e-12
    got5 = oracle.expand_stored(U, scale, offset, voxel_min=5.0, slice_len=slice_len)
    assert np.array_equal(got5, (U.astype(np.float64) - 5.0) * scale[sl, :] + offset[sl, :])
    F = rng.normal(size=(V, n)).astype(np.float32)
    assert np.array_equal(oracle.expand_stored(F), F.astype(np.float64))


def test_fdr_restatement(oracle):
    """p.adjust(p, "fdr") and mincFDR's p-values / thresholds (R/minc_FDR.R:385-505).  Known answers: the BH example
    worked by hand, and textbook critical values of the t and F distributions."""
    # BH by hand: p = (0.01, 0.04, 0.03, 0.005) -> sorted 0.005, 0.01, 0.03, 0.04 -> 4/1*.005, 4/2*.01, 4/3*.03, 4/4*.04
    q = oracle.p_adjust_bh([0.01, 0.04, 0.03, 0.005])
    assert q == pytest.approx([0.02, 0.04, 0.04, 0.02], rel=1e-14)
    # NA p-values stay NA and do not count in n
    q = oracle.p_adjust_bh([0.01, np.nan, 0.04, 0.03, 0.005, np.nan])
    assert np.isnan(q[[1, 5]]).all() and q[[0, 2, 3, 4]] == pytest.approx([0.02, 0.04, 0.04, 0.02], rel=1e-14)
    assert oracle.p_adjust_bh([0.9, 0.95]) == pytest.approx([0.95, 0.95])
    # critical values: t_{.975}(10) = 2.228139, t_{.995}(10) = 3.169273, F_{.95}(1,10) = 4.964603, F_{.95}(3,20) = 3.098391
    _, _, p = oracle.fdr([2.228139, -3.169273], "t", 10)
    assert p == pytest.approx([0.05, 0.01], rel=2e-6)
    _, _, p = oracle.fdr([4.964603], "F", 1, 10)
    assert p == pytest.approx([0.05], rel=2e-6)
    _, _, p = oracle.fdr([3.098391], "F", 3, 20)
    assert p == pytest.approx([0.05], rel=2e-6)
    # thresholds: the statistic of the largest accepted p-value; outside the mask q = 1
    rng = np.random.default_rng(3)
    t = np.r_[rng.standard_t(20, 900), rng.normal(6, 1, 100)]
    mask = (np.arange(1000) % 10 != 0).astype(float)
    q, th, p = oracle.fdr(t, "t", 20, mask=mask)
    assert (q[mask < 0.5] == 1).all() and np.all(np.diff(th[~np.isnan(th)]) <= 0)
    acc = np.abs(t[mask > 0.5])[q[mask > 0.5] <= 0.05]
    assert th[1] == pytest.approx(acc.min(), rel=1e-9)       # the weakest accepted |t| is the threshold


def test_fdr_other_methods_restatement(oracle):
    """mincFDR's other methods (R/minc_FDR.R:429-438).  p.adjust(., "BY") against scipy's Benjamini-Yekutieli and a case
    by hand; fast.qvalue's qvalue_min step against the REFERENCE'S OWN object code (src/modelling_functions.c:15-46),
    including its two quirks (pairwise, not running, minimum; the last-ranked element probes its successor in the
    vector); the qvalue package's form is pi0 times Benjamini-Hochberg."""
    from scipy import stats as S
    rng = np.random.default_rng(8)
    # BY by hand: n = 4, c = 1 + 1/2 + 1/3 + 1/4 = 25/12; sorted p .005 .01 .03 .04 -> c*4/i*p = .041667 .041667 .083333 .083333
    q = oracle.p_adjust_by([0.01, 0.04, 0.03, 0.005])
    assert q == pytest.approx([25 / 12 * 0.02, 25 / 12 * 0.04, 25 / 12 * 0.04, 25 / 12 * 0.02], rel=1e-14)
    p = np.r_[rng.random(500), rng.random(60) * 1e-4]
    assert oracle.p_adjust_by(p) == pytest.approx(S.false_discovery_control(p, method="by"), rel=1e-12)
    pn = p.copy()
    pn[::7] = np.nan
    qn = oracle.p_adjust_by(pn)
    assert np.isnan(qn[::7]).all()
    ok = ~np.isnan(pn)
    assert qn[ok] == pytest.approx(S.false_discovery_control(pn[ok], method="by"), rel=1e-12)
    # qvalue_min: restatement == the reference's object code, on raw q-values straddling 1, with ties, with the largest p
    # last in the vector (the reference reads one past the end there) and not last
    if oracle.have_ref():
        for trial in range(20):
            m = int(rng.integers(2, 60))
            pp = np.round(rng.random(m), 2 if trial % 2 else 6)             # rounding makes ties
            if trial % 3 == 0:
                pp[-1] = 1.0                                                 # largest p is the last element
            pi0 = float(rng.uniform(0.3, 1.0))
            u = np.argsort(pp, kind="stable") + 1
            v = np.searchsorted(np.sort(pp), pp, side="right")
            raw = pi0 * m * pp / v * rng.choice([1.0, 3.0], m)               # some far above 1
            assert np.array_equal(oracle.qvalue_min(raw, u), oracle.qvalue_min(raw, u, use_ref=True)), trial
            assert np.array_equal(oracle.fast_qvalue(pp, pi0), oracle.fast_qvalue(pp, pi0, use_ref=True)), trial
    # by hand: p = (.5, .01, .02), pi0 = 1, m = 3: raw q = 3 p / rank = (.5, .03, .03); u = (2, 3, 1)
    #   rank-last (element 1): probes qvalue[u[2]] = qvalue[1] (0-based) = .03, not above 1 -> keeps .5
    #   element 3: min(.03, .5) = .03; element 2: min(.03, .03) = .03
    assert oracle.fast_qvalue([0.5, 0.01, 0.02], 1.0) == pytest.approx([0.5, 0.03, 0.03], rel=1e-15)
    # NOT a running minimum: p = (.001, .5, .6, .9) * ..., raw = 4p/rank = (.004, 1, .8, .9): element 2 takes min(1, .8),
    # element 1 takes min(.004, 1 [raw, not .8])
    assert oracle.fast_qvalue([0.001, 0.5, 0.6, 0.9], 1.0) == pytest.approx([0.004, 0.8, 0.8, 0.9], rel=1e-15)
    assert oracle.qvalue_package(p, 0.7) == pytest.approx(0.7 * oracle.p_adjust_bh(p), rel=1e-15)
    assert oracle.pi0_fractions([0.1, 0.5, 0.5, 0.9], [0.0, 0.5, 0.55]) == pytest.approx([1.0, 0.75, 0.25])
    # thresholds come from max(p[q <= level]) (a set, not a prefix: fast.qvalue's q is not monotone in p)
    t = np.r_[rng.standard_t(30, 2000), rng.normal(5, 1, 200)]
    for method, pi0 in (("BY", 1.0), ("pFDR", 0.85), ("qvalue", 0.85)):
        q, th, pv = oracle.fdr(t, "t", 30, method=method, pi0=pi0)
        for j, lev in enumerate(oracle.FDR_LEVELS):
            acc = np.abs(t)[q <= lev]
            assert (np.isnan(th[j]) and acc.size == 0) or th[j] == pytest.approx(acc.min(), rel=1e-9)


def test_smoothing_spline_for_pi0():
    """fast.qvalue's pi0 = predict(smooth.spline(lambda, pi0, df = 3), max(lambda)) (R/minc_misc.R:446-447): the host
    smoother of the API layer against scipy's independent smoothing-spline solver at the same penalty, and its trace."""
    from scipy.interpolate import make_smoothing_spline
    from rminc_b200 import api
    rng = np.random.default_rng(4)
    lam = np.arange(0, 19) * 0.05
    y = 0.8 + 0.1 * lam ** 2 + 0.01 * rng.standard_normal(19)
    got = api.smooth_spline_at_last_knot(lam, y, 3.0)
    # find the penalty whose smoother matrix has trace 3 with scipy alone (columns of the hat matrix = fits of unit vectors)
    def trace(l):
        return sum(make_smoothing_spline(lam, np.eye(19)[i], lam=l)(lam[i]) for i in range(19))
    lo, hi = -30.0, 10.0
    for _ in range(80):
        mid = 0.5 * (lo + hi)
        lo, hi = (mid, hi) if trace(np.exp(mid)) > 3.0 else (lo, mid)
    want = float(make_smoothing_spline(lam, y, lam=np.exp(0.5 * (lo + hi)))(lam[-1]))
    assert got == pytest.approx(want, rel=1e-8)
    # a straight line is reproduced exactly (the penalty does not see it); df -> n interpolates
    assert api.smooth_spline_at_last_knot(lam, 0.3 + 0.5 * lam, 3.0) == pytest.approx(0.3 + 0.5 * 0.9, rel=1e-12)
    assert api.smooth_spline_at_last_knot(lam, y, 19.0) == pytest.approx(y[-1], rel=1e-6)


def _icosphere():
    """A closed triangle mesh (icosahedron, each face split once): 42 vertices, 80 consistently oriented triangles."""
    t = (1 + 5 ** 0.5) / 2
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1),
         (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
         (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    verts = [np.array(p, float) / np.linalg.norm(p) for p in v]
    mid, faces = {}, []

    def m(a, b):
        k = (min(a, b), max(a, b))
        if k not in mid:
            p = verts[a] + verts[b]
            verts.append(p / np.linalg.norm(p))
            mid[k] = len(verts) - 1
        return mid[k]
    for a, b, c in f:
        ab, bc, ca = m(a, b), m(b, c), m(c, a)
        faces += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
    return np.array(verts).T * [[1.0], [1.3], [0.8]], np.array(faces).T + 1


def test_mesh_area_is_the_reference_object_code(oracle, lib):
    """RMINC:::mesh_area (src/vertex_tfce.cpp:221-257), vertexTFCE's default weights for a bic_obj surface: the product's
    host routine and the restatement equal the reference's compiled function bit for bit -- including its first
    cross-product component, written d1y*d2z - d1z*d2x."""
    from rminc_b200 import api, core
    vm, tm = _icosphere()
    rng = np.random.default_rng(3)
    vm = vm + 0.01 * rng.standard_normal(vm.shape)
    got = core.mesh_area(vm, tm)
    assert np.array_equal(got, oracle.mesh_area(vm, tm))
    if oracle.have_ref_tfce():
        assert np.array_equal(got, oracle.mesh_area(vm, tm, use_ref=True))
    # the literal formula is NOT the true area (true: |d1 x d2| / 2); the quirk is restated, not repaired
    d1, d2 = vm[:, tm[1] - 1] - vm[:, tm[0] - 1], vm[:, tm[2] - 1] - vm[:, tm[0] - 1]
    assert not np.allclose(got.sum(), 0.5 * np.linalg.norm(np.cross(d1.T, d2.T), axis=1).sum(), rtol=1e-3)
    with pytest.raises(core.RmincGpuError, match="refers to vertex"):
        core.mesh_area(vm, np.array([[1], [2], [99]]))
    # obj_to_graph's edge list as written: every 1-2 and 1-3 edge, plus the 2-3 edge of the first triangle only
    adj = api.obj_to_adjacency(api.BicObj(vm, tm))
    edges = {tuple(sorted(e)) for t in (tm - 1).T for e in ((t[0], t[1]), (t[0], t[2]))} | {tuple(sorted((tm[1, 0] - 1, tm[2, 0] - 1)))}
    assert {(a, int(b)) for a, nb in enumerate(adj) for b in nb if a < b} == edges
    assert all(a in adj[int(b)] for a, nb in enumerate(adj) for b in nb)          # symmetric, as rmg_graph_tfce requires


def test_graph_tfce_restatement(oracle):
    """graph_tfce_wqu / graph_tfce (src/vertex_tfce.cpp:92-157): equals the reference's own object code bit for bit,
    and a case small enough to enhance by hand."""
    from helpers import grid_adjacency
    # path graph 0-1-2-3, map (2, 1, 0, 1), nsteps = 2, E = 1, H = 1, unit weights: thresholds 2, 1, 0 (dh = 1)
    #   t = 2: {0}                 -> v0 += 2 * 1 * 1
    #   t = 1: {0,1} and {3}       -> v0, v1 += 1 * 2 * 1 ; v3 += 1 * 1 * 1
    #   t = 0: everything          -> += 0
    ptr, idx = oracle.adjacency_csr([[1], [0, 2], [1, 3], [2]])
    got = oracle.graph_tfce([2.0, 1.0, 0.0, 1.0], ptr, idx, E=1.0, H=1.0, nsteps=2, side="positive")
    assert got == pytest.approx([4.0, 2.0, 0.0, 1.0], abs=1e-15)
    both = oracle.graph_tfce([2.0, 1.0, 0.0, -1.0], ptr, idx, E=1.0, H=1.0, nsteps=2, side="both")
    assert both == pytest.approx([4.0, 2.0, 0.0, -0.5 * 1 * 0.5 - 1.0 * 1 * 0.5], abs=1e-15)   # -(t=1: 1*1*.5 + t=.5: .5*1*.5)
    rng = np.random.default_rng(12)
    adj = grid_adjacency(25, 30)
    ptr, idx = oracle.adjacency_csr(adj)
    x = rng.standard_normal(750)
    x[100:130] += 3.0
    x[400] = 0.0
    w = rng.uniform(0.5, 2.0, 750)
    for side in ("both", "positive", "negative"):
        for weights in (None, w):
            for E, H, ns in ((0.5, 2.0, 100), (1.0, 1.0, 37)):
                a = oracle.graph_tfce(x, ptr, idx, E=E, H=H, nsteps=ns, side=side, weights=weights)
                if oracle.have_ref_tfce():
                    assert np.array_equal(a, oracle.graph_tfce(x, ptr, idx, E=E, H=H, nsteps=ns, side=side, weights=weights,
                                                               use_ref=True))
                assert (a >= 0).all() if side == "positive" else True
    assert (oracle.graph_tfce(-np.abs(x) - 1, ptr, idx, side="positive") == 0).all()        # hmax <= 0: all zero


def test_volume_tfce_restatement_agrees_with_the_graph_algorithm_on_a_neighbour_list(oracle, lib):
    """The reference pins mincTFCE only through vertexTFCE on RMINC:::neighbour_list(.., 6) with weights = voxel volume
    (tests/testthat/test_mincTFCE.R:153-170).  The scipy-labelling restatement of the volume enhancement and the
    restated graph_tfce (pinned to the reference's object code) must then agree when the graph algorithm is given the
    same thresholds: with hmax an exact multiple of d, nsteps = hmax / d makes graph_tfce's thresholds hmax - k * d."""
    from rminc_b200 import core
    rng = np.random.default_rng(5)
    x, y, z = 7, 6, 5                                    # neighbour_list's (fastest, middle, slowest)
    dims = (z, y, x)                                      # file order: slowest first
    vol = np.round(rng.normal(0, 1, x * y * z), 2)
    d = 0.25
    vol[17] = 3.0                                         # hmax = 12 d exactly
    vol[vol > 3.0] = 2.5
    nl = oracle.neighbour_list(x, y, z, 6)
    ptr, idx = oracle.adjacency_csr(nl)
    got = oracle.volume_tfce(vol, dims, d=d, side="positive", voxel_volume=0.001)
    want = oracle.graph_tfce(vol, ptr, idx, nsteps=12, side="positive", weights=np.full(vol.size, 0.001))
    assert np.allclose(got, want, rtol=1e-9, atol=1e-15)
    # the product's neighbour_list (host code, no GPU) is the reference's, for every connectivity
    for n in (6, 18, 26):
        p2, i2 = core.neighbour_list(4, 3, 5, n)
        want_nl = oracle.neighbour_list(4, 3, 5, n)
        assert [sorted(i2[p2[v]:p2[v + 1]].tolist()) for v in range(60)] == [sorted(w) for w in want_nl]
    with pytest.raises(core.RmincGpuError, match="connectivity must be 6,18, or 26"):
        core.neighbour_list(3, 3, 3, 8)


def test_grid_tfce_golden_vectors_of_the_reference_object_code(oracle, lib):
    """Golden case F (tests/golden/make_golden.py): the reference's own compiled neighbour_list and graph_tfce on a
    6 x 5 x 4 voxel grid with weights = voxel volume -- the enhancement the reference's tests equate with mincTFCE
    (tests/testthat/test_mincTFCE.R:153-170).  Pins (a) the restated neighbour_list, (b) the product's
    rmg_neighbour_list (host code), (c) the scipy-labelling restatement of the volume enhancement, for every
    connectivity and side; and, where the reference is present, the compiled function itself once more."""
    from rminc_b200 import core
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_golden.npz"))
    z, y, x = (int(v) for v in G["F_dims"])
    vol = G["F_vol"]
    for conn in (6, 18, 26):
        ptr, idx = G[f"F_ptr{conn}"], G[f"F_idx{conn}"]
        want_nl = [sorted(idx[ptr[v]:ptr[v + 1]].tolist()) for v in range(x * y * z)]
        assert [sorted(a) for a in oracle.neighbour_list(x, y, z, conn)] == want_nl
        p2, i2 = core.neighbour_list(x, y, z, conn)
        assert [sorted(i2[p2[v]:p2[v + 1]].tolist()) for v in range(x * y * z)] == want_nl
        if oracle.have_ref_tfce():
            assert [sorted(a) for a in oracle.neighbour_list(x, y, z, conn, use_ref=True)] == want_nl
        for side in ("positive", "negative", "both"):
            want = G[f"F_tfce{conn}_{side}"]
            got = oracle.volume_tfce(vol, (z, y, x), d=0.25, side=side, connectivity=conn, voxel_volume=0.001)
            assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max(), (conn, side)
            assert np.array_equal(got == 0, want == 0)
            # and the restated graph algorithm on the golden adjacency, with the reference's nsteps
            g2 = oracle.graph_tfce(vol, ptr, idx, nsteps=8, side=side, weights=np.full(vol.size, 0.001))
            assert np.array_equal(g2, want), (conn, side)

