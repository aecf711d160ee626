"""The host-side mirror of RMINC's R interface (rminc_b200/api.py), tested the way the reference's
testthat files test mincLm / vertexLm / anatLm / mincRandomize: one row of the result against an
independent lm() of the same data (tests/testthat/test_mincLm.R:34-56, :135-208,
test_vertexLm.R:66-74, test_anatLm.R:30-44, :152-157).

Each test runs twice: with the numerical back end replaced by the CPU oracle (host logic only, no
GPU needed) and, marked gpu, through the real CUDA path."""
import numpy as np
import pandas as pd
import pytest

from test_oracle import numpy_ols


@pytest.fixture(params=["oracle-backend", pytest.param("cuda", marks=pytest.mark.gpu)])
def api(request, monkeypatch, oracle, lib):
    from rminc_b200 import api as A
    from rminc_b200 import core
    if request.param == "oracle-backend":
        def lm_fit(Y, X, mask=None, mask_lo=1.0, mask_hi=99999999.0, device=0, n_gpus=None):
            return oracle.minc2_model_lm(Y, (1, 1, Y.shape[0]), X, mask=mask, lo=mask_lo, hi=mask_hi)

        def vertex_lm(Y, X, device=0):
            return oracle.vertex_lm_loop(Y, X)

        def randomize(Y, X, idx, cols, two_sided=True, mask=None, mask_lo=1.0, mask_hi=99999999.0, device=0,
                      n_gpus=None):
            return oracle.randomize(Y, X, idx, cols, two_sided=two_sided, mask=mask, lo=mask_lo, hi=mask_hi)

        def anova_fit(Y, X, asgn, mask=None, mask_lo=1.0, mask_hi=99999999.0, device=0):
            return oracle.anova(Y, X, asgn, mask=mask, lo=mask_lo, hi=mask_hi)

        def vertex_wlm(Y, X, weights, device=0):
            return oracle.vertex_wlm_loop(Y, X, weights)

        def lm_fit_cov(Y, Z, Xs=None, mask=None, mask_lo=1.0, mask_hi=99999999.0, device=0, n_gpus=None):
            return oracle.lm_loop_cov(Y, Z, Xs, mask=mask, lo=mask_lo, hi=mask_hi)

        def lm_fit_stored(U, X, scale=None, offset=None, voxel_min=0.0, slice_len=None, mask=None, mask_lo=1.0,
                          mask_hi=99999999.0, device=0, n_gpus=None):
            Y = oracle.expand_stored(U, scale, offset, voxel_min=voxel_min, slice_len=slice_len)
            return oracle.minc2_model_lm(Y, (1, 1, Y.shape[0]), X, mask=mask, lo=mask_lo, hi=mask_hi)

        def fdr(stat, stat_type, df1, df2=0.0, mask=None, device=0, method="FDR", pi0=1.0):
            q, th, _ = oracle.fdr(stat, "t" if stat_type == core.STAT_T else "F", df1, df2, mask=mask, method=method, pi0=pi0)
            return q, th

        def fdr_pi0_fractions(stat, stat_type, df1, df2=0.0, mask=None, lambdas=None, device=0):
            _, _, p = oracle.fdr(stat, "t" if stat_type == core.STAT_T else "F", df1, df2, mask=mask)
            return oracle.pi0_fractions(p, lambdas)

        def graph_tfce(maps, adjacency, E=0.5, H=2.0, nsteps=100, side="both", weights=None, device=0):
            ptr, aidx = core.adjacency_csr(adjacency)
            M = np.asarray(maps, dtype=np.float64)
            if M.ndim == 1:
                return oracle.graph_tfce(M, ptr, aidx, E=E, H=H, nsteps=nsteps, side=side, weights=weights)
            return np.column_stack([oracle.graph_tfce(M[:, j], ptr, aidx, E=E, H=H, nsteps=nsteps, side=side, weights=weights)
                                    for j in range(M.shape[1])])

        def randomize_tfce(Y, X, idx, cols, adjacency, two_sided=True, E=0.5, H=2.0, nsteps=100, side="both", weights=None,
                           device=0, n_gpus=None):
            from helpers import tfce_randomize_oracle
            return tfce_randomize_oracle(oracle, np.asfortranarray(Y), X, np.asarray(idx), list(cols), adjacency,
                                         two_sided=two_sided, E=E, H=H, nsteps=nsteps, side=side, weights=weights)

        def randomize_stored(U, X, idx, cols, scale=None, offset=None, voxel_min=0.0, slice_len=None, two_sided=True,
                             mask=None, mask_lo=1.0, mask_hi=99999999.0, device=0, n_gpus=None):
            Y = oracle.expand_stored(U, scale, offset, voxel_min=voxel_min, slice_len=slice_len)
            return oracle.randomize(Y, X, idx, cols, two_sided=two_sided, mask=mask, lo=mask_lo, hi=mask_hi)

        def volume_tfce(maps, dims, d=0.1, E=0.5, H=2.0, side="both", connectivity=6, voxel_volume=1.0, device=0):
            M = np.asarray(maps, dtype=np.float64)
            kw = dict(d=d, E=E, H=H, side=side, connectivity=connectivity, voxel_volume=voxel_volume)
            if M.ndim == 1:
                return oracle.volume_tfce(M, dims, **kw)
            return np.column_stack([oracle.volume_tfce(M[:, j], dims, **kw) for j in range(M.shape[1])])

        def randomize_volume_tfce(Y, X, idx, cols, dims, two_sided=True, d=0.1, E=0.5, H=2.0, side="both", connectivity=6,
                                  voxel_volume=1.0, mask=None, mask_lo=1.0, mask_hi=99999999.0, n_gpus=None):
            return oracle.randomize_volume_tfce(Y, X, idx, cols, dims, two_sided=two_sided, mask=mask, lo=mask_lo, hi=mask_hi,
                                                d=d, E=E, H=H, side=side, connectivity=connectivity, voxel_volume=voxel_volume)

        monkeypatch.setattr(core, "volume_tfce", volume_tfce)
        monkeypatch.setattr(core, "randomize_volume_tfce", randomize_volume_tfce)
        monkeypatch.setattr(core, "randomize_stored", randomize_stored)
        monkeypatch.setattr(core, "graph_tfce", graph_tfce)
        monkeypatch.setattr(core, "randomize_tfce", randomize_tfce)
        monkeypatch.setattr(core, "fdr", fdr)
        monkeypatch.setattr(core, "fdr_pi0_fractions", fdr_pi0_fractions)
        monkeypatch.setattr(core, "lm_fit_stored", lm_fit_stored)
        monkeypatch.setattr(core, "lm_fit_cov", lm_fit_cov)
        monkeypatch.setattr(core, "vertex_wlm", vertex_wlm)
        monkeypatch.setattr(core, "anova_fit", anova_fit)
        monkeypatch.setattr(core, "lm_fit", lm_fit)
        monkeypatch.setattr(core, "vertex_lm", vertex_lm)
        monkeypatch.setattr(core, "randomize", randomize)

        def randomize_cov(Y, Z, Xs, idx, cols, two_sided=True, mask=None, mask_lo=1.0, mask_hi=99999999.0, device=0,
                          n_gpus=None):
            return oracle.randomize_cov(Y, Z, Xs, idx, cols, two_sided=two_sided, mask=mask, lo=mask_lo, hi=mask_hi)
        monkeypatch.setattr(core, "randomize_cov", randomize_cov)
    return A


@pytest.fixture()
def study(tmp_path):
    """10 subjects of 10x10x10 volumes, as tests/testthat/test_mincLm.R:12-23."""
    rng = np.random.default_rng(42)
    n, dims = 10, (10, 10, 10)
    sex = np.array(list("MFMFMFMFMF"))
    scale = np.round(rng.normal(1.0, 0.2, n), 3)
    coil = np.array([1, 2, 3, 1, 2, 3, 1, 2, 3, 1])
    grp = np.array(list("abcabcabca"))
    files = []
    for i in range(n):
        vol = rng.normal(0, 0.2, dims) + 0.3 * (sex[i] == "M") + 0.5 * scale[i]
        f = tmp_path / f"subj{i:02d}.npy"
        np.save(f, vol)
        files.append(str(f))
    gf = pd.DataFrame(dict(jacobians=files, Sex=sex, Scale=scale, Coil=coil, Grp=grp))
    mask = np.zeros(dims)
    mask[2:8, 2:8, 2:8] = 1
    np.save(tmp_path / "mask.npy", mask)
    return gf, str(tmp_path / "mask.npy"), np.stack([np.load(f).reshape(-1) for f in files], axis=1)


def check_row(res, v, y, X, names):
    w = numpy_ols(y, X)
    p = X.shape[1]
    row = np.asarray(res)[v]
    assert row[0] == pytest.approx(w["F"], rel=1e-9)
    assert row[1] == pytest.approx(w["R2"], rel=1e-9)
    assert row[2:2 + p] == pytest.approx(w["coef"], rel=1e-8, abs=1e-10)
    assert row[2 + p:2 + 2 * p] == pytest.approx(w["t"], rel=1e-8, abs=1e-10)
    assert row[-1] == pytest.approx(w["logLik"], rel=1e-10)
    assert res.colnames == (["F-statistic", "R-squared"] + [f"beta-{r}" for r in names]
                            + [f"tvalue-{r}" for r in names] + ["logLik"])


def test_mincLm_two_level_factor(api, study):
    gf, _, Y = study
    vs = api.mincLm("jacobians ~ Sex", gf)
    X = np.column_stack([np.ones(10), gf.Sex == "M"]).astype(float)
    assert vs.shape == (1000, 7) and vs.r_class == ("mincLm", "mincMultiDim", "matrix")
    check_row(vs, 0, Y[0], X, ["(Intercept)", "SexM"])
    check_row(vs, 567, Y[567], X, ["(Intercept)", "SexM"])
    # attributes as R/minc_voxel_statistics.R:908-935
    assert vs.attrs["likeVolume"] == gf.jacobians[0] and vs.attrs["filenames"] == list(gf.jacobians)
    assert np.array_equal(vs.attrs["model"], X)
    assert vs.attrs["stat-type"] == ["F", "R-squared", "beta", "beta", "t", "t", "logLik"]
    assert vs.attrs["df"] == [[1, 8], 8, 8]
    # AIC / BIC consumers use the logLik column and the model attribute (R/minc_model_selection.R:1-68)
    ll, k = vs.col("logLik")[0], X.shape[1] + 1
    w = numpy_ols(Y[0], X)
    assert -2 * ll + 2 * k == pytest.approx(-2 * w["logLik"] + 2 * k)


def test_mincLm_interaction_three_level_and_two_covariates(api, study):
    gf, _, Y = study
    m, s = (gf.Sex == "M").astype(float), gf.Scale.values
    vs = api.mincLm("jacobians ~ Sex*Scale", gf)
    check_row(vs, 1, Y[1], np.column_stack([np.ones(10), m, s, m * s]), ["(Intercept)", "SexM", "Scale", "SexM:Scale"])
    vs = api.mincLm("jacobians ~ Grp", gf)
    Xg = np.column_stack([np.ones(10), gf.Grp == "b", gf.Grp == "c"]).astype(float)
    check_row(vs, 1, Y[1], Xg, ["(Intercept)", "Grpb", "Grpc"])
    assert vs.attrs["df"] == [[2, 7], 7, 7, 7]
    vs = api.mincLm("jacobians ~ Scale*Coil", gf.iloc[:10])
    c = gf.Coil.values.astype(float)
    check_row(vs, 999, Y[999], np.column_stack([np.ones(10), s, c, s * c]), ["(Intercept)", "Scale", "Coil", "Scale:Coil"])


def test_mincLm_mask_maskval_subset(api, study):
    gf, maskfile, Y = study
    vs = api.mincLm("jacobians ~ Sex", gf, mask=maskfile)
    inside = np.load(maskfile).reshape(-1) > 0.5
    arr = np.asarray(vs)
    assert (arr[~inside] == 0).all() and (arr[inside, 0] != 0).all()
    labels = (np.arange(1000) % 3).astype(float)
    vs2 = api.mincLm("jacobians ~ Sex", gf, mask=labels, maskval=2)
    assert np.array_equal(np.asarray(vs2)[:, 0] != 0, labels == 2)
    # subset: rows of the data frame, drop.unused.levels
    keep = np.arange(10) != 3
    vs3 = api.mincLm("jacobians ~ Sex", gf, subset=keep)
    X = np.column_stack([np.ones(9), gf.Sex[keep] == "M"]).astype(float)
    check_row(vs3, 5, Y[5][keep], X, ["(Intercept)", "SexM"])
    with pytest.raises(ValueError, match="dimension mismatch"):
        api.mincLm("jacobians ~ Sex", gf, mask=np.ones(999))
    bad = gf.copy()
    bad.loc[2, "jacobians"] = "/nonexistent/file.npy"
    with pytest.raises(FileNotFoundError, match="Error opening input file"):
        api.mincLm("jacobians ~ Sex", bad)


def test_vertexLm_and_anatLm(api, tmp_path):
    rng = np.random.default_rng(7)
    n, V = 12, 40
    dx = np.array(["AD", "CN"] * 6)
    age = np.round(rng.normal(70, 5, n), 1)
    files = []
    for i in range(n):
        f = tmp_path / f"thick{i}.txt"
        np.savetxt(f, np.column_stack([rng.normal(3, 0.3, V) + 0.2 * (dx[i] == "CN"), rng.normal(size=V)]))
        files.append(str(f))
    gf = pd.DataFrame(dict(thickness=files, Dx=dx, Age=age))
    Y = np.column_stack([np.loadtxt(f)[:, 0] for f in files])
    res = api.vertexLm("thickness ~ Dx + Age", gf)
    X = np.column_stack([np.ones(n), dx == "CN", age]).astype(float)
    assert res.r_class == ("vertexLm", "vertexMultiDim", "matrix") and res.shape == (V, 9)
    check_row(res, 0, Y[0], X, ["(Intercept)", "DxCN", "Age"])
    res2 = api.vertexLm("thickness ~ Dx + Age", gf, column=2)
    Y2 = np.column_stack([np.loadtxt(f)[:, 1] for f in files])
    check_row(res2, 3, Y2[3], X, ["(Intercept)", "DxCN", "Age"])
    # anatLm: subjects x structures matrix, rows of the result are structures
    anat = pd.DataFrame(rng.normal(10, 1, (n, 5)) + 0.5 * (dx == "CN")[:, None], columns=list("ABCDE"))
    ra = api.anatLm("~ Dx", gf, anat)
    assert ra.r_class == ("anatModel", "matrix") and ra.rownames == list("ABCDE") and ra.shape == (5, 7)
    check_row(ra, 2, anat.values[:, 2], X[:, :2], ["(Intercept)", "DxCN"])
    # weighted anatLm vs an independent WLS (tests/testthat/test_anatLm.R:121-148)
    w = rng.uniform(0.5, 2.0, n)
    rw = api.anatLm("~ Dx", gf, anat, weights=w)
    Xw, yv = X[:, :2], anat.values[:, 3]
    W = np.diag(w)
    beta = np.linalg.solve(Xw.T @ W @ Xw, Xw.T @ W @ yv)
    res = yv - Xw @ beta
    wrss = (w * res ** 2).sum()
    se = np.sqrt(np.diag(np.linalg.inv(Xw.T @ W @ Xw)) * wrss / (n - 2))
    assert np.asarray(rw)[3, 2:4] == pytest.approx(beta, rel=1e-9)
    assert np.asarray(rw)[3, 4:6] == pytest.approx(beta / se, rel=1e-9)
    assert np.asarray(rw)[3, 6] == pytest.approx(0.5 * (np.log(w).sum() - n * (np.log(2 * np.pi) + 1 - np.log(n) + np.log(wrss))), rel=1e-10)
    with pytest.raises(ValueError, match="Incorrect number of weights"):
        api.anatLm("~ Dx", gf, anat, weights=np.ones(n - 1))
    gm = gf.copy()
    gm["M"] = [np.ones(3)] * n
    with pytest.raises(ValueError, match="A term in your model is a matrix"):
        api.anatLm("~ M", gm, anat)                                    # tests/testthat/test_anatLm.R:152-157


def test_mincRandomize_and_thresholds(api, study):
    gf, maskfile, Y = study
    vs = api.mincLm("jacobians ~ Sex", gf, mask=maskfile)
    rand = api.mincRandomize(vs, R=9, seed=3)
    th = api.thresholds(rand)
    assert th.shape == (4, 2)                                          # tests/testthat/test_mincLm.R:261-267
    assert th.rownames == ["1%", "5%", "10%", "20%"] and th.colnames == ["tvalue-(Intercept)", "tvalue-SexM"]
    assert rand.r_class == ("mincLm_randomization", "minc_randomization")
    assert rand["args"]["alternative"] == "two.sided"
    assert np.array_equal(np.asarray(rand["stats"]), np.asarray(vs)[:, [4, 5]])
    # every row of the null distribution is the column max of |t| of a refit with permuted predictors
    idx = api.draw_indices(10, 9, seed=3)
    X = vs.attrs["model"]
    inside = np.load(maskfile).reshape(-1) > 0.5
    for r in (0, 8):
        tmax = np.zeros(2)
        for v in np.flatnonzero(inside):
            tmax = np.maximum(tmax, np.abs(numpy_ols(Y[v], X[idx[:, r]])["t"]))
        assert np.asarray(rand["randomization_dist"])[r] == pytest.approx(tmax, rel=1e-8)
    one = api.mincRandomize(vs, R=5, alternative="greater", seed=3, columns=[5])
    assert np.asarray(one["randomization_dist"]).shape == (5, 1) and (np.asarray(one["randomization_dist"]) >= 0).all()
    with pytest.raises(ValueError):
        api.mincRandomize(vs, R=3, alternative="less")


def test_mincAnova_vertexAnova_anatAnova(api, study, tmp_path):
    """tests/testthat/test_mincAnova.R:22-109 pattern: F per term at one voxel == anova(lm(...))."""
    gf, maskfile, Y = study
    res = api.mincAnova("jacobians ~ Sex*Scale", gf, mask=maskfile)
    assert res.colnames == ["Sex", "Scale", "Sex:Scale"] and res.shape == (1000, 3)
    assert res.attrs["stat-type"] == ["F", "F", "F"] and res.attrs["df"] == [[1, 6], [1, 6], [1, 6]]
    m, s = (gf.Sex == "M").astype(float).values, gf.Scale.values
    X = np.column_stack([np.ones(10), m, s, m * s])

    def rss(Xs, y):
        b = np.linalg.lstsq(Xs, y, rcond=None)[0]
        r = y - Xs @ b
        return r @ r

    inside = np.load(maskfile).reshape(-1) > 0.5
    v = int(np.flatnonzero(inside)[7])
    r = [rss(X[:, :c], Y[v]) for c in (1, 2, 3, 4)]
    want = [(r[i] - r[i + 1]) / (r[3] / 6) for i in range(3)]
    assert np.asarray(res)[v] == pytest.approx(want, rel=1e-8)
    assert (np.asarray(res)[~inside] == 0).all()
    res3 = api.mincAnova("jacobians ~ Grp", gf)
    assert res3.colnames == ["Grp"] and res3.attrs["df"] == [[2, 7]]
    Xg = np.column_stack([np.ones(10), gf.Grp == "b", gf.Grp == "c"]).astype(float)
    assert np.asarray(res3)[3, 0] == pytest.approx(((rss(Xg[:, :1], Y[3]) - rss(Xg, Y[3])) / 2) / (rss(Xg, Y[3]) / 7), rel=1e-8)
    # vertexAnova / anatAnova share the native loop
    rng = np.random.default_rng(9)
    anat = pd.DataFrame(rng.normal(10, 1, (10, 4)) + 0.5 * m[:, None], columns=list("ABCD"))
    ra = api.anatAnova("~ Sex + Scale", gf, anat)
    assert ra.rownames == list("ABCD") and ra.colnames == ["Sex", "Scale"] and ra.r_class == ("anatModel", "matrix")
    X2 = np.column_stack([np.ones(10), m, s])
    y = anat.values[:, 1]
    assert np.asarray(ra)[1] == pytest.approx([(rss(X2[:, :1], y) - rss(X2[:, :2], y)) / (rss(X2, y) / 7),
                                               (rss(X2[:, :2], y) - rss(X2, y)) / (rss(X2, y) / 7)], rel=1e-8)
    files = []
    for i in range(10):
        f = tmp_path / f"v{i}.txt"
        np.savetxt(f, anat.values[i])
        files.append(str(f))
    gv = gf.assign(thick=files)
    rv = api.vertexAnova("thick ~ Sex + Scale", gv)
    assert np.array_equal(np.asarray(rv), np.asarray(ra)) and rv.r_class == ("vertexMultiDim", "matrix")


def test_file_covariate_on_the_right_hand_side(api, study, tmp_path):
    """mincLm(jacobians ~ Sex + otherVolume): parseLmFormula finds the file term (R/minc_interface.R:1266-1324) and
    the per-voxel design is [model.matrix(~ Sex) | z_v]; with the file as the only term it is [1 | z_v]."""
    gf, maskfile, Y = study
    rng = np.random.default_rng(11)
    n = len(gf)
    zfiles = []
    for i in range(n):
        f = tmp_path / f"cov{i:02d}.npy"
        np.save(f, rng.normal(50, 4, (10, 10, 10)))
        zfiles.append(str(f))
    gf = gf.assign(other=zfiles)
    Z = np.stack([np.load(f).reshape(-1) for f in zfiles], axis=1)
    sexm = (gf.Sex == "M").astype(float).values
    vs = api.mincLm("jacobians ~ Sex + other", gf)
    assert vs.shape == (1000, 9) and vs.r_class == ("mincLm", "mincMultiDim", "matrix")
    for v in (0, 431, 999):
        check_row(vs, v, Y[v], np.column_stack([np.ones(n), sexm, Z[v]]), ["(Intercept)", "SexM", "other"])
    # the reference derives model / df from the static part alone (R/minc_voxel_statistics.R:908, :929-935)
    assert np.array_equal(vs.attrs["model"], np.column_stack([np.ones(n), sexm]))
    assert vs.attrs["df"] == [[1, 8], 8, 8, 8]
    assert vs.attrs["stat-type"] == ["F", "R-squared"] + ["beta"] * 3 + ["t"] * 3 + ["logLik"]
    # term order does not matter; mask honoured
    vm = api.mincLm("jacobians ~ other + Sex", gf, mask=maskfile)
    inside = np.load(maskfile).reshape(-1) > 0.5
    assert np.allclose(np.asarray(vm)[inside], np.asarray(vs)[inside], rtol=1e-12) and (np.asarray(vm)[~inside] == 0).all()
    # mincRandomize re-evaluates the whole call with the predictor rows re-sampled: the covariate FILES move with Sex
    # (R/minc_voxel_statistics.R:1468-1477), so every row of the null is the column max of |t| of a refit on
    # [1, Sex, z_v][idx, ]
    rz = api.mincRandomize(vm, R=5, seed=7)
    assert rz["randomization_dist"].colnames == ["tvalue-(Intercept)", "tvalue-SexM", "tvalue-other"]
    idx = api.draw_indices(n, 5, seed=7)
    for r in (0, 4):
        tmax = np.zeros(3)
        for v in np.flatnonzero(inside):
            Xv = np.column_stack([np.ones(n), sexm, Z[v]])[idx[:, r]]
            tmax = np.maximum(tmax, np.abs(numpy_ols(Y[v], Xv)["t"]))
        assert np.asarray(rz["randomization_dist"])[r] == pytest.approx(tmax, rel=1e-7)
    # the file as the only term: [1 | z], coefficient names 'Intercept' and the term (:1270)
    v1 = api.mincLm("jacobians ~ other", gf)
    assert v1.shape == (1000, 7)
    check_row(v1, 77, Y[77], np.column_stack([np.ones(n), Z[77]]), ["Intercept", "other"])
    assert v1.attrs["df"] == [[0, 0], 0, 0]
    with pytest.raises(api.FormulaError, match="Only \\+ sign allowed"):
        api.mincLm("jacobians ~ Sex * other", gf)
    with pytest.raises(api.FormulaError, match="Only 2 terms allowed"):
        api.mincLm("jacobians ~ Sex + Scale + other", gf)
    # vertexLm with a vertex-file covariate (R/minc_vertex_statistics.R:393-401)
    V = 30
    yf, zf = [], []
    for i in range(n):
        a, b = tmp_path / f"y{i}.txt", tmp_path / f"z{i}.txt"
        np.savetxt(a, rng.normal(3, 0.3, V))
        np.savetxt(b, rng.normal(1, 0.1, V))
        yf.append(str(a))
        zf.append(str(b))
    vf = pd.DataFrame(dict(thick=yf, Sex=gf.Sex.values, curv=zf))
    rv = api.vertexLm("thick ~ Sex + curv", vf)
    Yv = np.column_stack([np.loadtxt(f) for f in yf])
    Zv = np.column_stack([np.loadtxt(f) for f in zf])
    assert rv.r_class == ("vertexLm", "vertexMultiDim", "matrix") and rv.shape == (V, 9)
    check_row(rv, 5, Yv[5], np.column_stack([np.ones(n), sexm, Zv[5]]), ["(Intercept)", "SexM", "curv"])


def test_mincLm_on_stored_volumes(api, study):
    """A reader that hands over the stored voxels (uint16 + per-slice real range, as MINC keeps them) instead of
    doubles: same result as the fit on the doubles the CPU reader would have produced."""
    gf, maskfile, _ = study
    rng = np.random.default_rng(2)
    store = {}
    for f in gf.jacobians:
        real = np.load(f)
        lo, hi = real.min(axis=(1, 2)) - 0.01, real.max(axis=(1, 2)) + 0.01          # per-slice image-min / image-max
        vox = np.round((real - lo[:, None, None]) / (hi - lo)[:, None, None] * 65535.0).astype(np.uint16)
        store[f] = api.StoredVolume(vox, image_min=lo, image_max=hi)

    def stored_reader(item):
        return store[item] if item in store else api.default_reader(item)

    vs = api.mincLm("jacobians ~ Sex + Scale", gf, mask=maskfile, reader=stored_reader)
    Yq = np.stack([store[f].real() for f in gf.jacobians], axis=1)                   # the quantised doubles
    X = np.column_stack([np.ones(10), gf.Sex == "M", gf.Scale.values]).astype(float)
    inside = np.load(maskfile).reshape(-1) > 0.5
    v = int(np.flatnonzero(inside)[17])
    check_row(vs, v, Yq[v], X, ["(Intercept)", "SexM", "Scale"])
    assert (np.asarray(vs)[~inside] == 0).all()
    # the permutation test of a stored-type fit runs on the stored voxels too, and equals the test on the doubles
    idx = api.draw_indices(10, 12, seed=5)
    rs = api.mincRandomize(vs, R=12, idx=idx)
    from rminc_b200 import core
    m = np.load(maskfile).reshape(-1).astype(float)
    tcols = [i for i, c in enumerate(vs.colnames) if c.startswith("tvalue-")]
    np.testing.assert_allclose(np.asarray(rs["randomization_dist"]), core.randomize(Yq, X, idx, tcols, mask=m), rtol=1e-9)
    assert api.thresholds(rs).shape == (4, 3)
    # mincTFCE of a stored-type fit: the permutation test sees the doubles the CPU reader would have produced
    tf = api.mincTFCE(vs, R=3, idx=idx[:, :3], d=0.2, like_volume=(10, 10, 10))
    want_tf = core.randomize_volume_tfce(Yq, X, idx[:, :3], tcols, (10, 10, 10), d=0.2, mask=m)
    np.testing.assert_allclose(np.asarray(tf["randomization_dist"]), want_tf, rtol=1e-9)
    np.testing.assert_array_equal(api._expand_stored(**vs.attrs["_stored"]), Yq)
    # float32 volumes need no range tables
    f32 = {f: api.StoredVolume(np.load(f).astype(np.float32)) for f in gf.jacobians}
    v32 = api.mincLm("jacobians ~ Sex", gf, reader=lambda item: f32[item])
    Y32 = np.stack([np.load(f).astype(np.float32).astype(np.float64).reshape(-1) for f in gf.jacobians], axis=1)
    check_row(v32, 3, Y32[3], X[:, :2], ["(Intercept)", "SexM"])
    with pytest.raises(TypeError):
        api.StoredVolume(np.zeros(4, dtype=np.int32))


def test_mincFDR(api, study):
    """mincFDR on a mincLm result (tests/testthat/test_mincFDR.R style): q-values of the F and t columns, 1 outside
    the mask, thresholds attribute 5 x ncols."""
    from scipy import stats as S
    gf, maskfile, Y = study
    vs = api.mincLm("jacobians ~ Sex", gf, mask=maskfile)
    qv = api.mincFDR(vs, mask=maskfile)
    assert qv.r_class == ("mincQvals", "mincMultiDim", "matrix")
    assert qv.colnames == ["qvalue-F-statistic", "qvalue-tvalue-(Intercept)", "qvalue-tvalue-SexM"]
    assert qv.attrs["thresholds"].shape == (5, 3) and qv.attrs["thresholds"].rownames == ["0.01", "0.05", "0.1", "0.15", "0.2"]
    assert api.thresholds(qv) is qv.attrs["thresholds"]                      # thresholds.mincQvals
    assert qv.attrs["DF"] == [[1, 8], 8, 8]
    inside = np.load(maskfile).reshape(-1) > 0.5
    assert (np.asarray(qv)[~inside] == 1).all()
    # independent BH on the SexM column
    t = vs.col("tvalue-SexM")[inside]
    p = 2 * S.t.sf(np.abs(t), 8)
    o = np.argsort(p)
    bh = np.minimum.accumulate((p[o] * len(p) / np.arange(1, len(p) + 1))[::-1])[::-1]
    want = np.empty_like(p)
    want[o] = np.minimum(bh, 1)
    assert np.asarray(qv)[inside, 2] == pytest.approx(want, rel=1e-9)
    # for one numerator degree of freedom F = t^2, so both columns carry the same q-values
    assert np.asarray(qv)[inside, 0] == pytest.approx(np.asarray(qv)[inside, 2], rel=1e-8)
    th = np.asarray(qv.attrs["thresholds"])[:, 2]
    acc = np.abs(t)[want <= 0.2]
    if acc.size:
        assert th[4] == pytest.approx(acc.min(), rel=1e-8)
    else:
        assert np.isnan(th[4])
    sub = api.mincFDR(vs, columns=["tvalue-SexM"], mask=maskfile)
    assert sub.shape == (1000, 1) and np.array_equal(np.asarray(sub)[:, 0], np.asarray(qv)[:, 2])
    with pytest.raises(ValueError, match="need to specify the degrees of freedom"):
        bad = api.MincMatrix(np.asarray(vs), colnames=vs.colnames, attrs={"stat-type": vs.attrs["stat-type"]})
        api.mincFDR(bad)
    # the other methods (R/minc_FDR.R:429-438): BY = c(m) * BH capped at 1; pFDR / qvalue estimate pi0 from the
    # mean(p >= lambda) grid through the smoothing spline and scale by it
    by = api.mincFDR(vs, columns=["tvalue-SexM"], mask=maskfile, method="BY")
    c = np.sum(1.0 / np.arange(1, inside.sum() + 1))
    assert np.asarray(by)[inside, 0] == pytest.approx(np.minimum(1.0, c * bh[np.argsort(o)]), rel=1e-9)
    lam = np.arange(1, 20) * 0.05
    pi0 = min(api.smooth_spline_at_last_knot(lam, np.array([np.mean(p >= l) for l in lam]) / (1 - lam), 3.0), 1.0)
    qq = api.mincFDR(vs, columns=["tvalue-SexM"], mask=maskfile, method="qvalue")
    assert np.asarray(qq)[inside, 0] == pytest.approx(pi0 * want, rel=1e-8)
    pf = api.mincFDR(vs, columns=["tvalue-SexM"], mask=maskfile, method="pFDR")
    assert pf.shape == (1000, 1) and (np.asarray(pf)[~inside] == 1).all() and (np.asarray(pf)[inside] <= 1).all()
    with pytest.raises(ValueError, match="unknown method"):
        api.mincFDR(vs, method="bonferroni")


def test_vertexTFCE(api, tmp_path, oracle):
    """vertexTFCE on a vector, a matrix and a vertexLm result (R/minc_vertex_statistics.R:731-948); the surface is
    given as the 0-based adjacency list the reference accepts."""
    from helpers import grid_adjacency, tfce_randomize_oracle
    rng = np.random.default_rng(17)
    a, b, n = 12, 15, 12
    V = a * b
    adj = grid_adjacency(a, b)
    ptr, aidx = oracle.adjacency_csr(adj)
    dx = np.array(["AD", "CN"] * 6)
    files = []
    for i in range(n):
        f = tmp_path / f"t{i}.txt"
        y = rng.normal(3, 0.3, V)
        y[40:70] += 0.6 * (dx[i] == "CN")
        np.savetxt(f, y)
        files.append(str(f))
    gf = pd.DataFrame(dict(thickness=files, Dx=dx))
    vs = api.vertexLm("thickness ~ Dx", gf)
    t = vs.col("tvalue-DxCN")
    one = api.vertexTFCE(t, adj)
    w_one = oracle.graph_tfce(t, ptr, aidx)
    assert one.shape == (V,) and np.abs(one - w_one).max() <= 1e-10 * np.abs(w_one).max()
    pos = api.vertexTFCE(t, adj, side="positive", E=1.0, H=1.0, nsteps=20)
    assert np.allclose(pos, oracle.graph_tfce(t, ptr, aidx, E=1.0, H=1.0, nsteps=20, side="positive"), rtol=1e-9, atol=1e-12)
    mat = api.vertexTFCE(np.column_stack([t, -t]), adj)
    assert mat.shape == (V, 2) and np.allclose(np.asarray(mat)[:, 1], -np.asarray(mat)[:, 0], rtol=1e-9, atol=1e-12)
    res = api.vertexTFCE(vs, adj, R=7, seed=4, nsteps=40)
    assert res.r_class == ("vertexTFCE_randomization", "vertex_randomization")
    assert res["tfce"].colnames == ["tvalue-(Intercept)", "tvalue-DxCN"] and res["tfce"].shape == (V, 2)
    assert res["randomization_dist"].shape == (7, 2) and res["args"]["side"] == "both"
    idx = api.draw_indices(n, 7, seed=4)
    want = tfce_randomize_oracle(oracle, vs.attrs["_Y"], vs.attrs["model"], idx, [4, 5], adj, nsteps=40)
    assert np.asarray(res["randomization_dist"]) == pytest.approx(want, rel=1e-8)
    th = api.thresholds(res)
    assert th.shape == (4, 2)
    # vertexTFCE.character: one file is a vector, several files the columns of a matrix
    assert np.allclose(api.vertexTFCE(files[0], adj), oracle.graph_tfce(np.loadtxt(files[0]), ptr, aidx), rtol=1e-9, atol=1e-12)
    assert api.vertexTFCE(files[:3], adj).shape == (V, 3)
    # vertexRandomize.vertexLm: mincRandomize's loop with the vertex classes
    vr = api.vertexRandomize(vs, R=6, seed=2)
    assert vr.r_class == ("vertexLm_randomization", "vertex_randomization") and vr["randomization_dist"].shape == (6, 2)
    assert np.allclose(np.asarray(vr["randomization_dist"]), np.asarray(api.mincRandomize(vs, R=6, seed=2)["randomization_dist"]))
    with pytest.raises(TypeError):
        api.vertexRandomize(np.zeros((3, 3)))
    with pytest.raises(ValueError, match="should be one of"):
        api.vertexTFCE(t, adj, side="sideways")
    # a bic_obj surface: default weights = mesh_area, adjacency = obj_to_graph's edge list (R/minc_vertex_statistics.R:751-781)
    from test_oracle import _icosphere
    vm, tm = _icosphere()
    obj = api.BicObj(vm, tm)
    xm = rng.standard_normal(vm.shape[1])
    ptr_o, idx_o = oracle.adjacency_csr(api.obj_to_adjacency(obj))
    want_o = oracle.graph_tfce(xm, ptr_o, idx_o, weights=oracle.mesh_area(vm, tm))
    assert np.abs(api.vertexTFCE(xm, obj) - want_o).max() <= 1e-10 * np.abs(want_o).max()
    want_1 = oracle.graph_tfce(xm, ptr_o, idx_o)
    assert np.abs(api.vertexTFCE(xm, obj, weights=np.ones(vm.shape[1])) - want_1).max() <= 1e-10 * np.abs(want_1).max()


def test_mincTFCE(api, study, oracle):
    """mincTFCE on a volume, a matrix and a mincLm result (R/minc_voxel_statistics.R:1186-1439); the shapes and classes
    of tests/testthat/test_mincTFCE.R:77-149, the values against the restatement."""
    gf, maskfile, _ = study
    dims = (10, 10, 10)
    vs = api.mincLm("jacobians ~ Sex", gf, mask=maskfile)
    t = vs.col("tvalue-SexM")
    one = api.mincTFCE(t, like_volume=dims, d=0.05)
    assert one.shape == (1000,) and np.allclose(one, oracle.volume_tfce(t, dims, d=0.05), rtol=1e-9, atol=1e-12)
    assert (one[np.load(maskfile).reshape(-1) < 0.5] == 0).all()
    # a 3-D array carries its own dimensions; voxel volume and the other sides
    vol3 = t.reshape(dims)
    pos = api.mincTFCE(vol3, d=0.1, E=1.0, H=1.0, side="positive", voxel_volume=0.001)
    assert np.allclose(pos, oracle.volume_tfce(t, dims, d=0.1, E=1.0, H=1.0, side="positive", voxel_volume=0.001),
                       rtol=1e-9, atol=1e-15)
    # matrix: every column, NaN -> 0, one d per column; third column of zeros stays zero (test_mincTFCE.R:96-118)
    tn = t.copy()
    tn[5] = np.nan
    mat = api.mincTFCE(np.column_stack([tn, t, np.zeros_like(t)]), like_volume=gf.jacobians[0], d=[0.05, 0.1, 0.05])
    t0 = np.where(np.isnan(tn), 0.0, tn)
    assert mat.shape == (1000, 3) and (np.asarray(mat)[:, 2] == 0).all()
    assert np.allclose(np.asarray(mat)[:, 0], oracle.volume_tfce(t0, dims, d=0.05), rtol=1e-9, atol=1e-12)
    assert np.allclose(np.asarray(mat)[:, 1], oracle.volume_tfce(t, dims, d=0.1), rtol=1e-9, atol=1e-12)
    # mincLm result: permutation test with the enhancement as post-processing
    res = api.mincTFCE(vs, R=5, seed=3, d=0.1)
    assert res.r_class == ("mincTFCE_randomization", "minc_randomization")
    assert res["tfce"].colnames == ["tvalue-(Intercept)", "tvalue-SexM"] and res["tfce"].shape == (1000, 2)
    assert res["randomization_dist"].shape == (5, 2) and res["args"]["side"] == "both"
    idx = api.draw_indices(10, 5, seed=3)
    want = oracle.randomize_volume_tfce(vs.attrs["_Y"], vs.attrs["model"], idx, [4, 5], dims, mask=vs.attrs["_mask"], d=0.1)
    assert np.asarray(res["randomization_dist"]) == pytest.approx(want, rel=1e-8)
    assert api.thresholds(res).shape == (4, 2)
    with pytest.raises(ValueError, match="should be one of"):
        api.mincTFCE(t, like_volume=dims, side="sideways")
    with pytest.raises(ValueError):
        api.mincTFCE(t, like_volume=(10, 10))


def test_model_selection_on_the_logLik_column(api, study):
    """AIC / BIC at a row against the closed-form lm values (tests/testthat/test_mincLm.R:54-55), AICc's correction and
    compare_models / its summary (test_mincLm.R:211-232)."""
    gf, maskfile, _ = study
    fits = [api.mincLm(f, gf) for f in ("jacobians ~ Sex", "jacobians ~ Scale", "jacobians ~ Sex + Scale", "jacobians ~ Coil")]
    vs = fits[0]
    n, v = 10, 7
    y = np.load(gf.jacobians[0]).reshape(-1)[v], None
    Y = np.array([np.load(f).reshape(-1)[v] for f in gf.jacobians])
    X = np.asarray(vs.attrs["model"])
    beta, rss = np.linalg.lstsq(X, Y, rcond=None)[:2]
    loglik = -0.5 * n * (np.log(2 * np.pi) + 1 - np.log(n) + np.log(rss[0]))
    p1 = X.shape[1] + 1
    assert api.AIC(vs)[v] == pytest.approx(-2 * loglik + 2 * p1, rel=1e-9)               # stats::AIC of the same lm
    assert api.BIC(vs)[v] == pytest.approx(-2 * loglik + np.log(n) * p1, rel=1e-9)       # stats::BIC
    assert api.AICc(vs)[v] == pytest.approx(api.AIC(vs)[v] + 2 * (p1 + 1) * (p1 + 2) / (n - p1 - 2), rel=1e-12)
    comp1 = api.compare_models(*fits, metric=api.AIC)
    comp2 = api.compare_models(*fits)
    params = np.array([np.asarray(m.attrs["model"]).shape[1] + 1 for m in fits])
    corr = 2 * (params + 1) * (params + 2) / (n - params - 2)
    assert comp2.r_class == ("model_comparison", "matrix") and comp2.shape == (1000, 4)
    np.testing.assert_allclose(np.asarray(comp2), np.asarray(comp1) + corr[None, :], rtol=1e-12)
    assert comp2.attrs["formulae"] == ["jacobians ~ Sex", "jacobians ~ Scale", "jacobians ~ Sex + Scale", "jacobians ~ Coil"]
    summ = api.summary_model_comparison(comp2)
    assert sum(r["wins"] for r in summ) >= 1000 and [r["wins"] for r in summ] == sorted(r["wins"] for r in summ)
    with pytest.raises(ValueError, match="more than one model"):
        api.compare_models(vs)
    with pytest.raises(TypeError):
        api.AIC(np.zeros((3, 3)))



def test_mincLm_on_minc2_files(api, study, tmp_path):
    """The whole front of the path on MINC 2 files (tests/testthat/test_mincLm.R runs on .mnc files): the default reader
    decodes the study in parallel through the library's own HDF5 / MINC 2 reader, the uint16 voxels and their per-slice
    ranges go to the stored-type fit, the mask is a uint8 label volume, and mincTFCE takes the dimensions from the
    likeVolume's header -- same numbers as the fit on the doubles libminc's reader would have produced."""
    import h5mini
    gf, maskfile, _ = study
    files, Yq = [], []
    for i, f in enumerate(gf.jacobians):
        real = np.load(f)
        lo, hi = real.min(axis=(1, 2)) - 0.01, real.max(axis=(1, 2)) + 0.01
        vox = np.round((real - lo[:, None, None]) / (hi - lo)[:, None, None] * 65535.0).astype(np.uint16)
        p = str(tmp_path / f"subj{i:02d}.mnc")
        h5mini.write_minc2(p, vox, image_min=lo, image_max=hi, valid_range=(0, 65535), chunks=(1, 10, 10), compress=4,
                           steps=(0.1, 0.1, 0.1))
        files.append(p)
        Yq.append(api.StoredVolume(vox, image_min=lo, image_max=hi).real())
    Yq = np.stack(Yq, axis=1)
    gm = gf.assign(jacobians=files)
    mpath = str(tmp_path / "mask.mnc")
    h5mini.write_minc2(mpath, np.load(maskfile).astype(np.uint8), valid_range=(0, 255), chunks=(10, 10, 10))
    vs = api.mincLm("jacobians ~ Sex + Scale", gm, mask=mpath)
    X = np.column_stack([np.ones(10), gf.Sex == "M", gf.Scale.values]).astype(float)
    inside = np.load(maskfile).reshape(-1) > 0.5
    for v in np.flatnonzero(inside)[[0, 17, 101]]:
        check_row(vs, int(v), Yq[v], X, ["(Intercept)", "SexM", "Scale"])
    assert (np.asarray(vs)[~inside] == 0).all() and vs.attrs["likeVolume"] == files[0]
    np.testing.assert_array_equal(api._expand_stored(**vs.attrs["_stored"]), Yq)
    # a right-hand-side volume and other stored types take the real-valued route through the same reader
    idx = api.draw_indices(10, 3, seed=5)
    tf = api.mincTFCE(vs, R=3, idx=idx, d=0.2)                   # dimensions from the MINC header of likeVolume
    assert np.asarray(tf["tfce"]).shape == (1000, 3) and np.asarray(tf["randomization_dist"]).shape == (3, 3)
    i16 = []
    for i, f in enumerate(gf.jacobians):
        p = str(tmp_path / f"short{i:02d}.mnc")
        real = np.load(f)
        vox = np.round((real - real.min()) / (real.max() - real.min()) * 65535.0 - 32768.0).astype(np.int16)
        h5mini.write_minc2(p, vox, image_min=real.min(), image_max=real.max(), valid_range=(-32768, 32767), chunks=(2, 10, 10))
        i16.append(p)
    v16 = api.mincLm("jacobians ~ Sex", gf.assign(jacobians=i16))
    Y16 = np.stack([api.default_reader(p) for p in i16], axis=1)
    assert "_Y" in v16.attrs and np.array_equal(v16.attrs["_Y"], Y16)
    check_row(v16, 7, Y16[7], X[:, :2], ["(Intercept)", "SexM"])
    with pytest.raises(FileNotFoundError, match="Error opening input file"):
        api.mincLm("jacobians ~ Sex", gm.assign(jacobians=[str(tmp_path / "gone.mnc")] * 10))
    # the other callers of the path read the same files: mincAnova (real values of the uint16 volumes), mincRandomize and
    # mincFDR on the fit
    an = api.mincAnova("jacobians ~ Sex + Scale", gm, mask=mpath)
    want_an = api.mincAnova("jacobians ~ Sex + Scale", gf.assign(jacobians=[Yq[:, j].reshape(10, 10, 10) for j in range(10)]),
                            mask=np.load(maskfile))
    np.testing.assert_allclose(np.asarray(an), np.asarray(want_an), rtol=1e-9, atol=0)
    rs = api.mincRandomize(vs, R=6, idx=api.draw_indices(10, 6, seed=3))
    assert np.asarray(rs["randomization_dist"]).shape == (6, 3)
    fd = api.mincFDR(vs, mask=mpath)
    assert np.asarray(fd).shape[0] == 1000
