"""The resident dataset (rmg_dataset_*) and the single-process multi-GPU forms (rmg_*_multi), through the C ABI.

Reference behaviour: mincLm -> mincRandomize -> mincAnova -> mincFDR on the same files
(R/minc_voxel_statistics.R:911-912, :1457-1477) and the parallel= sharding (:848-906).  One upload, every routine on
the resident shards; results equal the one-call forms and the CPU oracle.
"""
import numpy as np
import pytest

from helpers import assert_rows_match, cfg1_data, design_cfg2, ellipsoid_mask

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.fixture(scope="module")
def core(lib):
    from rminc_b200 import core as c
    assert c.device_count() > 0, "GPU tests need a CUDA device"
    return c


def perm_indices(n, R, seed=4, replace=False):
    rng = np.random.default_rng(seed)
    if replace:
        return np.column_stack([rng.integers(0, n, n) for _ in range(R)])
    return np.column_stack([rng.permutation(n) for _ in range(R)])


def gpu_counts(core):
    return sorted({1, min(2, core.device_count()), core.device_count()})


def test_fit_randomize_anova_fdr_on_one_upload(core, oracle):
    """mincLm, mincRandomize (twice: the second reuses the resident per-voxel sums), mincAnova and mincFDR on ONE handle."""
    Y, X, mask = cfg1_data(dims=(16, 20, 13))
    V, n = Y.shape
    idx = perm_indices(n, 37)
    want_fit = oracle.minc2_model_lm(Y, (16, 20, 13), X, mask=mask)
    want_perm = oracle.randomize(Y, X, idx, [4, 5], two_sided=True, mask=mask)
    for G in gpu_counts(core):
        with core.Dataset(Y, mask=mask, n_gpus=G) as ds:
            assert ds.n_gpus == G and ds.V == V and ds.n == n
            got = ds.lm_fit(X)
            assert_rows_match(got, want_fit, RTOL, f"dataset fit G={G}")
            assert (got[mask < 0.5] == 0).all()
            np.testing.assert_allclose(ds.randomize(X, idx, [4, 5]), want_perm, rtol=RTOL, atol=0)
            np.testing.assert_allclose(ds.randomize(X, idx[:, :9], [5], two_sided=False),
                                       oracle.randomize(Y, X, idx[:, :9], [5], two_sided=False, mask=mask), rtol=RTOL, atol=0)
            # a design without the constant in its span: un-shifted sums (a second resident set on the same handle)
            X0 = np.column_stack([X[:, 1], np.linspace(-1, 1, n)])
            np.testing.assert_allclose(ds.randomize(X0, idx[:, :11], [4, 5]),
                                       oracle.randomize(Y, X0, idx[:, :11], [4, 5], mask=mask), rtol=RTOL, atol=0)
            # non-t columns take the refit path on the resident shards
            np.testing.assert_allclose(ds.randomize(X, idx[:, :7], [0, 1, 3, 6]),
                                       oracle.randomize(Y, X, idx[:, :7], [0, 1, 3, 6], mask=mask), rtol=RTOL, atol=0)
            # FDR on the resident t column of the LAST fit (fit again: the anova below replaces the resident result)
            ds.lm_fit(X, want_result=False)
            q, th = ds.fdr(5, core.STAT_T, n - 2)
            wq, wth, _ = oracle.fdr(want_fit[:, 5], "t", n - 2, mask=mask)
            np.testing.assert_allclose(q, wq, rtol=RTOL, atol=0)
            assert np.array_equal(np.isnan(th), np.isnan(wth))
            np.testing.assert_allclose(th[~np.isnan(wth)], wth[~np.isnan(wth)], rtol=1e-8)
            qf, _ = ds.fdr(0, core.STAT_F, 1, n - 2)
            np.testing.assert_allclose(qf, oracle.fdr(want_fit[:, 0], "F", 1, n - 2, mask=mask)[0], rtol=RTOL, atol=0)
            # the other methods of mincFDR and fast.qvalue's pi0 grid on the resident column (R/minc_FDR.R:429-438)
            fr = ds.fdr_pi0_fractions(5, core.STAT_T, n - 2)
            np.testing.assert_allclose(fr, core.fdr_pi0_fractions(got[:, 5], core.STAT_T, n - 2, mask=mask), rtol=1e-12, atol=0)
            for method, pi0 in (("BY", 1.0), ("pFDR", 0.8), ("qvalue", 0.8)):
                qm, thm = ds.fdr(5, core.STAT_T, n - 2, method=method, pi0=pi0)
                wm, wthm, _ = oracle.fdr(want_fit[:, 5], "t", n - 2, mask=mask, method=method, pi0=pi0)
                np.testing.assert_allclose(qm, wm, rtol=RTOL, atol=0, err_msg=method)
                assert np.array_equal(np.isnan(thm), np.isnan(wthm)), method
            a = ds.anova(X, [0, 1])
            assert_rows_match(a, core.anova_fit(Y, X, [0, 1], mask=mask), RTOL, f"dataset anova G={G}")
            with pytest.raises(core.RmincGpuError, match="out of range"):
                ds.fdr(3, core.STAT_F, 1, n - 2)
        with pytest.raises(core.RmincGpuError, match="closed"):
            ds.lm_fit(X)


def test_dataset_no_mask_odd_voxel_counts_and_blocks(core, oracle, monkeypatch):
    """Odd V (the last voxel of the last shard takes the explicit-residual kernel), several voxel blocks per shard,
    pageable input through the staging ring, p = 4 with a large mean."""
    rng = np.random.default_rng(21)
    V, n = 5001, 48
    X = design_cfg2(n)
    Y = np.asfortranarray(3000 + 40 * rng.standard_normal((V, n)))
    Y[::17] = 0.0
    idx = perm_indices(n, 41)
    cols = [6, 7, 8, 9]
    want = oracle.randomize(Y, X, idx, cols)
    want_fit = oracle.vertex_lm_loop(Y, X)
    for chunk, rows in ((None, None), ("1024", None), ("700", "5")):
        for k, v in (("RMG_PERM_CHUNK_VOXELS", chunk), ("RMG_STAGING_ROWS", rows)):
            if v:
                monkeypatch.setenv(k, v)
            else:
                monkeypatch.delenv(k, raising=False)
        for G in gpu_counts(core):
            monkeypatch.setenv("RMG_PERM_PATH", "gemm")
            with core.Dataset(Y, n_gpus=G) as ds:
                np.testing.assert_allclose(ds.randomize(X, idx, cols), want, rtol=RTOL, atol=0, err_msg=f"G={G} chunk={chunk}")
                assert_rows_match(ds.lm_fit(X), want_fit, RTOL, f"fit G={G} chunk={chunk}")
                one = ds.randomize(X, idx, cols, two_sided=False)
            monkeypatch.setenv("RMG_PERM_PATH", "refit")
            np.testing.assert_allclose(one, core.randomize(Y, X, idx, cols, two_sided=False), rtol=1e-10, atol=0)
    monkeypatch.delenv("RMG_PERM_PATH")


def test_dataset_from_stored_volumes(core, oracle):
    rng = np.random.default_rng(8)
    V, n, slice_len = 6002, 36, 700
    U = np.asfortranarray(np.round(rng.normal(30000, 400, (V, n))).astype(np.uint16))
    nsl = (V + slice_len - 1) // slice_len
    scale = rng.uniform(1.5, 2.5, (nsl, n)) / 65535.0
    offset = rng.normal(-0.8, 0.05, (nsl, n))
    X = design_cfg2(n)
    Yx = oracle.expand_stored(U, scale, offset, voxel_min=1.0, slice_len=slice_len)
    mask = (np.arange(V) % 4 > 0).astype(float)
    idx = perm_indices(n, 19)
    cols = [6, 7, 8, 9]
    for G in gpu_counts(core):
        with core.Dataset(U, mask=mask, n_gpus=G, scale=scale, offset=offset, voxel_min=1.0, slice_len=slice_len) as ds:
            assert_rows_match(ds.lm_fit(X), oracle.minc2_model_lm(Yx, (1, 1, V), X, mask=mask), RTOL, f"stored dataset G={G}")
            np.testing.assert_allclose(ds.randomize(X, idx, cols), oracle.randomize(Yx, X, idx, cols, mask=mask), rtol=RTOL, atol=0)
    F = np.asfortranarray(rng.normal(0, 0.2, (3000, n)).astype(np.float32))
    with core.Dataset(F) as ds:
        assert_rows_match(ds.lm_fit(X), oracle.vertex_lm_loop(F.astype(np.float64), X), RTOL, "float32 dataset")


def test_dataset_from_minc2_files(core, oracle, tmp_path):
    """Files -> resident dataset -> fit, permutation test, anova on one decode and one upload; the mask is a MINC 2 label
    volume.  Equals the oracle on the doubles libminc's reader would have produced."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import h5mini
    from rminc_b200 import api
    rng = np.random.default_rng(12)
    dims, n = (9, 14, 12), 24
    V = int(np.prod(dims))
    files, cols_real = [], []
    for j in range(n):
        vox = np.round(rng.normal(30000, 400, dims)).astype(np.uint16)
        lo = rng.normal(-0.8, 0.05, dims[0])
        hi = lo + rng.uniform(1.5, 2.5, dims[0])
        f = str(tmp_path / f"s{j:02d}.mnc")
        h5mini.write_minc2(f, vox, image_min=lo, image_max=hi, valid_range=(0, 65535), chunks=(1, 14, 12))
        files.append(f)
        cols_real.append(api.StoredVolume(vox, image_min=lo, image_max=hi).real())
    Yx = np.asfortranarray(np.stack(cols_real, axis=1))
    mask = ellipsoid_mask(dims)
    mfile = str(tmp_path / "mask.mnc")
    h5mini.write_minc2(mfile, mask.reshape(dims).astype(np.uint8), valid_range=(0, 255))
    X = design_cfg2(n)
    idx = perm_indices(n, 11)
    for G in gpu_counts(core):
        with core.Dataset.from_minc2(files, mask=mfile, n_gpus=G) as ds:
            assert (ds.V, ds.n) == (V, n)
            assert_rows_match(ds.lm_fit(X), oracle.minc2_model_lm(Yx, (1, 1, V), X, mask=mask), RTOL, f"minc2 dataset G={G}")
            np.testing.assert_allclose(ds.randomize(X, idx, [6, 7, 8, 9]), oracle.randomize(Yx, X, idx, [6, 7, 8, 9], mask=mask),
                                       rtol=RTOL, atol=0)


def test_dataset_error_contract(core):
    Y, X, mask = cfg1_data(dims=(8, 8, 8))
    n = X.shape[0]
    with core.Dataset(Y, mask=mask, n_gpus=1) as ds:
        Xs = np.column_stack([X, np.zeros(n)])                   # a zero column: R[3,3] == 0 exactly, dpotri info = 3
        with pytest.raises(core.SingularDesignError, match=r"element \(3, 3\) is zero"):
            ds.lm_fit(Xs)
        bad = perm_indices(n, 6, replace=True)
        bad[:, 2] = 0                                            # a resample with one level only
        with pytest.raises(core.SingularDesignError, match="permutation 3"):
            ds.randomize(X, bad, [4, 5])
        with pytest.raises(core.RmincGpuError, match="fit first"):
            ds.fdr(0, core.STAT_F, 1, n - 2)
        with pytest.raises(ValueError):
            ds.lm_fit(X[:-1])
    with pytest.raises(core.RmincGpuError, match="GPUs requested"):
        core.Dataset(Y, n_gpus=core.device_count() + 1)


# ------------------------------------------------------------------ the single-process multi-GPU forms
def _visible_gpus():
    try:
        from rminc_b200 import _lib
        return _lib.load().rmg_device_count()
    except Exception:  # noqa: BLE001  (library not built yet: the fixture builds it; this only decides the skip)
        return 0


need2 = pytest.mark.skipif(_visible_gpus() < 2, reason="needs two visible GPUs")


@need2
def test_every_multi_entry_point_on_two_gpus(core, oracle):
    """rmg_lm_fit_multi, rmg_lm_fit_cov_multi, rmg_lm_fit_stored_multi, rmg_randomize_multi, rmg_randomize_stored_multi,
    rmg_randomize_tfce_multi and rmg_randomize_volume_tfce(n_gpus = 2): each equals its one-GPU form and the oracle."""
    from helpers import grid_adjacency
    rng = np.random.default_rng(33)
    G = core.device_count()
    Y, X, mask = cfg1_data(dims=(17, 21, 13))                     # V = 4641 (odd: uneven shards)
    V, n = Y.shape
    want = oracle.minc2_model_lm(Y, (17, 21, 13), X, mask=mask)
    for g in sorted({2, G}):
        got = core.lm_fit(Y, X, mask=mask, n_gpus=g)
        assert_rows_match(got, want, RTOL, f"lm_fit_multi {g}")
        assert (got[mask < 0.5] == 0).all()
        idx = perm_indices(n, 23)
        np.testing.assert_allclose(core.randomize(Y, X, idx, [4, 5], mask=mask, n_gpus=g),
                                   oracle.randomize(Y, X, idx, [4, 5], mask=mask), rtol=RTOL, atol=0)
        np.testing.assert_allclose(core.randomize(Y, X, idx, [0, 3], mask=mask, n_gpus=g, two_sided=False),
                                   oracle.randomize(Y, X, idx, [0, 3], mask=mask, two_sided=False), rtol=RTOL, atol=0)
    # covariate and stored-type fits
    Z = np.asfortranarray(50 + 4 * rng.standard_normal((V, n)))
    assert_rows_match(core.lm_fit_cov(Y, Z, X, mask=mask, n_gpus=2), oracle.lm_loop_cov(Y, Z, X, mask=mask), RTOL, "cov multi")
    U = np.asfortranarray(np.round(rng.normal(30000, 400, (V, n))).astype(np.uint16))
    sl = 17 * 21
    nsl = (V + sl - 1) // sl
    scale, offset = rng.uniform(1.5, 2.5, (nsl, n)) / 65535.0, rng.normal(-0.8, 0.05, (nsl, n))
    Yx = oracle.expand_stored(U, scale, offset, slice_len=sl)
    assert_rows_match(core.lm_fit_stored(U, X, scale, offset, slice_len=sl, n_gpus=2), oracle.vertex_lm_loop(Yx, X), RTOL, "stored multi")
    idx = perm_indices(n, 13)
    np.testing.assert_allclose(core.randomize_stored(U, X, idx, [4, 5], scale=scale, offset=offset, slice_len=sl, n_gpus=2),
                               oracle.randomize(Yx, X, idx, [4, 5]), rtol=RTOL, atol=0)
    # TFCE permutation tests shard over permutations
    adj = grid_adjacency(13, 17)
    Yg = np.asfortranarray(Y[:13 * 17])
    a = core.randomize_tfce(Yg, X, idx, [4, 5], adj, nsteps=20)
    b = core.randomize_tfce(Yg, X, idx, [4, 5], adj, nsteps=20, n_gpus=2)
    np.testing.assert_allclose(b, a, rtol=1e-12, atol=0)
    dims = (6, 7, 5)
    Yv = np.asfortranarray(Y[:210])
    a = core.randomize_volume_tfce(Yv, X, idx, [5], dims, d=0.3)
    b = core.randomize_volume_tfce(Yv, X, idx, [5], dims, d=0.3, n_gpus=2)
    np.testing.assert_allclose(b, a, rtol=1e-12, atol=0)
