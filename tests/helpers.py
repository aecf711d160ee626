"""Synthetic workloads shared by the tests and comparison helpers."""
import numpy as np


def design_cfg1(n=20):
    g = np.r_[np.zeros(n // 2), np.ones(n - n // 2)]
    return np.column_stack([np.ones(n), g])


def design_cfg2(n=500, seed=1):
    rng = np.random.default_rng(seed)
    group = (np.arange(n) % 2).astype(float)
    age = np.round(rng.normal(60, 10, n), 1)
    sex = rng.integers(0, 2, n).astype(float)
    return np.column_stack([np.ones(n), group, age, sex])


def design_cov(n, p, seed=3):
    rng = np.random.default_rng(seed)
    return np.column_stack([np.ones(n)] + [rng.normal(size=n) for _ in range(p - 1)])


def ellipsoid_mask(dims, frac=0.45):
    z, y, x = np.meshgrid(*[np.arange(d) for d in dims], indexing="ij")
    c = [(d - 1) / 2 for d in dims]
    r = [frac * d for d in dims]
    inside = ((z - c[0]) / r[0]) ** 2 + ((y - c[1]) / r[1]) ** 2 + ((x - c[2]) / r[2]) ** 2 <= 1
    return inside.astype(np.float64).reshape(-1)


def cfg1_data(dims=(60, 80, 50), n=20, seed=20261019):
    """Config 1: y ~ group, 2-level group, masked; Y[v,j] = 0.2 z + 0.15 g_j s_v."""
    rng = np.random.default_rng(seed)
    V = int(np.prod(dims))
    X = design_cfg1(n)
    s = rng.random(V)
    Y = np.asfortranarray(0.2 * rng.standard_normal((V, n)) + 0.15 * np.outer(s, X[:, 1]))
    return Y, X, ellipsoid_mask(dims)


def assert_rows_match(got, want, rtol=1e-9, what=""):
    """Result matrices agree: same NaN/Inf pattern, finite entries within rtol (relative to the
    column scale for entries that are numerically zero), exact zeros where the oracle has zeros."""
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, (got.shape, want.shape)
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    assert np.array_equal(nan_g, nan_w), f"{what}: NaN pattern differs at {np.argwhere(nan_g != nan_w)[:5]}"
    inf_g, inf_w = np.isinf(got), np.isinf(want)
    assert np.array_equal(inf_g, inf_w), f"{what}: Inf pattern differs"
    assert np.array_equal(np.sign(got[inf_g]), np.sign(want[inf_w])), f"{what}: Inf sign differs"
    fin = np.isfinite(want)
    g, w = got[fin], want[fin]
    err = np.abs(g - w)
    scale = np.abs(w)
    bad = err > rtol * scale + 1e-300
    if bad.any():
        # entries that are ~0 relative to their column (e.g. a beta of 1e-17) are compared on the column scale
        colscale = np.broadcast_to(np.nanmax(np.where(np.isfinite(want), np.abs(want), 0), axis=0), want.shape)[fin]
        bad &= err > rtol * colscale * 1e-3
    assert not bad.any(), (f"{what}: {bad.sum()} entries differ; worst rel err "
                           f"{(err[bad] / np.maximum(scale[bad], 1e-300)).max():.3e}")
