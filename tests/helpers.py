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
    """Result matrices agree: identical NaN / +-Inf pattern, and every finite entry within
    rtol * max(|want|, column RMS).

    The column-RMS floor matters only for entries that are numerically ~0 relative to their column
    (a beta of 1e-7 in a column of O(1) betas): such an entry is a difference of O(column scale)
    quantities, so both the reference and the kernel carry an absolute error of eps * scale and a
    relative comparison on the entry itself would test rounding noise, not parity."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    if got.ndim == 1:
        got, want = got[:, None], want[:, None]
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    assert np.array_equal(nan_g, nan_w), f"{what}: NaN pattern differs at {np.argwhere(nan_g != nan_w)[:5]}"
    inf_g, inf_w = np.isinf(got), np.isinf(want)
    assert np.array_equal(inf_g, inf_w), f"{what}: Inf pattern differs"
    assert np.array_equal(np.sign(got[inf_g]), np.sign(want[inf_w])), f"{what}: Inf sign differs"
    fin = np.isfinite(want)
    wz = np.where(fin, want, 0.0)
    cnt = np.maximum(fin.sum(axis=0), 1)
    rms = np.sqrt((wz ** 2).sum(axis=0) / cnt)
    scale = np.maximum(np.abs(wz), np.broadcast_to(rms, want.shape))
    err = np.where(fin, np.abs(np.where(fin, got, 0.0) - wz), 0.0)
    bad = err > rtol * scale
    assert not bad.any(), (f"{what}: {bad.sum()} entries differ; worst err/scale "
                           f"{(err[bad] / np.maximum(scale[bad], 1e-300)).max():.3e} at {np.argwhere(bad)[:3].tolist()}")


def grid_adjacency(a, b):
    """0-based adjacency list of an a x b grid graph (4-neighbourhood), the list form vertexTFCE accepts."""
    adj = []
    for i in range(a):
        for j in range(b):
            nb = []
            if i > 0:
                nb.append((i - 1) * b + j)
            if i < a - 1:
                nb.append((i + 1) * b + j)
            if j > 0:
                nb.append(i * b + j - 1)
            if j < b - 1:
                nb.append(i * b + j + 1)
            adj.append(nb)
    return adj


def tfce_randomize_oracle(oracle, Y, X, idx, cols, adjacency, two_sided=True, E=0.5, H=2.0, nsteps=100, side="both",
                          weights=None):
    """mincRandomize_core with post_proc = vertexTFCE.matrix, literally (R/minc_voxel_statistics.R:1465-1497):
    refit with permuted design rows, non-finite -> 0, enhance every column, abs, column maxima."""
    ptr, aidx = oracle.adjacency_csr(adjacency)
    R = idx.shape[1]
    dist = np.empty((R, len(cols)))
    for r in range(R):
        res = oracle.vertex_lm_loop(Y, X[idx[:, r]])
        res[~np.isfinite(res)] = 0.0
        for j, c in enumerate(cols):
            tf = oracle.graph_tfce(res[:, c], ptr, aidx, E=E, H=H, nsteps=nsteps, side=side, weights=weights)
            dist[r, j] = np.max(np.abs(tf) if two_sided else tf)
    return dist
