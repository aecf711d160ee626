"""BASELINE.json full sizes on the GPU: the oracle cannot run 7 M voxels, so parity is checked
(a) against the oracle on a random sample of voxels drawn from the full-size problem and
(b) through size-independent properties of OLS over ALL voxels: affine equivariance, linearity
of beta in Y, invariance under a joint permutation of subjects, masked rows bit-exact zero."""
import numpy as np
import pytest

from helpers import assert_rows_match, design_cfg2, design_cov, ellipsoid_mask

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.fixture(scope="module")
def torch_mod(lib):
    import torch
    from rminc_b200 import core
    assert core.device_count() > 0
    free, _ = torch.cuda.mem_get_info()
    return torch, core, free


def gen(torch, V, n, seed, mean=0.0, sd=0.2):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    Yt = torch.empty((n, V), dtype=torch.float64, device="cuda")
    for j0 in range(0, n, 25):
        Yt[j0:j0 + 25].normal_(mean, sd, generator=g)
    return Yt


def sample_check(torch, core, oracle, Yt, X, out, mask=None, k=1500, seed=0, what=""):
    V = Yt.shape[1]
    rng = np.random.default_rng(seed)
    sel = np.sort(rng.choice(V, size=k, replace=False))
    sel[0], sel[-1] = 0, V - 1                        # first and last row (int64 indexing at the far end)
    st = torch.from_numpy(sel).cuda()
    Ys = Yt[:, st].cpu().numpy().T                    # k x n
    want = oracle.vertex_lm_loop(Ys, X)
    got = out[:, st].cpu().numpy().T
    if mask is not None:
        m = mask[sel] > 0.5
        assert (got[~m] == 0).all() and not np.signbit(got[~m]).any(), what
        got, want = got[m], want[m]
    assert_rows_match(got, want, RTOL, what)


def test_config2_full_size(torch_mod, oracle):
    """mincLm(y ~ group + age + sex), 500 x 180x252x156: V = 7,076,160, n = 500, p = 4."""
    torch, core, free = torch_mod
    dims, n = (180, 252, 156), 500
    V = int(np.prod(dims))
    if free < 2.3 * V * n * 8:
        pytest.skip("not enough free HBM for two copies of config 2")
    X = design_cfg2(n)
    p = X.shape[1]
    Yt = gen(torch, V, n, seed=1, mean=30000.0, sd=50.0)     # ushort-like intensities (hard part H1)
    out = core.lm_fit_torch(Yt, X)
    torch.cuda.synchronize()
    sample_check(torch, core, oracle, Yt, X, out, what="cfg2 sample")
    # masked: rows outside the ellipsoid are +0.0 bit-exact, rows inside unchanged bit-for-bit
    mask = ellipsoid_mask(dims)
    mt = torch.from_numpy(mask).cuda()
    outm = core.lm_fit_torch(Yt, X, mask=mt)
    torch.cuda.synchronize()
    inside = mt > 0.5
    assert 0.3 < float(inside.double().mean()) < 0.45
    assert bool((outm[:, ~inside] == 0).all()) and not bool(torch.signbit(outm[:, ~inside]).any())
    assert torch.equal(outm[:, inside], out[:, inside])
    del outm, mt
    # affine equivariance over ALL voxels: y -> a*y + b
    a, b = -2.5, 1234.5
    Y2 = Yt * a + b
    out2 = core.lm_fit_torch(Y2, X)
    torch.cuda.synchronize()
    del Y2

    def close(u, v, tol=1e-8):
        scale = torch.maximum(v.abs(), v.square().mean().sqrt())
        return bool(((u - v).abs() <= tol * scale).all())

    assert close(out2[0], out[0]) and close(out2[1], out[1])                       # F, R2 invariant
    assert close(out2[2], a * out[2] + b)                                          # intercept
    for i in range(1, p):
        assert close(out2[2 + i], a * out[2 + i])                                  # slopes scale
        assert close(out2[2 + p + i], np.sign(a) * out[2 + p + i])                 # t flips sign
    assert close(out2[2 * p + 2], out[2 * p + 2] - n * np.log(abs(a)))             # logLik shift
    del out2
    # joint permutation of subjects (rows of X and of Y): same model, same numbers
    perm = np.random.default_rng(3).permutation(n)
    Yp = Yt[torch.from_numpy(perm).cuda()]
    outp = core.lm_fit_torch(Yp, X[perm])
    torch.cuda.synchronize()
    for c in range(2 * p + 3):
        assert close(outp[c], out[c], 1e-9), c


def test_config3_vertex_shape(torch_mod, oracle):
    """vertexLm: 81,924 vertices x 2000 subjects, 6 covariates (p = 7): small V, large n (split-n)."""
    torch, core, _ = torch_mod
    V, n = 81924, 2000
    X = design_cov(n, 7)
    Yt = gen(torch, V, n, seed=3, mean=2.5, sd=0.4)
    out = core.lm_fit_torch(Yt, X)
    torch.cuda.synchronize()
    sample_check(torch, core, oracle, Yt, X, out, k=600, what="cfg3 sample")
    # linearity of beta in Y over all vertices
    Y2 = gen(torch, V, n, seed=4, mean=-1.0, sd=0.3)
    o2 = core.lm_fit_torch(Y2, X)
    o12 = core.lm_fit_torch(Yt + Y2, X)
    torch.cuda.synchronize()
    b, b2, b12 = out[2:9], o2[2:9], o12[2:9]
    scale = b12.abs().mean(dim=1, keepdim=True) + b12.abs()
    assert bool(((b + b2 - b12).abs() <= 1e-9 * scale).all())


def test_config5_one_shard(torch_mod, oracle):
    """40 um volume 400x560x340 x 200 subjects over 8 GPUs: one GPU's shard is 9,520,000 voxels
    (15.2 GB); the full problem has 76,160,000 rows -> 64-bit indexing end to end."""
    torch, core, free = torch_mod
    V, n = 9_520_000, 200
    if free < 1.3 * V * n * 8:
        pytest.skip("not enough free HBM")
    from rminc_b200 import parallel
    assert int(parallel.shard_bounds(400 * 560 * 340, 8)[1]) == V
    X = design_cfg2(n)
    Yt = gen(torch, V, n, seed=5)
    mask = (torch.arange(V, device="cuda") % 3 != 0).double()
    out = core.lm_fit_torch(Yt, X, mask=mask)
    torch.cuda.synchronize()
    sample_check(torch, core, oracle, Yt, X, out, mask=mask.cpu().numpy(), k=1000, what="cfg5 shard sample")


def test_config4_randomize_full_size(torch_mod, oracle, monkeypatch):
    """mincRandomize on config 2 data: GEMM path equals the refit path on the full volume, and
    equals the oracle's literal refits on a voxel subset."""
    torch, core, free = torch_mod
    dims, n = (180, 252, 156), 500
    V = int(np.prod(dims))
    if free < 1.4 * V * n * 8:
        pytest.skip("not enough free HBM")
    X = design_cfg2(n)
    Yt = gen(torch, V, n, seed=1)
    mt = torch.from_numpy(ellipsoid_mask(dims)).cuda()
    rng = np.random.default_rng(4)
    idx = np.column_stack([rng.permutation(n) for _ in range(24)])
    cols = [6, 7, 8, 9]
    monkeypatch.setenv("RMG_PERM_PATH", "gemm")
    dg = core.randomize_torch(Yt, X, idx, cols, two_sided=True, mask=mt).cpu().numpy()
    monkeypatch.setenv("RMG_PERM_PATH", "refit")
    dr = core.randomize_torch(Yt, X, idx[:, :6], cols, two_sided=True, mask=mt).cpu().numpy()
    np.testing.assert_allclose(dg[:, :6], dr, rtol=1e-10)
    assert (dg > 4.0).all() and (dg < 9.0).all()            # max |t| over 2.7 M null voxels, 496 df
    # oracle on a contiguous slab that intersects the mask
    a = V // 2
    sl = slice(a, a + 3000)
    Ys = Yt[:, sl].cpu().numpy().T
    ms = mt[sl].cpu().numpy()
    want = oracle.randomize(Ys, X, idx[:, :4], cols, two_sided=True, mask=ms)
    monkeypatch.setenv("RMG_PERM_PATH", "gemm")
    got = core.randomize_torch(Yt[:, sl], X, idx[:, :4], cols, two_sided=True, mask=mt[sl].contiguous()).cpu().numpy().T
    np.testing.assert_allclose(got, want, rtol=RTOL)


def test_config2_full_size_stored_and_covariate(torch_mod, oracle):
    """Config 2 through the next-row kernels at full size: uint16 voxels with per-slice real ranges (TMA ring) and a
    voxel-wise covariate.  Checked against the oracle on a voxel sample, and over ALL voxels through properties: the
    stored-type fit equals the double fit of the expanded samples; a covariate that does not enter y leaves the
    static coefficients' estimates those of a model that projects it out (Frisch-Waugh: beta_z of y := y + c z moves
    by exactly c)."""
    torch, core, free = torch_mod
    dims, n = (180, 252, 156), 500
    V = int(np.prod(dims))
    if free < 2.4 * V * n * 8:
        pytest.skip("not enough free HBM")
    X = design_cfg2(n)
    p = X.shape[1]
    nsl, slice_len = dims[0], dims[1] * dims[2]
    g = torch.Generator(device="cuda")
    g.manual_seed(11)
    Ut = torch.empty((n, V), dtype=torch.uint16, device="cuda")
    for j0 in range(0, n, 25):
        blk = torch.empty((25, V), dtype=torch.float32, device="cuda").normal_(30000.0, 400.0, generator=g)
        Ut[j0:j0 + 25] = blk.round_().clamp_(0, 32767).to(torch.int16).view(torch.uint16)
    del blk
    rs = np.random.default_rng(5)
    smin = rs.normal(-0.8, 0.05, (n, nsl))
    smax = smin + rs.uniform(1.5, 2.5, (n, nsl))
    sc_h, of_h = np.ascontiguousarray((smax - smin) / 65535.0), np.ascontiguousarray(smin)
    sc_t, of_t = torch.from_numpy(sc_h).cuda(), torch.from_numpy(of_h).cuda()
    out = core.lm_fit_stored_torch(Ut, X, sc_t, of_t, slice_len=slice_len)
    torch.cuda.synchronize()
    assert core.last_fit_variant() == "lm_fit_packed_tma_kernel"
    # (a) oracle on a sample of voxels (expanded by the oracle's own restatement of the conversion)
    rng = np.random.default_rng(0)
    sel = np.sort(rng.choice(V, size=1200, replace=False))
    sel[0], sel[-1] = 0, V - 1
    sel[1:40] = np.arange(slice_len - 20, slice_len + 19)                 # across the first slice boundary
    sel = np.unique(sel)
    st = torch.from_numpy(sel).cuda()
    Us = Ut.view(torch.int16)[:, st].cpu().numpy().view(np.uint16).T      # k x n
    sl = sel // slice_len
    Ys = (Us.astype(np.float64) - 0.0) * sc_h.T[sl, :] + of_h.T[sl, :]
    assert_rows_match(out[:, st].cpu().numpy().T, oracle.vertex_lm_loop(Ys, X), RTOL, "stored sample")
    # (b) ALL voxels: equal to the double kernel on the expanded samples, slab by slab (a slab = 12 slices)
    for s0 in range(0, nsl, 12):
        v0, v1 = s0 * slice_len, min(V, (s0 + 12) * slice_len)
        scv = sc_t[:, s0:s0 + 12].repeat_interleave(slice_len, dim=1)[:, :v1 - v0]
        ofv = of_t[:, s0:s0 + 12].repeat_interleave(slice_len, dim=1)[:, :v1 - v0]
        Ye = Ut.view(torch.int16)[:, v0:v1].to(torch.float64) * scv + ofv
        ref = core.lm_fit_torch(Ye, X)
        torch.cuda.synchronize()
        scale = torch.maximum(ref.abs(), ref.square().mean(dim=1, keepdim=True).sqrt())
        assert bool(((out[:, v0:v1] - ref).abs() <= 1e-9 * scale).all()), f"slab {s0}"
        del Ye, ref, scv, ofv, scale
    del Ut, out
    # ---- covariate: [X[:, :3] | z_v]
    Yt = gen(torch, V, n, seed=1, mean=30000.0, sd=50.0)
    Zt = gen(torch, V, n, seed=2, mean=100.0, sd=3.0)
    Xs = X[:, :3]
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    oc = core.lm_fit_cov_torch(Yt, Zt, Xs, status=status)
    torch.cuda.synchronize()
    assert int(status.item()) == 0
    sel = np.sort(rng.choice(V, size=800, replace=False))
    sel[0], sel[-1] = 0, V - 1
    st = torch.from_numpy(sel).cuda()
    want = oracle.lm_loop_cov(Yt[:, st].cpu().numpy().T, Zt[:, st].cpu().numpy().T, Xs)
    assert_rows_match(oc[:, st].cpu().numpy().T, want, RTOL, "covariate sample")
    # ALL voxels: y + c z has the covariate's beta shifted by exactly c, the static betas and rss-derived columns unchanged
    c = 7.25
    Yt.add_(Zt, alpha=c)
    o2 = core.lm_fit_cov_torch(Yt, Zt, Xs)
    torch.cuda.synchronize()

    def close(u, v, tol=1e-7):
        scale = torch.maximum(v.abs(), v.square().mean().sqrt())
        return bool(((u - v).abs() <= tol * scale).all())

    assert close(o2[2 + 3], oc[2 + 3] + c)
    for i in range(3):
        assert close(o2[2 + i], oc[2 + i], 1e-6), i                       # static betas (intercept ~3e4: looser scale)
    assert close(o2[2 * p + 2], oc[2 * p + 2], 1e-9)                      # logLik: same residuals


def test_config5_whole_volume_as_uint16_on_one_gpu(torch_mod, oracle):
    """Config 5 (40 um, 400x560x340 x 200 subjects) is 121.9 GB as doubles and needs 8 GPUs; in its stored type the
    whole volume is 30.5 GB and fits one B200.  V = 76,160,000 exercises the int32 TMA coordinates and 64-bit offsets."""
    torch, core, free = torch_mod
    dims, n = (400, 560, 340), 200
    V = int(np.prod(dims))
    if free < V * (2 * n + 8 * 11) * 1.25:
        pytest.skip("not enough free HBM")
    X = design_cfg2(n)
    nsl, slice_len = dims[0], dims[1] * dims[2]
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    Ut = torch.empty((n, V), dtype=torch.uint16, device="cuda")
    for j in range(n):
        blk = torch.empty(V, dtype=torch.float32, device="cuda").normal_(20000.0, 300.0, generator=g)
        Ut[j] = blk.round_().clamp_(0, 32767).to(torch.int16).view(torch.uint16)
    del blk
    rs = np.random.default_rng(6)
    smin = rs.normal(-0.5, 0.02, (n, nsl))
    sc_h, of_h = np.ascontiguousarray(rs.uniform(1.0, 1.2, (n, nsl)) / 65535.0), np.ascontiguousarray(smin)
    out = core.lm_fit_stored_torch(Ut, X, torch.from_numpy(sc_h).cuda(), torch.from_numpy(of_h).cuda(), slice_len=slice_len)
    torch.cuda.synchronize()
    assert core.last_fit_variant() == "lm_fit_packed_tma_kernel"
    sel = np.sort(np.random.default_rng(0).choice(V, size=1500, replace=False))
    sel[0], sel[-1] = 0, V - 1
    st = torch.from_numpy(sel).cuda()
    Us = Ut.view(torch.int16)[:, st].cpu().numpy().view(np.uint16).T
    sl = sel // slice_len
    Ys = (Us.astype(np.float64) - 0.0) * sc_h.T[sl, :] + of_h.T[sl, :]
    assert_rows_match(out[:, st].cpu().numpy().T, oracle.vertex_lm_loop(Ys, X), RTOL, "cfg5 uint16 sample")
    assert bool(torch.isfinite(out).all())


def test_randomize_host_pipeline_at_a_realistic_size(torch_mod):
    """rmg_randomize from (pageable) host memory with its default blocking -- 8 voxel blocks uploaded through the pinned
    ring while the permutation GEMM runs -- equals the device-resident entry point on the same data; and so does the
    same matrix shipped as float32 stored volumes (rmg_randomize_stored)."""
    torch, core, free = torch_mod
    V, n, R = 1_200_000, 64, 48                              # 614 MB of doubles: above the staging threshold
    X = design_cfg2(n)
    p = X.shape[1]
    Yt = gen(torch, V, n, seed=5).to(torch.float32).to(torch.float64)      # float32-representable, for the stored leg
    Yt[:, 1000:1010] = 0.0                                   # degenerate voxels: statistics -> 0
    mask = ellipsoid_mask((100, 100, 120))
    mt = torch.from_numpy(mask).cuda()
    rng = np.random.default_rng(6)
    idx = np.column_stack([rng.permutation(n) for _ in range(R)]).astype(np.int32)
    cols = list(range(p + 2, 2 * p + 2))
    want = core.randomize_torch(Yt, X, idx, cols, mask=mt).cpu().numpy().T
    Yh = np.asfortranarray(Yt.cpu().numpy().T)               # plain numpy: pageable
    got = core.randomize(Yh, X, idx, cols, mask=mask)
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=0)
    got32 = core.randomize_stored(np.asfortranarray(Yh.astype(np.float32)), X, idx, cols, mask=mask)
    np.testing.assert_allclose(got32, want, rtol=1e-10, atol=0)
    assert np.isfinite(got).all() and (got > 0).all()

