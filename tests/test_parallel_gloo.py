"""The N > 1 host logic on CPU: world_size-2 gloo process group, voxel shards, the single
all-reduce(MAX) of mincRandomize and the row stitching of mincLm.  The per-shard compute is the
CPU oracle here (tests may use it as the checker); on the GPU box the same functions run the CUDA
path (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from helpers import cfg1_data
    from oracle import oracle as O
    from rminc_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        Y, X, mask = cfg1_data(dims=(9, 11, 7))          # V = 693 (odd): ragged shards
        V, n = Y.shape
        a, b = parallel.my_shard(V, rank, world)
        Yt_local = torch.from_numpy(np.ascontiguousarray(Y[a:b].T))
        m_local = torch.from_numpy(mask[a:b].copy())
        rng = np.random.default_rng(1)
        idx = np.column_stack([rng.permutation(n) for _ in range(9)])
        cols = [4, 5]

        def fit(Yt, X, mask=None, mask_lo=1.0, mask_hi=99999999.0):
            out = O.minc2_model_lm(Yt.numpy().T, (1, 1, Yt.shape[1]), X, mask=None if mask is None else mask.numpy(),
                                   lo=mask_lo, hi=mask_hi)
            return torch.from_numpy(np.ascontiguousarray(out.T))

        def perm(Yt, X, idx, cols, two_sided=True, mask=None, mask_lo=1.0, mask_hi=99999999.0):
            d = O.randomize(Yt.numpy().T, X, idx, cols, two_sided=two_sided,
                            mask=None if mask is None else mask.numpy(), lo=mask_lo, hi=mask_hi)
            return torch.from_numpy(np.ascontiguousarray(d.T))

        # the next-row paths shard the same way: covariate (two operands) and stored type (slices keep their global index)
        def fit_cov(Yt, Zt, Xs, mask=None, mask_lo=1.0, mask_hi=99999999.0):
            out = O.lm_loop_cov(Yt.numpy().T, Zt.numpy().T, Xs, mask=None if mask is None else mask.numpy(), lo=mask_lo, hi=mask_hi)
            return torch.from_numpy(np.ascontiguousarray(out.T))

        def fit_stored(Ut, X, scale, offset, voxel_min=0.0, slice_len=None, mask=None, mask_lo=1.0, mask_hi=99999999.0,
                       v_offset=0):
            U = Ut.numpy().T.astype(np.float64)
            sl = (v_offset + np.arange(U.shape[0])) // slice_len
            Yx = (U - voxel_min) * scale.numpy().T[sl, :] + offset.numpy().T[sl, :]
            return fit(torch.from_numpy(np.ascontiguousarray(Yx.T)), X, mask=mask, mask_lo=mask_lo, mask_hi=mask_hi)

        rs = np.random.default_rng(7)
        Z = np.asfortranarray(rs.normal(10, 1, Y.shape))
        U = rs.integers(0, 4000, Y.shape).astype(np.int32)            # (torch has no uint16 arithmetic on CPU)
        slice_len, nsl = 100, 7
        scale, offset = rs.uniform(1, 2, (nsl, n)) / 4000.0, rs.normal(0, 0.1, (nsl, n))
        cov_local = parallel.lm_fit_cov_sharded(Yt_local, torch.from_numpy(np.ascontiguousarray(Z[a:b].T)), X, fit=fit_cov)
        st_local = parallel.lm_fit_stored_sharded(torch.from_numpy(np.ascontiguousarray(U[a:b].T)), X,
                                                  torch.from_numpy(np.ascontiguousarray(scale.T)),
                                                  torch.from_numpy(np.ascontiguousarray(offset.T)), 0.0, slice_len, a,
                                                  fit=fit_stored)
        cov_full, st_full = parallel.gather_rows(cov_local, V), parallel.gather_rows(st_local, V)

        out_local = parallel.lm_fit_sharded(Yt_local, X, mask_local=m_local, fit=fit)
        full = parallel.gather_rows(out_local, V)
        d = parallel.randomize_sharded(Yt_local, X, idx, cols, two_sided=True, mask_local=m_local, perm=perm)
        res = dict(rank=rank, shard=(a, b), dist=d.numpy().T.copy())
        if rank == 0:
            res["cov"], res["want_cov"] = cov_full.numpy().T.copy(), O.lm_loop_cov(Y, Z, X)
            sl = np.arange(V) // slice_len
            res["stored"] = st_full.numpy().T.copy()
            res["want_stored"] = O.vertex_lm_loop((U.astype(np.float64) - 0.0) * scale[sl, :] + offset[sl, :], X)
            res["full"] = full.numpy().T.copy()
            res["want_full"] = O.minc2_model_lm(Y, (9, 11, 7), X, mask=mask)
            res["want_dist"] = O.randomize(Y, X, idx, cols, two_sided=True, mask=mask)
        q.put(res)
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_and_align():
    sys.path.insert(0, ROOT)
    from rminc_b200 import parallel
    for V in (0, 1, 2, 7, 693, 7076160, 76160000):
        for w in (1, 2, 3, 4, 8):
            b = parallel.shard_bounds(V, w)
            assert b[0] == 0 and b[-1] == V and (np.diff(b) >= 0).all()
            assert all(int(x) % 2 == 0 for x in b[1:-1])
    assert list(np.diff(parallel.shard_bounds(76160000, 8))) == [9520000] * 8     # config 5


@pytest.mark.timeout(300)
def test_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort(key=lambda r: r["rank"])
    assert res[0]["shard"][1] == res[1]["shard"][0] and res[1]["shard"][1] == 693
    # stitched rows == unsharded fit, bit for bit (each row depends only on X and its own y)
    assert np.array_equal(res[0]["full"], res[0]["want_full"], equal_nan=True)
    assert np.array_equal(res[0]["cov"], res[0]["want_cov"], equal_nan=True)
    assert np.array_equal(res[0]["stored"], res[0]["want_stored"], equal_nan=True)
    # all-reduce(MAX) of the two shards' maxima == maxima over all voxels, on every rank
    for r in res:
        assert np.array_equal(r["dist"], res[0]["want_dist"])
