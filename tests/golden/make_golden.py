"""Regenerate tests/golden/ref_golden.npz from the REFERENCE'S OWN object code.

Runs only where /root/reference exists: `python tests/golden/make_golden.py`.  It builds
oracle/_ref/librminc_ref.so (the reference's src/modelling_functions.c, vertex_modelling.c,
minc_anova.c, slice_loop_functions.c compiled where they lie; see oracle/Makefile) and records
the outputs of the reference's vertex_lm_loop and minc2_model(method="lm") on small seeded
inputs.  The inputs are stored too, so the fixture is self-contained on machines that have
neither the reference nor a GPU.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402


def main():
    O.build(ref=True)
    assert O.have_ref(), "needs /root/reference"
    rng = np.random.default_rng(20261019)
    out = {}
    # case A: vertex_lm_loop, y ~ group + age (p = 3), V = 96, n = 24
    n, V = 24, 96
    X = np.column_stack([np.ones(n), np.arange(n) % 2, np.round(rng.normal(60, 10, n), 1)])
    Y = np.asfortranarray(rng.normal(1.0, 0.2, (V, n)) + 0.1 * np.outer(rng.random(V), X[:, 1]))
    Y[5, :] = 0.0  # degenerate all-zero vertex
    out["A_X"], out["A_Y"] = X, Y
    out["A_out"] = O.vertex_lm_loop(Y, X, use_ref=True)
    # case B: minc2_model lm with an integer-labelled mask, default range and maskval = 2
    dims = (4, 6, 5)
    V = int(np.prod(dims))
    n = 12
    X = np.column_stack([np.ones(n), (np.arange(n) >= n // 2).astype(float)])
    Y = np.asfortranarray(30000 + 50 * rng.standard_normal((V, n)) + 40 * np.outer(rng.random(V), X[:, 1]))
    mask = rng.integers(0, 4, V).astype(float)
    out["B_dims"], out["B_X"], out["B_Y"], out["B_mask"] = np.array(dims), X, Y, mask
    out["B_out_default"] = O.minc2_model_lm(Y, dims, X, mask=mask, use_ref=True, fill=0.0)
    out["B_out_maskval2"] = O.minc2_model_lm(Y, dims, X, mask=mask, lo=2, hi=2, use_ref=True, fill=0.0)
    out["B_out_nomask"] = O.minc2_model_lm(Y, dims, X, mask=None, use_ref=True, fill=0.0)
    # case C: rank-deficient design [1, g, 1-g, age]: pivoted betas, trailing zero
    n, V = 16, 8
    g = (np.arange(n) % 2).astype(float)
    X = np.column_stack([np.ones(n), g, 1 - g, np.round(rng.normal(60, 10, n), 1)])
    Y = np.asfortranarray(rng.normal(0, 1, (V, n)))
    out["C_X"], out["C_Y"] = X, Y
    out["C_out"] = O.vertex_lm_loop(Y, X, use_ref=True)
    # case D: per_voxel_anova (slice loop + mask), 3-level factor * covariate
    n, dims = 18, (3, 4, 5)
    V = int(np.prod(dims))
    g3 = np.array(list("abc" * 6))
    age = np.round(rng.normal(60, 10, n), 1)
    X = np.column_stack([np.ones(n), g3 == "b", g3 == "c", age, (g3 == "b") * age, (g3 == "c") * age]).astype(float)
    asgn = np.array([0, 1, 1, 2, 3, 3], dtype=np.int32)
    Y = np.asfortranarray(rng.normal(1.0, 0.3, (V, n)))
    mask = (rng.random(V) > 0.3).astype(float)
    out["D_X"], out["D_Y"], out["D_asgn"], out["D_mask"] = X, Y, asgn, mask
    out["D_out"] = O.anova(Y, X, asgn, mask=mask, use_ref=True, sizes=dims)
    # case E: voxel-wise covariate (a file term on the right-hand side): vertex_lm_loop with data_right,
    # static part [1, group] and no static part ([1 | z]); one vertex with a constant covariate (rank p-1)
    n, V = 20, 40
    Xs = np.column_stack([np.ones(n), np.arange(n) % 2]).astype(float)
    Y = np.asfortranarray(rng.normal(2.0, 0.5, (V, n)))
    Z = np.asfortranarray(rng.normal(100.0, 3.0, (V, n)))
    Z[7, :] = 42.0
    out["E_Xs"], out["E_Y"], out["E_Z"] = Xs, Y, Z
    out["E_out_static"] = O.lm_loop_cov(Y, Z, Xs, use_ref=True)
    out["E_out_nostatic"] = O.lm_loop_cov(Y, Z, None, use_ref=True)
    # case F: the reference's own neighbour_list + graph_tfce on a voxel grid (what tests/testthat/test_mincTFCE.R:153-170
    # equates with mincTFCE): a 6 x 5 x 4 volume (x fastest), weights = voxel volume, hmax an exact multiple of the step
    x, y, z = 6, 5, 4
    for conn in (6, 18, 26):
        nl = O.neighbour_list(x, y, z, conn, use_ref=True)
        ptr, idx = O.adjacency_csr(nl)
        out[f"F_ptr{conn}"], out[f"F_idx{conn}"] = ptr, idx
    vol = np.round(rng.normal(0, 1, x * y * z), 2)
    vol[vol > 2.0] = 1.5
    vol[vol < -2.0] = -1.5
    vol[31], vol[77] = 2.0, -2.0                      # hmax = 2.0 on both sides = 8 steps of 0.25
    out["F_dims"], out["F_vol"] = np.array([z, y, x]), vol
    for conn in (6, 18, 26):
        for side in ("positive", "negative", "both"):
            out[f"F_tfce{conn}_{side}"] = O.graph_tfce(vol, out[f"F_ptr{conn}"], out[f"F_idx{conn}"], nsteps=8, side=side,
                                                       weights=np.full(vol.size, 0.001), use_ref=True)
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "ref_golden.npz"), **out)
    print("wrote ref_golden.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
