"""Multi-GPU execution: one process per GPU, torch.distributed for the plumbing.

The path shards naturally over voxels, replacing the reference's
process-level sharding (create_parallel_mask + mclapply, R/minc_voxel_statistics.R:848-906,
and groupingVector over permutations, :1503-1509):

* mincLm / vertexLm / anatLm: every voxel row depends only on X and its own y -> contiguous
  voxel shards, NO data-path collective; rows are stitched (gathered) only if the caller wants
  the whole matrix on one rank.
* mincRandomize: every rank evaluates all R permutations on its voxel shard and keeps a local
  R x ncols maximum; the single exchange step of the path is one all-reduce(MAX) of that small
  buffer (32 KB at R = 1000, 4 columns) over NCCL / NVLink.

The per-shard compute is injected (`fit` / `perm`) so the host logic can be exercised on CPU
with the gloo backend (tests/test_parallel_gloo.py); the default is the CUDA path.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence, Tuple

import numpy as np


def shard_bounds(V: int, world: int) -> np.ndarray:
    """Contiguous voxel shards with even boundaries (keeps every shard 16-byte aligned).

    Any contiguous partition gives identical rows (the reference's groupingVector,
    R/minc_parallel.R:613-622, makes the FIRST group the short one; here the last is)."""
    b = np.array([min(V, ((V * g // world) + 1) & ~1) for g in range(world + 1)], dtype=np.int64)
    b[0], b[-1] = 0, V
    return b


def my_shard(V: int, rank: int, world: int) -> Tuple[int, int]:
    b = shard_bounds(V, world)
    return int(b[rank]), int(b[rank + 1])


def randomize_allreduce_max(local, group=None):
    """The path's one collective: element-wise max over ranks of the (ncols, R) buffer."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(local, op=dist.ReduceOp.MAX, group=group)
    return local


def randomize_sharded(Yt_local, X, idx, cols, two_sided=True, mask_local=None, mask_lo=1.0, mask_hi=99999999.0,
                      group=None, perm: Optional[Callable] = None):
    """mincRandomize over voxel shards.  Yt_local: this rank's (n, V_local) slab.  Returns the
    (ncols, R) tensor of global maxima on every rank."""
    if perm is None:
        from . import core
        perm = core.randomize_torch
    local = perm(Yt_local, X, idx, cols, two_sided=two_sided, mask=mask_local, mask_lo=mask_lo, mask_hi=mask_hi)
    return randomize_allreduce_max(local, group)


def lm_fit_sharded(Yt_local, X, mask_local=None, mask_lo=1.0, mask_hi=99999999.0, fit: Optional[Callable] = None):
    """mincLm on this rank's voxel shard: (2p+3, V_local).  No collective."""
    if fit is None:
        from . import core
        fit = core.lm_fit_torch
    return fit(Yt_local, X, mask=mask_local, mask_lo=mask_lo, mask_hi=mask_hi)


def lm_fit_cov_sharded(Yt_local, Zt_local, Xs=None, mask_local=None, mask_lo=1.0, mask_hi=99999999.0,
                       fit: Optional[Callable] = None):
    """Voxel-wise covariate design [Xs | z_v] on this rank's shard of both operands.  No collective."""
    if fit is None:
        from . import core
        fit = core.lm_fit_cov_torch
    return fit(Yt_local, Zt_local, Xs, mask=mask_local, mask_lo=mask_lo, mask_hi=mask_hi)


def lm_fit_stored_sharded(Ut_local, X, scale, offset, voxel_min: float, slice_len: int, v_offset: int, mask_local=None,
                          mask_lo=1.0, mask_hi=99999999.0, fit: Optional[Callable] = None):
    """The fit on this rank's shard of volumes in their stored type.  `scale` / `offset` are the WHOLE volume's
    (n, nslices) tables (replicated, a few hundred KB); v_offset = first voxel of the shard, so every voxel looks up
    the slice it has in the file.  No collective."""
    if fit is None:
        from . import core
        fit = core.lm_fit_stored_torch
    return fit(Ut_local, X, scale, offset, voxel_min=voxel_min, slice_len=slice_len, mask=mask_local, mask_lo=mask_lo,
               mask_hi=mask_hi, v_offset=v_offset)


def gather_rows(out_local, V: int, group=None, dst: int = 0):
    """Stitch the per-rank (ncol, V_local) blocks back into (ncol, V) on rank `dst` (what
    Reduce(rbind) + result_fleshed_out do in R/minc_voxel_statistics.R:873,898-905)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return out_local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    b = shard_bounds(V, world)
    ncol = out_local.shape[0]
    full = torch.empty((ncol, V), dtype=out_local.dtype, device=out_local.device) if rank == dst else None
    if rank == dst:
        full[:, b[rank]:b[rank + 1]] = out_local
        for r in range(world):
            if r == dst or b[r + 1] == b[r]:
                continue
            buf = torch.empty((ncol, int(b[r + 1] - b[r])), dtype=out_local.dtype, device=out_local.device)
            dist.recv(buf, src=r, group=group)
            full[:, b[r]:b[r + 1]] = buf
    elif out_local.shape[1] > 0:
        dist.send(out_local.contiguous(), dst=dst, group=group)
    return full
