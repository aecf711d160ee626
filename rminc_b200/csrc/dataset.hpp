// dataset.hpp -- the resident-dataset engine (dataset.cu): interfaces shared with the other C-ABI translation units.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "capi_common.hpp"
#include "perm_gemm.cuh"

struct rmg_dataset;

namespace rmg {

// One resident voxel shard as the permutation engine sees it (a dataset shard, or a matrix borrowed from the caller).
struct PermShard {
    int device = 0, sm_count = 148;
    const double *dY = nullptr;
    int64_t cv = 0, ld = 0;
    const double *dM = nullptr;              // this shard's mask (nullable)
    cudaStream_t st = nullptr;               // compute stream
    std::vector<int64_t> blk_v0, blk_cv;     // voxel blocks of the shard (upload granularity)
    std::vector<cudaEvent_t> landed;         // landed[b] fires when block b is resident (empty / null: already there)
    PermPre *pre = nullptr;                  // [2] resident per-voxel sums indexed by `shift` (null: per-call workspace)
    std::vector<char> *pre_done = nullptr;   // [2] per block: sums already computed
    double *ddist = nullptr;                 // device R x ncols result of THIS shard (nullable)
    double *host_part = nullptr;             // host R x ncols result of THIS shard (nullable)
};

// Factor every permuted design X[idx[, r], ]; translates the reference's error contract.  0 or RMG_ERR_*.
int factor_permutations(const double *X, int n, int p, const int32_t *idx, int R, PermSet &pd, char *err, size_t errlen);
int check_perm_args(int p, const int32_t *idx, int R, const int32_t *cols, int ncols, const void *dist, char *err, size_t errlen);

// All R permutations on resident shards (one per device), queued from the calling thread.  sync_host: wait for the
// devices before returning (host results); otherwise everything stays queued on the shards' streams.
int randomize_shards(std::vector<PermShard> &S, int n, int p, const double *X, double mask_lo, double mask_hi,
                     const int32_t *idx, int R, const int32_t *cols, int ncols, int two_sided, bool sync_host, char *err,
                     size_t errlen);

// devices 0..n_gpus-1 (n_gpus <= 0: every visible device)
int resolve_gpus(int n_gpus, std::vector<int> &devices, char *err, size_t errlen);

}  // namespace rmg
