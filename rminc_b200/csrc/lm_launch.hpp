// lm_launch.hpp -- host-visible launch interface of the fit kernels.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "lm_common.cuh"

namespace rmg {

// Device pointers only.  Y: V x n (ld ldY), out: V x (2p+3) (ld ldOut), mask: V or null.
struct LmArgs {
    const double *Y;
    int64_t V, ldY;
    int n, p;
    const double *Qt;          // [n][KP] row-major, KP = lm_padded_p(p), columns >= rank are zero
    const DevDesign *design;
    const double *mask;
    double mask_lo, mask_hi;
    double *out;
    int64_t ldOut;
    int out_vec2 = 0;          // set by launch_lm_fit: out is 16-byte aligned with an even ld
    int ncol_out = 0;          // result columns: 0 -> 2p+3 (lm); nterms for the anova epilogue
    const double *rowscale = nullptr;  // n doubles sqrt(w_j) for weighted fits (direct-load kernel only), else null
    // mask compaction (direct-load kernel): ascending list of voxel PAIRS with at least one selected
    // voxel and their count, both device-resident; null = process every pair
    const uint32_t *pair_list = nullptr;
    const unsigned int *n_pairs = nullptr;
    int l2_hint = 0;           // direct-load kernel: L2 prefetch size of the sample loads in bytes (0 = the default policy)
};

enum class LmVariant { Auto = 0, Tma = 1, Ldg = 2 };

struct LmPlan {
    LmVariant variant = LmVariant::Auto;
    int split = 0;       // LDG path: slices of the subject axis (0 = choose)
    int sm_count = 148;
    int tma_cfg = -1;    // TMA path: tile-shape variant (-1 = choose by V; see lm_kernels.cu)
    int no_compact = 0;  // direct-load path: 1 = never compact masked voxels (debug / comparison)
    int q_stream = -1;   // TMA path: 1 = stream Q' with Y, 0 = keep Q' resident in smem, -1 = choose
    int stored_exact = 0;  // stored uint16, TMA path: 1 = expand with the reader's three roundings instead of one FMA
};

struct LmLaunchInfo {
    int launches = 0;
    int used_tma = 0;
    int tma_cfg = -1;
    int split = 1;
};

// Reusable device scratch owned by the caller (mask tile flags / compacted pair list).
struct LmScratch {
    unsigned char *flags = nullptr;
    size_t flags_cap = 0;
    cudaError_t ensure_flags(size_t n, cudaStream_t st)
    {
        if (n <= flags_cap) return cudaSuccess;
        release(st);
        cudaError_t e = cudaMallocAsync(reinterpret_cast<void **>(&flags), n, st);
        if (e == cudaSuccess) flags_cap = n;
        return e;
    }
    cudaError_t ensure_bytes(size_t n, cudaStream_t st) { return ensure_flags(n, st); }
    // stream-ordered: safe while kernels that read the flags are still queued on st
    void release(cudaStream_t st)
    {
        if (flags) cudaFreeAsync(flags, st);
        flags = nullptr;
        flags_cap = 0;
    }
};

int lm_padded_p(int p);

// Voxel-wise covariate designs [Xs | z_v] (lm_cov.cu).  Z has Y's shape and leading dimension; Qt / design
// describe the static part Xs (n x ps, full column rank); out is V x (2(ps+1)+3).
struct CovArgs {
    const double *Y, *Z;
    int64_t V, ldY;
    int n, ps;
    const double *Qt;          // [n][lm_padded_p(ps)] row-major
    const DevDesign *design;
    const double *mask;
    double mask_lo, mask_hi;
    double *out;
    int64_t ldOut;
    int *status;               // device int (nullable): set to p when an unmasked voxel's covariate is all zero
    const int32_t *zrow;       // device [n] (nullable): subject j pairs with covariate plane zrow[j] -- a permuted design
                               // P [Xs | z_v] (mincRandomize re-samples the predictor rows, R/minc_voxel_statistics.R:1468-1474)
};

cudaError_t launch_lm_cov(const CovArgs &a, const LmPlan &plan, cudaStream_t st, LmLaunchInfo *info);

// The fit on volumes in their stored type (lm_packed.cu).  U: V x n samples of `dtype`, leading dimension ldU
// (in samples).  A sample's double is ((double)u - voxel_min) * scale[slice, j] + offset[slice, j] with
// slice = (v_base + v) / slice_len; scale / offset are device arrays [nslices x n], slice fastest (null for f32).
constexpr int kStoredU16 = 1, kStoredF32 = 2;   // == RMG_STORED_U16 / RMG_STORED_F32
struct PackedArgs {
    const void *U;
    int dtype;
    int64_t V, ldU;
    int n, p;
    const double *Qt;
    const DevDesign *design;
    const double *mask;
    double mask_lo, mask_hi;
    double *out;
    int64_t ldOut;
    const double *scale, *offset;
    double cmagic;             // 2^52 + voxel_min (u16)
    int64_t slice_len, nslices, v_base;
    int out_vec2 = 0;
    int zero = 0;              // always 0: a kernel parameter the compiler cannot fold (low words of built doubles)
    int shift_mode = 0;        // FOLD kernels: 0 = tiles / warps vote for per-voxel shifts, 1 = never, 2 = always (RMG_STORED_SHIFT)
};

cudaError_t launch_lm_packed(const PackedArgs &a, const LmPlan &plan, LmScratch &scr, cudaStream_t st, LmLaunchInfo *info);

// Stored samples -> doubles ((u - voxel_min) * scale + offset, the same three roundings as the fused kernels) into a
// resident V x n matrix; scale / offset as in PackedArgs (device, [nslices x n]); cmagic = 2^52 + voxel_min.
cudaError_t launch_expand_stored(const void *U, int dtype, int64_t V, int64_t ldU, int n, const double *scale,
                                 const double *offset, double cmagic, int64_t slice_len, int64_t nslices, int64_t v_base,
                                 double *Y, int64_t ldY, cudaStream_t st);

cudaError_t launch_lm_fit(const LmArgs &a, const LmPlan &plan, LmScratch &scr, cudaStream_t st,
                          LmLaunchInfo *info);

}  // namespace rmg
