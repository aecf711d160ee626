// perm_gemm.cuh -- permutation test as one stacked FP64 tensor-core GEMM (piece K3).
//
// For permutation r the refit needs e_r(v) = Q_r' y_v for every voxel v, where Q_r is the
// thin orthonormal factor of the permuted design X[idx_r, ].  Stacking the Q_r' of many
// permutations gives A (rows = permutation x component, cols = subjects) and the whole test is
//     E = A (R*KP x n)  *  Y' (n x V)                      2*KP*n*V flop per permutation
// after which t_{r,i}(v) = (Rinv_r e_r)_i / sqrt(ct_{r,i} * rss_r(v) / (n-p)) with
// rss_r(v) = |y_v - y0_v|^2 - |e_r - y0_v Q_r'1|^2.  Only the maximum over voxels of each
// statistic is kept (R/minc_voxel_statistics.R:1484-1488), so nothing of size R x V is ever
// written: the epilogue reduces straight into an R x ncols buffer.
//
// Kernel shape (sm_100a; tcgen05/UMMA has no FP64 kind, so the tensor-core path for FP64 is
// mma.sync DMMA m8n8k4):
//   CTA tile  = 4x2 consumer warps = (32*PG permutations x KP components) x 64 voxels,
//               + 1 producer warp.  Persistent CTAs; voxel tiles outer, permutation tiles
//               inner, so each CTA's 64 x n slab of Y is fetched from HBM once and re-read from
//               L2 for the other permutation tiles, while A (16 MB at R=1000) lives in L2.
//   staging   = cp.async.bulk (TMA 1-D) into a 3-4 stage shared-memory ring (32 or 16 subjects per
//               stage) guarded by mbarriers.  A is pre-packed on the host in MMA-fragment order, so one
//               contiguous bulk copy per stage lands it conflict-free; Y rows are copied with
//               a pitch of 68 doubles (== 4 mod 16), which makes the B-fragment loads
//               conflict-free without a transpose.
//   epilogue  = per (permutation, voxel): rss, beta for the requested t-columns, the
//               monotone proxy sign(beta) beta^2 / rss, non-finite -> 0, running max in
//               registers, one ordered-integer atomicMax per (permutation, column) per warp.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "design.hpp"
#include "lm_common.cuh"

namespace rmg {

// ---- order-preserving map double -> uint64 (so atomicMax on integers is a max on doubles)
__host__ __device__ __forceinline__ unsigned long long ordered_key(double x)
{
#ifdef __CUDA_ARCH__
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
#else
    unsigned long long b;
    std::memcpy(&b, &x, 8);
#endif
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double ordered_value(unsigned long long k)
{
    unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)b);
#else
    double x;
    std::memcpy(&x, &b, 8);
    return x;
#endif
}

// new_mod[!is.finite(new_mod)] <- 0; abs() if two-sided  (R/minc_voxel_statistics.R:1478,1484)
__device__ __forceinline__ double perm_stat_filter(double s, int two_sided)
{
    if (!isfinite(s)) s = 0.0;
    return two_sided ? fabs(s) : s;
}

struct PermGemmArgs {
    const double *Y;
    int64_t V, ldY;
    int n, p;
    const double *mask;
    double mask_lo, mask_hi;
    const int32_t *hcols;  // HOST copy of the requested result columns
    int ncols, two_sided;
    unsigned long long *keys;  // device, [ncols][R] ordered keys, initialised to key(-inf)
    int R;
};

constexpr int kPermNT = 64;       // voxels per CTA tile
constexpr int kPermPitch = 68;    // doubles per staged Y row (68 == 4 mod 16: conflict-free B fragments)
constexpr int kPermMaxCols = 8;

// per-permutation constants read by the epilogue (KP <= 8)
struct PermConst {
    double q1[8];
    double Rinv[64];   // row-major 8 x 8
    double ct[8];
    double rdf;
    double c0;         // the constant value of Q_r[:, 0] when the intercept row is skipped
    double pad[6];
};

// The R permuted designs of one call.  A column of idx that is a true permutation only re-orders the
// rows of X (X_pi = P X => Q_pi = P Q, same R / pivoting / rank / (R'R)^-1), so it is represented by
// the base factorisation plus the index column; only columns with repeated indices (replace = TRUE)
// carry a factorisation of their own.
struct PermSet {
    Design base;
    std::vector<Design> own;        // own[r].n == 0  <=>  idx[, r] is a true permutation
    const int32_t *idx = nullptr;   // n x R, 0-based
    int R = 0;
    bool all_intercept = true;
    bool is_perm(int r) const { return own[r].n == 0; }
    const Design &design(int r) const { return is_perm(r) ? base : own[r]; }
    const double *qcol_own(int r, int k) const { return own[r].Q.data() + (size_t)k * own[r].n; }
    // materialise permutation r as a stand-alone Design (refit path only)
    Design materialise(int r) const
    {
        if (!is_perm(r)) return own[r];
        Design D = base;
        const int n = base.n;
        for (int c = 0; c < base.p; ++c)
            for (int i = 0; i < n; ++i) D.Q[i + (size_t)c * n] = base.Q[idx[i + (size_t)r * n] + (size_t)c * n];
        return D;
    }
};

bool perm_gemm_supported(int n, int p, int64_t V, int64_t ldY, const double *dY, const int32_t *hcols, int ncols);

// ---- the GEMM path in pieces, so that one call can drive several devices (voxel shards) and several resident
//      voxel blocks per device from ONE host-side packing of the permuted designs:
//        perm_pack_plan      geometry of the stacked operand (invariant-row skip, tile shape) + pinned host staging
//        perm_pack_batch     host threads pack one batch of permutations in MMA-fragment order
//        perm_dev_begin      per device: workspace (packed operand, per-permutation constants, per-voxel sums, flags)
//        perm_dev_prepass    per resident voxel block: y0, |y - y0|^2, sum(y - y0)   (skipped when the caller holds them)
//        perm_dev_upload     per device and batch: packed operand -> device on a side stream
//        perm_dev_launch     per device, batch and voxel block: the GEMM + fused statistics
//        perm_dev_end        per device: 0 floor, exact refits of flagged voxels, proxy -> t
struct PermKernelParams {
    const double *Y;
    int64_t V, ldY;
    int n, p, R, ncols, two_sided, shift;   // R: permutations of THIS launch (a batch)
    int Rstride;                             // permutations of the whole call (row stride of keys)
    int tcol[kPermMaxCols];          // coefficient index of each requested t-column
    const double *Apack;             // [n_pt][n_kc][A chunk] fragment-major
    const PermConst *pc;             // [R]
    const double *y0;                // [V] first sample (0 when !shift)
    const double *yy;                // [V] |y - y0|^2, or -1 for voxels outside the mask
    const double *ys;                // [V] sum_j (y_j - y0)  (only when the intercept row is skipped)
    unsigned long long *keys;        // [ncols][R]
    // geometry of Apack, so the (rare) exact path can read Q_r[j][k] back out of the packed operand
    int inv, kg, pg, wm, warps_m, a_chunk, perms_per_cta;
    unsigned char *vflag;            // [V] set when some permutation's one-pass rss lost too many digits
    int n_pt, n_kc;
    // per-permutation constants of the statistics epilogue, compact: [n_pt * perms_per_cta][epi_stride] doubles =
    // q1[p], c0, then Rinv row of every requested column (ncols x p); one bulk copy per permutation tile
    const double *epi;
    int epi_stride;
};

// Per-voxel quantities every permutation shares (the prepass output), device-resident.  A resident dataset keeps
// them across calls (they depend on Y, the mask and `shift` only).
struct PermPre {
    double *y0 = nullptr, *yy = nullptr, *ys = nullptr;
    int *any_zero = nullptr;         // raised when some voxel is outside the mask (the global 0 floor, Q6)
};

struct PermPack {
    int R = 0, n = 0, p = 0;
    int INV = 0, KG = 0, PG = 0, WM = 0, KC = 16, WARPS_M = 4, A_CHUNK = 0, perms_per_cta = 0, n_pt = 0, n_kc = 0;
    int tiles_per_batch = 1, first_tiles = 1, nbatch = 0;
    int shift = 0;                   // every permuted design spans the constant: voxels are centred on their first sample
    std::vector<double> c0;          // [R] constant value of Q_r[:, 0] (INV)
    size_t apack_doubles = 0;
    double *hA = nullptr;            // pinned host [n_pt][n_kc][A_CHUNK]
    PermConst *hpc = nullptr;        // pinned host [R]
    double *hE = nullptr;            // pinned host [n_pt * perms_per_cta][epi_stride]
    int ncols = 0, epi_stride = 0;
    int tcol[kPermMaxCols] = {0};    // coefficient index of each requested t-column
    void *lease = nullptr;           // pinned-pool lease behind hA / hpc (perm_pack_release)
    // Batch 0 is a single permutation tile, so that the first launch waits for the packing of 8 * PG * 4 permutations
    // only (the host packs the next, full-size batch behind it); batches b >= 1 hold tiles_per_batch tiles.
    int batch_pt0(int b) const { return b <= 0 ? 0 : std::min(n_pt, first_tiles + (b - 1) * tiles_per_batch); }
    int batch_pt1(int b) const { return std::min(n_pt, first_tiles + b * tiles_per_batch); }
    int batch_r0(int b) const { return std::min(R, batch_pt0(b) * perms_per_cta); }
    int batch_r1(int b) const { return std::min(R, batch_pt1(b) * perms_per_cta); }
};

cudaError_t perm_pack_plan(const PermSet &ps, int n, int p, const int32_t *hcols, int ncols, PermPack &pk);
void perm_pack_batch(const PermSet &ps, PermPack &pk, int b);
void perm_pack_release(PermPack &pk);

struct PermDevice {
    int device = 0, sm_count = 148;
    cudaStream_t st = nullptr;       // compute stream (the caller's)
    cudaStream_t cs = nullptr;       // side stream of the operand uploads
    char *ws = nullptr;              // one stream-ordered allocation behind everything below
    PermKernelParams P{};            // whole-shard parameters; launches offset them per voxel block / batch
    PermPre pre;                     // points into ws unless the caller supplied resident sums
    bool own_pre = true;
    std::vector<cudaEvent_t> ev;     // ev[b]: batch b's operand has landed
    int64_t V = 0;
};

// keys: device [ncols][R], initialised to key(-inf).  resident_pre: per-voxel sums kept by the caller (else null).
cudaError_t perm_dev_begin(PermDevice &d, const PermPack &pk, int device, int sm_count, const PermGemmArgs &a, cudaStream_t st,
                           const PermPre *resident_pre, int *launches);
cudaError_t perm_dev_prepass(PermDevice &d, const PermGemmArgs &a, int64_t v0, int64_t cv, int *launches);
cudaError_t perm_dev_upload(PermDevice &d, const PermPack &pk, int b);
// batches [b0, b1) of permutations on the voxel block [v0, v0 + cv) of the shard, in one launch
cudaError_t perm_dev_launch(PermDevice &d, const PermPack &pk, int b0, int b1, int64_t v0, int64_t cv, int *launches);
cudaError_t perm_dev_end(PermDevice &d, const PermPack &pk, int *launches);
void perm_dev_release(PermDevice &d);     // waits for the side stream; frees ws in stream order

// One device, one resident block: plan + begin + prepass + (pack, upload, launch per batch) + end.
cudaError_t perm_gemm_run(const PermGemmArgs &a, const PermSet &ps, int sm_count, cudaStream_t st, int *launches);

}  // namespace rmg
