// lm_kernels.cu -- the fit kernels (SURVEY.md pieces K1/K2), hand-written for sm_100a.
//
// Replaces, for every voxel at once, the per-voxel body of minc2_model's "lm" branch
// (src/modelling_functions.c:1057-1175), vertex_lm_loop (src/vertex_modelling.c:104-158)
// and voxel_lm (src/modelling_functions.c:448-576).
//
// Data layout in HBM: Y is V x n column-major (ld = ldY): subject j's volume is the
// contiguous run Y[j*ldY .. j*ldY+V).  One voxel's n samples are therefore strided by
// ldY, and consecutive voxels of one subject are contiguous -- every warp load is a
// fully coalesced row segment.  The kernel streams Y exactly once (8n bytes/voxel),
// accumulates e = Q_k'(y - y0) and |y - y0|^2 per voxel in registers, and a fused
// epilogue turns them into the 2p+3 result columns.  Arithmetic intensity is
// (2p+2) flop per 8 bytes, so the roofline that bounds it is HBM bandwidth.
//
// Two implementations of the same contract:
//   lm_fit_tma_kernel : persistent CTAs; one producer thread issues 2-D TMA tile loads
//                       (cp.async.bulk.tensor, FP64 boxes of NJ subjects x up to 256 voxels)
//                       plus the matching rows of Q' into a shared-memory ring guarded by
//                       mbarriers; consumer warps reduce out of shared memory.  Default
//                       whenever TMA can describe Y (16-byte aligned base and row pitch).
//   lm_fit_ldg_kernel : thread-per-voxel(-pair) direct streaming loads with an optional
//                       split of the subject axis across threadIdx.y (K2, "split-n") and a
//                       shared-memory reduction, for layouts TMA cannot describe (odd ldY,
//                       unaligned base), p > 16, or on request (RMG_FIT_VARIANT=ldg).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <cstdio>

#include "async_ptx.cuh"
#include "lm_common.cuh"
#include "lm_launch.hpp"

namespace rmg {

// ------------------------------------------------------------------ mask tile flags
// flags[t] = 1 iff any voxel of tile t (tv voxels) is selected by the mask.
__global__ void mask_tile_flags_kernel(const double *__restrict__ mask, int64_t V, int tv, double lo, double hi,
                                       unsigned char *__restrict__ flags, int64_t ntiles)
{
    const int64_t t = blockIdx.x;
    if (t >= ntiles) return;
    const int64_t v0 = t * tv;
    int any = 0;
    for (int i = threadIdx.x; i < tv; i += blockDim.x) {
        const int64_t v = v0 + i;
        if (v < V && mask_selects(mask[v], lo, hi)) any = 1;
    }
    any = __syncthreads_or(any);
    if (threadIdx.x == 0) flags[t] = (unsigned char)any;
}

// ------------------------------------------------------------------ mask compaction
// With a mask only the selected voxels need their n samples streamed.  Three small kernels turn
// the mask into an ascending list of voxel pairs (v, v+1) that contain a selected voxel; pairs
// with none get their +0.0 result rows written here and are never visited again.
constexpr int kCompactBlock = 1024;   // pairs per block

__device__ __forceinline__ bool pair_selected(const double *__restrict__ mask, int64_t pair, int64_t V, double lo,
                                              double hi)
{
    const int64_t v = pair * 2;
    if (v >= V) return false;
    const double2 m = *reinterpret_cast<const double2 *>(mask + v);   // V even, mask 16-byte aligned (host checks)
    return mask_selects(m.x, lo, hi) || mask_selects(m.y, lo, hi);
}

__global__ void __launch_bounds__(kCompactBlock)
mask_count_kernel(const double *__restrict__ mask, int64_t V, double lo, double hi, unsigned int *__restrict__ counts)
{
    const int64_t pair = (int64_t)blockIdx.x * kCompactBlock + threadIdx.x;
    const int c = __syncthreads_count(pair_selected(mask, pair, V, lo, hi));
    if (threadIdx.x == 0) counts[blockIdx.x] = (unsigned int)c;
}

// exclusive scan of the per-block counts by one block; total -> *n_pairs
__global__ void __launch_bounds__(1024)
mask_scan_kernel(unsigned int *__restrict__ counts, int nblocks, unsigned int *__restrict__ n_pairs)
{
    __shared__ unsigned int warp_sums[32];
    __shared__ unsigned int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned int x = i < nblocks ? counts[i] : 0u;
        unsigned int incl = x;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int y = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += y;
        }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned int w = warp_sums[threadIdx.x], wi = w;
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int y = __shfl_up_sync(0xffffffffu, wi, o);
                if (threadIdx.x >= o) wi += y;
            }
            warp_sums[threadIdx.x] = wi - w;   // exclusive
        }
        __syncthreads();
        const unsigned int excl = carry + warp_sums[threadIdx.x >> 5] + incl - x;
        if (i < nblocks) counts[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_pairs = carry;
}

__global__ void __launch_bounds__(kCompactBlock)
mask_compact_kernel(const double *__restrict__ mask, int64_t V, double lo, double hi,
                    const unsigned int *__restrict__ offsets, uint32_t *__restrict__ list, double *__restrict__ out,
                    int64_t ldOut, int ncol)
{
    __shared__ unsigned int warp_sums[32];
    const int64_t pair = (int64_t)blockIdx.x * kCompactBlock + threadIdx.x;
    const bool sel = pair_selected(mask, pair, V, lo, hi);
    const unsigned int ballot = __ballot_sync(0xffffffffu, sel);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_sums[warp] = __popc(ballot);
    __syncthreads();
    if (warp == 0) {
        unsigned int w = warp_sums[lane], wi = w;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int y = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += y;
        }
        warp_sums[lane] = wi - w;
    }
    __syncthreads();
    if (sel) {
        list[offsets[blockIdx.x] + warp_sums[warp] + __popc(ballot & ((1u << lane) - 1u))] = (uint32_t)pair;
    } else if (pair * 2 < V) {
        double2 *o = reinterpret_cast<double2 *>(out + pair * 2);   // out 16-byte aligned, ldOut even (host checks)
        for (int c = 0; c < ncol; ++c) o[(int64_t)c * (ldOut / 2)] = make_double2(0.0, 0.0);
    }
}

// One box of NJ subject rows out of shared memory into the two voxels' running sums.
template <int KP, int NJ, int ROWW, bool GUARD>
__device__ __forceinline__ void consume_box(const double *__restrict__ tile, const double *__restrict__ q,
                                            VoxelSums<KP> &s0, VoxelSums<KP> &s1, int nvalid)
{
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
        const double2 y = *reinterpret_cast<const double2 *>(tile + jj * ROWW);
        double ya = y.x - s0.y0, yb = y.y - s1.y0;
        if (GUARD && jj >= nvalid) ya = yb = 0.0;
        s0.yy = fma(ya, ya, s0.yy);
        s1.yy = fma(yb, yb, s1.yy);
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            const double qk = q[jj * KP + k];
            s0.e[k] = fma(qk, ya, s0.e[k]);
            s1.e[k] = fma(qk, yb, s1.e[k]);
        }
    }
}

// ------------------------------------------------------------------ TMA kernel
// CTA = 1 producer warp + CW consumer warps.  Tile = TV = CW*64 voxels (2 per consumer
// thread).  One pipeline stage holds NJ subjects: the Y box(es) -- 2-D TMA tiles of NJ x
// min(TV,256) doubles (a TMA box dimension is at most 256 elements) -- and the matching NJ
// rows of Q' (a 1-D bulk copy), so nothing of size n lives in shared memory and any n fits.
// Shared memory: [ring: STAGES x (NJ x TV + NJ x KP) doubles][SmemDesign][mbarriers]
// QSTREAM = false keeps the whole Q' (n x KP) resident in shared memory instead (loaded once per
// CTA); it needs n*KP*8 bytes of shared memory but issues no per-stage Q copy.
template <int KP, int CW, int NJ, int STAGES, int MINB, bool QSTREAM>
__global__ void __launch_bounds__((CW + 1) * 32, MINB)
lm_fit_tma_kernel(const __grid_constant__ CUtensorMap ymap, LmArgs a, const unsigned char *__restrict__ flags)
{
    constexpr int TV = CW * 64;
    constexpr int BOXW = TV < 256 ? TV : 256;
    constexpr int NBOX = TV / BOXW;
    constexpr int STAGE_D = NJ * TV + (QSTREAM ? (NJ * KP + 15) / 16 * 16 : 0);  // doubles per stage; TMA dst needs 128 B
    static_assert(TV % BOXW == 0 && NJ % 2 == 0, "tile must be whole TMA boxes; Q rows must copy in 16-byte units");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *ring = reinterpret_cast<double *>(smem_raw);
    const int n = a.n;
    const int nit = (n + NJ - 1) / NJ;
    double *sQ = ring + (size_t)STAGES * STAGE_D;                    // resident Q' (only when !QSTREAM)
    SmemDesign<KP> &sd = *reinterpret_cast<SmemDesign<KP> *>(sQ + (QSTREAM ? 0 : (size_t)nit * NJ * KP));
    uint64_t *full = reinterpret_cast<uint64_t *>(&sd + 1);
    uint64_t *empty = full + STAGES;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;

    if (!QSTREAM)   // Qt is zero-padded to a multiple of 16 rows (capi.cu: make_qt), nit*NJ <= that
        for (int i = tid; i < nit * NJ * KP; i += blockDim.x) sQ[i] = a.Qt[i];
    load_design_to_smem<KP>(sd, a.design, tid, blockDim.x);
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CW);
        }
        fence_barrier_init();
    }
    __syncthreads();

    const int64_t ntiles = (a.V + TV - 1) / TV;
    constexpr uint32_t kStageBytes = (NJ * TV + (QSTREAM ? NJ * KP : 0)) * sizeof(double);

    if (warp == CW) {
        // ===== producer: one elected lane walks the same tile sequence as the consumers
        if ((tid & 31) == 0) {
            uint32_t stage = 0, phase = 0;
            for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
                if (flags && !flags[t]) continue;
                const int c0 = (int)(t * TV);
                for (int it = 0; it < nit; ++it) {
                    double *dst = ring + (size_t)stage * STAGE_D;
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], kStageBytes);
#pragma unroll
                    for (int b = 0; b < NBOX; ++b)
                        tma_load_2d(dst + (size_t)b * NJ * BOXW, &ymap, &full[stage], c0 + b * BOXW, it * NJ);
                    if (QSTREAM)
                        bulk_load_1d(dst + NJ * TV, a.Qt + (size_t)it * NJ * KP, NJ * KP * sizeof(double), &full[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        return;
    }

    // ===== consumers: thread owns voxels v0 + 2*tid, +1
    uint32_t stage = 0, phase = 0;
    const int lane = tid & 31;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t v = t * TV + 2 * tid;
        const bool in0 = v < a.V, in1 = v + 1 < a.V;
        bool act0 = in0, act1 = in1;
        if (a.mask) {
            act0 = in0 && mask_selects(a.mask[v], a.mask_lo, a.mask_hi);
            act1 = in1 && mask_selects(a.mask[v + 1], a.mask_lo, a.mask_hi);
        }
        double *o = a.out + v;
        if (flags && !flags[t]) {
            if (in0) zero_row<KP>(sd.ncol_out, o, a.ldOut);
            if (in1) zero_row<KP>(sd.ncol_out, o + 1, a.ldOut);
            continue;
        }
        VoxelSums<KP> s0, s1;
#pragma unroll
        for (int k = 0; k < KP; ++k) s0.e[k] = s1.e[k] = 0.0;
        s0.yy = s1.yy = 0.0;
        s0.y0 = s1.y0 = 0.0;
        for (int it = 0; it < nit; ++it) {
            mbar_wait(&full[stage], phase);
            const double *st = ring + (size_t)stage * STAGE_D;
            const double *tile = st + (size_t)((2 * tid) / BOXW) * NJ * BOXW + (2 * tid) % BOXW;
            if (it == 0 && sd.shift) {
                const double2 f = *reinterpret_cast<const double2 *>(tile);
                s0.y0 = f.x;
                s1.y0 = f.y;
            }
            const double *q = QSTREAM ? st + NJ * TV : sQ + (size_t)it * NJ * KP;
            // rows >= n of the last box are TMA zero-fill: their Q rows are zero, but their
            // (0 - y0)^2 must not enter yy, so the last box runs the guarded variant.
            if (it + 1 < nit) consume_box<KP, NJ, BOXW, false>(tile, q, s0, s1, NJ);
            else consume_box<KP, NJ, BOXW, true>(tile, q, s0, s1, n - it * NJ);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        // ---- fused epilogue
        double rss0 = s0.yy - sumsq_e(s0), mss0 = mss_from_e(s0, sd);
        double rss1 = s1.yy - sumsq_e(s1), mss1 = mss_from_e(s1, sd);
        if (act0 && rss0 <= kExactThreshold * s0.yy && s0.yy > 0.0) {
            const double2 ex = exact_pass<KP>(a.Y + v, a.ldY, n, a.Qt, s0.y0);
            rss0 = ex.x;
            mss0 = ex.y;
        }
        if (act1 && rss1 <= kExactThreshold * s1.yy && s1.yy > 0.0) {
            const double2 ex = exact_pass<KP>(a.Y + v + 1, a.ldY, n, a.Qt, s1.y0);
            rss1 = ex.x;
            mss1 = ex.y;
        }
        if (in1 && a.out_vec2) {
            finish_pair<KP>(s0, rss0, mss0, act0, s1, rss1, mss1, act1, sd, o, a.ldOut);
        } else {
            if (in0) {
                if (act0) finish_voxel<KP>(s0, rss0, mss0, sd, o, a.ldOut);
                else zero_row<KP>(sd.ncol_out, o, a.ldOut);
            }
            if (in1) {
                if (act1) finish_voxel<KP>(s1, rss1, mss1, sd, o + 1, a.ldOut);
                else zero_row<KP>(sd.ncol_out, o + 1, a.ldOut);
            }
        }
    }
}

// ------------------------------------------------------------------ LDG kernel (+ split-n)
// blockDim = (TX, S).  Thread (x, s) accumulates subjects [s*n/S, (s+1)*n/S) of voxels
// (blockIdx.x*TX + x)*VPT .. +VPT-1; slices are reduced through shared memory and slice 0
// runs the epilogue.  Shared memory: max(S x JT x KP Q rows, S x (KP+1) x VPT x TX partials).
// WGT: weighted least squares (voxel_wlm): every sample is multiplied by sqrt(w_j) (a.rowscale) after
// the shift; the staged tile row carries that factor as an extra KP-th value.
template <int KP, int VPT, bool WGT>
__global__ void __launch_bounds__(256)
lm_fit_ldg_kernel(LmArgs a, int JT)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ SmemDesign<KP> sd;
    constexpr int QW = KP + (WGT ? 1 : 0);   // doubles per staged row
    double *sQ = reinterpret_cast<double *>(smem_raw);
    const int TX = blockDim.x, S = blockDim.y;
    const int x = threadIdx.x, s = threadIdx.y;
    const int tid = s * TX + x, nthreads = TX * S;
    const int n = a.n;

    load_design_to_smem<KP>(sd, a.design, tid, nthreads);
    __syncthreads();

    int64_t v = ((int64_t)blockIdx.x * TX + x) * VPT;
    if (VPT == 2 && a.pair_list) {
        // compacted: slot -> pair index; slots past the count have nothing to do
        const int64_t slot = (int64_t)blockIdx.x * TX + x;
        v = slot < (int64_t)*a.n_pairs ? (int64_t)a.pair_list[slot] * 2 : a.V;
    }
    bool in[VPT], act[VPT];
    bool any = false;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        in[i] = v + i < a.V;
        act[i] = in[i] && (!a.mask || mask_selects(a.mask[v + i], a.mask_lo, a.mask_hi));
        any |= act[i];
    }
    VoxelSums<KP> sum[VPT];
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
#pragma unroll
        for (int k = 0; k < KP; ++k) sum[i].e[k] = 0.0;
        sum[i].yy = 0.0;
        sum[i].y0 = (sd.shift && act[i]) ? a.Y[v + i] : 0.0;
    }
    const int jb = (int)(((int64_t)s * n) / S), je = (int)(((int64_t)(s + 1) * n) / S);
    const int span = (n + S - 1) / S;  // longest slice
    const double *yp = a.Y + (any ? v : 0);
    if (a.pair_list && (int64_t)blockIdx.x * TX >= (int64_t)*a.n_pairs) return;   // whole block past the list

    for (int t0 = 0; t0 < span; t0 += JT) {
        // every slice stages rows [jb + t0, jb + t0 + JT) of Qt into its own sub-tile
        __syncthreads();
        for (int sl = 0; sl < S; ++sl) {
            const int b = (int)(((int64_t)sl * n) / S) + t0;
            const int e = (int)(((int64_t)(sl + 1) * n) / S);
            for (int i = tid; i < JT * QW; i += nthreads) {
                const int j = b + i / QW, c = i % QW;
                double val = 0.0;
                if (j < e) val = (c < KP) ? a.Qt[(size_t)j * KP + c] : a.rowscale[j];
                sQ[(size_t)sl * JT * QW + i] = val;
            }
        }
        __syncthreads();
        if (!any) continue;
        const double *q = sQ + (size_t)s * JT * QW;
        const int cnt = min(JT, je - (jb + t0));
        int jj = 0;
        constexpr int U = 8;
        for (; jj + U <= cnt; jj += U) {
            double y[U][VPT];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double *src = yp + (int64_t)(jb + t0 + jj + u) * a.ldY;
                if constexpr (VPT == 2) {
                    const double2 t = ldg_stream2_hint(src, a.l2_hint);
                    y[u][0] = t.x;
                    y[u][1] = t.y;
                } else {
                    y[u][0] = ldg_stream(src);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
#pragma unroll
                for (int i = 0; i < VPT; ++i) {
                    double yv = y[u][i] - sum[i].y0;
                    if (WGT) yv *= q[(jj + u) * QW + KP];
                    sum[i].yy = fma(yv, yv, sum[i].yy);
#pragma unroll
                    for (int k = 0; k < KP; ++k) sum[i].e[k] = fma(q[(jj + u) * QW + k], yv, sum[i].e[k]);
                }
            }
        }
        for (; jj < cnt; ++jj) {
            const double *src = yp + (int64_t)(jb + t0 + jj) * a.ldY;
#pragma unroll
            for (int i = 0; i < VPT; ++i) {
                // VPT == 2 is only launched when every pair is fully in range and aligned
                double yv = src[i] - sum[i].y0;
                if (WGT) yv *= q[jj * QW + KP];
                sum[i].yy = fma(yv, yv, sum[i].yy);
#pragma unroll
                for (int k = 0; k < KP; ++k) sum[i].e[k] = fma(q[jj * QW + k], yv, sum[i].e[k]);
            }
        }
    }

    if (S > 1) {
        // cross-slice reduction; partial layout [slice][component][voxel-in-block]
        __syncthreads();
        double *red = reinterpret_cast<double *>(smem_raw);
        const int W = TX * VPT;
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
#pragma unroll
            for (int k = 0; k < KP; ++k) red[((size_t)s * (KP + 1) + k) * W + x * VPT + i] = sum[i].e[k];
            red[((size_t)s * (KP + 1) + KP) * W + x * VPT + i] = sum[i].yy;
        }
        __syncthreads();
        if (s != 0) return;
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            for (int sl = 1; sl < S; ++sl) {
#pragma unroll
                for (int k = 0; k < KP; ++k) sum[i].e[k] += red[((size_t)sl * (KP + 1) + k) * W + x * VPT + i];
                sum[i].yy += red[((size_t)sl * (KP + 1) + KP) * W + x * VPT + i];
            }
        }
    }

#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        if (!in[i]) continue;
        double *o = a.out + v + i;
        if (!act[i]) {
            zero_row<KP>(sd.ncol_out, o, a.ldOut);
            continue;
        }
        double rss = sum[i].yy - sumsq_e(sum[i]), mss = mss_from_e(sum[i], sd);
        if (rss <= kExactThreshold * sum[i].yy && sum[i].yy > 0.0) {
            const double2 ex = exact_pass<KP>(a.Y + v + i, a.ldY, n, a.Qt, sum[i].y0, WGT ? a.rowscale : nullptr);
            rss = ex.x;
            mss = ex.y;
        }
        finish_voxel<KP>(sum[i], rss, mss, sd, o, a.ldOut);
    }
}

// ------------------------------------------------------------------ host-side launchers
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_tiled()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int round_up_kp(int p)
{
    if (p <= 8) return p < 1 ? 1 : p;
    if (p <= 12) return 12;
    if (p <= 16) return 16;
    if (p <= 24) return 24;
    return 32;
}

template <int KP, int CW, int NJ, int STAGES, bool QSTREAM>
size_t tma_smem_bytes(int n)
{
    const size_t stage = NJ * CW * 64 + (QSTREAM ? (NJ * KP + 15) / 16 * 16 : 0);
    const size_t qres = QSTREAM ? 0 : (size_t)((n + NJ - 1) / NJ) * NJ * KP;
    return (STAGES * stage + qres) * 8 + sizeof(SmemDesign<KP>) + 2 * STAGES * 8 + 128;
}

template <int KP, int CW, int NJ, int STAGES, int MINB, bool QSTREAM>
cudaError_t launch_tma_cfg(const LmArgs &a, const unsigned char *flags, int sm_count, cudaStream_t st)
{
    constexpr int TV = CW * 64;
    constexpr int BOXW = TV < 256 ? TV : 256;
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) return cudaErrorNotSupported;
    CUtensorMap map;
    const cuuint64_t dims[2] = {(cuuint64_t)a.V, (cuuint64_t)a.n};
    const cuuint64_t strides[1] = {(cuuint64_t)a.ldY * sizeof(double)};
    const cuuint32_t box[2] = {BOXW, NJ};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(a.Y), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    const size_t smem = tma_smem_bytes<KP, CW, NJ, STAGES, QSTREAM>(a.n);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    auto kern = lm_fit_tma_kernel<KP, CW, NJ, STAGES, MINB, QSTREAM>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int64_t ntiles = (a.V + TV - 1) / TV;
    const int64_t slots = (int64_t)sm_count * MINB;
    const int grid = (int)(ntiles < slots ? ntiles : slots);
    kern<<<grid, (CW + 1) * 32, smem, st>>>(map, a, flags);
    return cudaGetLastError();
}

template <int KP>
cudaError_t launch_kp(const LmArgs &a, const LmPlan &plan, LmScratch &scr, cudaStream_t st, LmLaunchInfo *info)
{
    const bool aligned16 = ((reinterpret_cast<uintptr_t>(a.Y) & 15) == 0) && (a.ldY % 2 == 0);
    // TMA needs a 16-byte aligned base and row pitch; int32 tile coordinates.
    const bool tma_ok = aligned16 && a.V <= 0x7fffffffLL && KP <= 16;
    // With a mask the direct-load kernel skips at warp (64-voxel) granularity where the TMA ring
    // skips whole 256-voxel tiles: measured 2.72 ms vs 3.02 ms on config 2 with the 38 % ellipsoid.
    bool use_tma = tma_ok && !a.rowscale && (plan.variant == LmVariant::Tma || (plan.variant == LmVariant::Auto && !a.mask));
    if (use_tma) {
        // Tile shapes (consumer warps, subjects per stage, ring depth, CTAs per SM):
        //   wide   4 x 64 = 256-voxel tiles, 4 CTAs/SM   -- large V (config 2: 4.5 ms, 98 % of the copy peak)
        //   narrow 2 x 64 = 128-voxel tiles, 8 CTAs/SM   -- small V: more, smaller tiles to balance 148 SMs
        //   cfg 0 (one 512-voxel CTA per SM) is kept for the record: too few warps to hide latency.
        int cfg = plan.tma_cfg;
        // wide tiles as soon as there are two of them per SM: 2 KB row segments use DRAM pages better than 1 KB ones
        // (config 3, 320 wide tiles on 148 SMs: 0.246 ms wide vs 0.285 ms narrow)
        if (cfg < 0) cfg = (a.V >= (int64_t)plan.sm_count * 2 * 256) ? 4 : 5;
        // ... and when all wide tiles are resident at once with 3 CTAs per SM (config 3: 321 tiles on 444 slots), every CTA
        // streams exactly one tile: twice the ring (8 subjects a stage) instead of a fourth CTA per SM that would have nothing
        // to do -- config 3: 0.231 ms against 0.240 ms (measured beside it: 8 x 6 stages x 2 CTAs/SM 0.276 ms -- 296 slots for
        // 321 tiles --, 4 x 8 stages x 3 CTAs/SM 0.244 ms)
        if (cfg == 4 && plan.tma_cfg < 0 && KP <= 8 && (a.V + 255) / 256 <= (int64_t)plan.sm_count * 3) cfg = 6;
        const int TV = cfg == 0 ? 512 : (cfg == 5 ? 128 : 256);
        const unsigned char *flags = nullptr;
        if (a.mask) {
            const int64_t ntiles = (a.V + TV - 1) / TV;
            cudaError_t e = scr.ensure_flags((size_t)ntiles, st);
            if (e != cudaSuccess) return e;
            mask_tile_flags_kernel<<<(unsigned)ntiles, 128, 0, st>>>(a.mask, a.V, TV, a.mask_lo, a.mask_hi, scr.flags, ntiles);
            if (info) info->launches++;
            flags = scr.flags;
        }
        cudaError_t e;
        // Q' resident in shared memory when 4 CTAs/SM still fit (n*KP*8 <= 24 KiB), else streamed
        const bool qres = plan.q_stream == 0 || (plan.q_stream < 0 && (size_t)a.n * KP * 8 <= 24 * 1024);
        if constexpr (KP <= 8) {
            switch (cfg * 2 + (qres ? 0 : 1)) {
            case 0: e = launch_tma_cfg<KP, 8, 8, 4, 1, false>(a, flags, plan.sm_count, st); break;
            case 1: e = launch_tma_cfg<KP, 8, 8, 4, 1, true>(a, flags, plan.sm_count, st); break;
            case 10: e = launch_tma_cfg<KP, 2, 4, 4, 8, false>(a, flags, plan.sm_count, st); break;
            case 11: e = launch_tma_cfg<KP, 2, 4, 4, 8, true>(a, flags, plan.sm_count, st); break;
            case 8: e = launch_tma_cfg<KP, 4, 4, 4, 4, false>(a, flags, plan.sm_count, st); break;
            case 12: case 13: e = launch_tma_cfg<KP, 4, 8, 4, 3, true>(a, flags, plan.sm_count, st); break;
            default: e = launch_tma_cfg<KP, 4, 4, 4, 4, true>(a, flags, plan.sm_count, st); break;
            }
        } else {
            // wide designs: more accumulators per thread -> fewer CTAs per SM
            e = cfg == 5 ? launch_tma_cfg<KP, 2, 4, 4, 4, true>(a, flags, plan.sm_count, st)
                         : launch_tma_cfg<KP, 4, 4, 4, 2, true>(a, flags, plan.sm_count, st);
        }
        if (info) { info->launches++; info->used_tma = 1; info->tma_cfg = cfg; }
        return e;
    }
    if (plan.variant == LmVariant::Tma && !a.rowscale) return cudaErrorInvalidValue;  // asked for TMA on a layout it cannot describe
    // ---- LDG path: pick voxels/thread and the split factor
    const bool vec2 = aligned16 && (a.V % 2 == 0) && (a.ldOut % 2 == 0) && KP <= 16 && !a.rowscale;
    int S = plan.split;
    if (S <= 0) {
        // enough threads for ~2 waves of 1024 threads/SM; split the subject axis otherwise
        S = 1;
        const int64_t want = (int64_t)plan.sm_count * 2048;
        while (S < 32 && (a.V / (vec2 ? 2 : 1)) * S < want && a.n / (S * 2) >= 32) S *= 2;
    }
    if (S > 32) S = 32;
    const int TX = S == 1 ? 128 : (256 / S < 8 ? 8 : 256 / S);
    const int VPT = vec2 ? 2 : 1;
    const int span = (a.n + S - 1) / S;
    int JT = span;
    const size_t budget = 64 * 1024;
    const int QW = KP + (a.rowscale ? 1 : 0);
    while ((size_t)S * JT * QW * 8 > budget) JT = (JT + 1) / 2;
    size_t smem = (size_t)S * JT * QW * 8;
    const size_t red = (size_t)S * (KP + 1) * VPT * TX * 8;
    if (S > 1 && red > smem) smem = red;
    const int64_t per_block = (int64_t)TX * VPT;
    const int64_t blocks = (a.V + per_block - 1) / per_block;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidValue;
    dim3 bd(TX, S);
    // ---- mask compaction: stream only the selected voxels (falls back to in-kernel skipping when the
    //      layout does not allow 16-byte pair accesses)
    LmArgs ac = a;
    const bool compact = a.mask && VPT == 2 && a.out_vec2 && (reinterpret_cast<uintptr_t>(a.mask) & 15) == 0 &&
                         a.V / 2 < 0xffffffffLL && !plan.no_compact;
    if (compact) {
        const int64_t npairs = a.V / 2;
        const int nb = (int)((npairs + kCompactBlock - 1) / kCompactBlock);
        const size_t off_list = ((size_t)nb * 4 + 4 + 255) & ~size_t(255);
        cudaError_t e = scr.ensure_bytes(off_list + (size_t)npairs * 4, st);
        if (e != cudaSuccess) return e;
        unsigned int *counts = reinterpret_cast<unsigned int *>(scr.flags);
        unsigned int *n_pairs = counts + nb;
        uint32_t *list = reinterpret_cast<uint32_t *>(scr.flags + off_list);
        mask_count_kernel<<<nb, kCompactBlock, 0, st>>>(a.mask, a.V, a.mask_lo, a.mask_hi, counts);
        mask_scan_kernel<<<1, 1024, 0, st>>>(counts, nb, n_pairs);
        mask_compact_kernel<<<nb, kCompactBlock, 0, st>>>(a.mask, a.V, a.mask_lo, a.mask_hi, counts, list, a.out, a.ldOut,
                                                           a.ncol_out);
        if (info) info->launches += 3;
        ac.pair_list = list;
        ac.n_pairs = n_pairs;
        // compacted runs: fetch no more than 64 bytes around what the loads touch (RMG_LDG_L2 = 0 / 64 / 128 / 256 overrides)
        static const int env_hint = [] { const char *s_ = std::getenv("RMG_LDG_L2"); return s_ ? std::atoi(s_) : -1; }();
        ac.l2_hint = env_hint >= 0 ? env_hint : 64;
    }
    auto go = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<(unsigned)blocks, bd, smem, st>>>(ac, JT);
        return cudaGetLastError();
    };
    cudaError_t e;
    constexpr int V2 = KP <= 16 ? 2 : 1;   // never instantiate the 2-voxel kernel for very wide designs
    if (a.rowscale) e = go(lm_fit_ldg_kernel<KP, 1, true>);       // weighted fits: anatLm-sized problems
    else if (VPT == 2) e = go(lm_fit_ldg_kernel<KP, V2, false>);
    else e = go(lm_fit_ldg_kernel<KP, 1, false>);
    if (info) { info->launches++; info->used_tma = 0; info->split = S; }
    return e;
}

}  // namespace

int lm_padded_p(int p) { return round_up_kp(p); }

cudaError_t launch_lm_fit(const LmArgs &a_in, const LmPlan &plan, LmScratch &scr, cudaStream_t st, LmLaunchInfo *info)
{
    if (a_in.V <= 0) return cudaSuccess;
    LmArgs a = a_in;
    a.out_vec2 = ((reinterpret_cast<uintptr_t>(a.out) & 15) == 0 && a.ldOut % 2 == 0) ? 1 : 0;
    if (a.ncol_out <= 0) a.ncol_out = 2 * a.p + 3;
    switch (round_up_kp(a.p)) {
#define RMG_CASE(K) case K: return launch_kp<K>(a, plan, scr, st, info);
        RMG_CASE(1) RMG_CASE(2) RMG_CASE(3) RMG_CASE(4) RMG_CASE(5) RMG_CASE(6) RMG_CASE(7) RMG_CASE(8)
        RMG_CASE(12) RMG_CASE(16) RMG_CASE(24) RMG_CASE(32)
#undef RMG_CASE
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace rmg
