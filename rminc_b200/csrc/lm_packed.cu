// lm_packed.cu -- the fit on volumes kept in their STORED type, sm_100a.
//
// Reference: minc2_model reads every slab with miget_real_value_hyperslab(..., MI_TYPE_DOUBLE, ...)
// (src/modelling_functions.c:1037-1055; src/minc_reader.c:294-331): libminc converts the stored voxels
// (usually unsigned 16-bit with a per-slice real range, or 32-bit float) to doubles on the CPU, and the
// doubles are what voxel_lm sees.  Shipping those doubles costs 8 bytes per sample over PCIe and HBM.
//
// Here the stored values travel as they are (2 or 4 bytes per sample) and the conversion
//     y = ((double)u - voxel_min) * scale[slice, subject] + offset[slice, subject]
// (three correctly rounded FP64 operations, no fused multiply-add: the doubles a C compiler produces from
// libminc's expression on a target without FMA contraction) happens in registers inside the fit kernel.  The
// accumulation and the epilogue are those of lm_fit_ldg_kernel, so the result equals the fit on the expanded
// doubles.  float32 volumes convert exactly and carry no scale.
//
// Roofline: HBM at 2n (u16) / 4n (f32) + 8(2p+3) bytes per voxel; at 2 bytes per sample the FP64 pipe
// ((5 + p) n operations per voxel) is the co-limiter.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <type_traits>

#include "async_ptx.cuh"
#include "lm_common.cuh"
#include "lm_launch.hpp"

namespace rmg {

namespace {

// Four consecutive stored samples -> doubles.  u16: 2^52 + u is exact, so (2^52 + u) - (2^52 + voxel_min) is the
// exactly rounded (double)u - voxel_min for integral voxel_min (checked on the host) in ONE FP64 add.
template <typename T>
struct Stored;

template <>
struct Stored<uint16_t> {
    static constexpr bool kScaled = true;
    using Raw = uint2;
    __device__ static __forceinline__ Raw load_raw(const uint16_t *p)
    {
        uint2 r;
        asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
        return r;
    }
    __device__ static __forceinline__ void convert(const Raw &r, double (&y)[4], double cmagic)
    {
        y[0] = __hiloint2double(0x43300000, (int)(r.x & 0xffffu)) - cmagic;
        y[1] = __hiloint2double(0x43300000, (int)(r.x >> 16)) - cmagic;
        y[2] = __hiloint2double(0x43300000, (int)(r.y & 0xffffu)) - cmagic;
        y[3] = __hiloint2double(0x43300000, (int)(r.y >> 16)) - cmagic;
    }
    __device__ static __forceinline__ double load1(const uint16_t *p, double cmagic)
    {
        return __hiloint2double(0x43300000, (int)__ldg(p)) - cmagic;
    }
};

template <>
struct Stored<float> {
    static constexpr bool kScaled = false;
    using Raw = float4;
    __device__ static __forceinline__ Raw load_raw(const float *p)
    {
        float4 r;
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
        return r;
    }
    __device__ static __forceinline__ void convert(const Raw &r, double (&y)[4], double)
    {
        y[0] = (double)r.x; y[1] = (double)r.y; y[2] = (double)r.z; y[3] = (double)r.w;
    }
    __device__ static __forceinline__ double load1(const float *p, double) { return (double)__ldg(p); }
};

// VPT (2 or 4) adjacent samples out of shared memory -> doubles (after the voxel_min subtraction for uint16)
template <typename T, int VPT>
__device__ __forceinline__ void load_convert_smem(const T *p, double (&y)[VPT], double cmagic)
{
    if constexpr (sizeof(T) == 2) {
        uint32_t w[VPT / 2];
        if constexpr (VPT == 8) {
            const uint4 r = *reinterpret_cast<const uint4 *>(p);
            w[0] = r.x; w[1] = r.y; w[2] = r.z; w[3] = r.w;
        } else if constexpr (VPT == 4) {
            const uint2 r = *reinterpret_cast<const uint2 *>(p);
            w[0] = r.x; w[1] = r.y;
        } else {
            w[0] = *reinterpret_cast<const uint32_t *>(p);
        }
#pragma unroll
        for (int i = 0; i < VPT / 2; ++i) {
            y[2 * i] = __hiloint2double(0x43300000, (int)(w[i] & 0xffffu)) - cmagic;
            y[2 * i + 1] = __hiloint2double(0x43300000, (int)(w[i] >> 16)) - cmagic;
        }
    } else {
        if constexpr (VPT == 8) {
            const float4 r = *reinterpret_cast<const float4 *>(p), q = *reinterpret_cast<const float4 *>(p + 4);
            y[0] = (double)r.x; y[1] = (double)r.y; y[2] = (double)r.z; y[3] = (double)r.w;
            y[4] = (double)q.x; y[5] = (double)q.y; y[6] = (double)q.z; y[7] = (double)q.w;
        } else if constexpr (VPT == 4) {
            const float4 r = *reinterpret_cast<const float4 *>(p);
            y[0] = (double)r.x; y[1] = (double)r.y; y[2] = (double)r.z; y[3] = (double)r.w;
        } else {
            const float2 r = *reinterpret_cast<const float2 *>(p);
            y[0] = (double)r.x; y[1] = (double)r.y;
        }
    }
}

// FOLD form of the uint16 conversion (TMA variant, default): the sample becomes w = 1 + u * 2^-16 by placing u in the
// top 16 mantissa bits of a double -- integer work only -- and its value is ONE fused multiply-add
//     y = A * w + C,   A = scale * 2^16 (exact),   C = fl(offset - scale * (voxel_min + 2^16)) (one rounding, host fma).
// A * w + C is evaluated exactly and rounded once; with the rounding of C the result lies within
// 2^-53 * |C| + ulp(y) / 2 of the real number the reader's expression denotes -- the same size as the reader's own two
// roundings -- but it is not always the reader's double bit for bit (RMG_STORED_EXACT=1 selects the three-rounding form).
template <int VPT>
__device__ __forceinline__ void load_unit_smem(const uint16_t *p, double (&w)[VPT])
{
    uint32_t r[VPT / 2];
    if constexpr (VPT == 8) {
        const uint4 q = *reinterpret_cast<const uint4 *>(p);
        r[0] = q.x; r[1] = q.y; r[2] = q.z; r[3] = q.w;
    } else if constexpr (VPT == 4) {
        const uint2 q = *reinterpret_cast<const uint2 *>(p);
        r[0] = q.x; r[1] = q.y;
    } else {
        r[0] = *reinterpret_cast<const uint32_t *>(p);
    }
#pragma unroll
    for (int i = 0; i < VPT / 2; ++i) {
        // (x & 0xffff0) | 0x3ff00000 as ONE three-input logic op (the compiler splits the two immediates into two)
        uint32_t h0, h1;
        asm("lop3.b32 %0, %1, 0x000ffff0, 0x3ff00000, 0xEA;" : "=r"(h0) : "r"(r[i] << 4));
        asm("lop3.b32 %0, %1, 0x000ffff0, 0x3ff00000, 0xEA;" : "=r"(h1) : "r"(r[i] >> 12));
        w[2 * i] = __hiloint2double((int)h0, 0);
        w[2 * i + 1] = __hiloint2double((int)h1, 0);
    }
}

// FOLD = 2 (default): the sample sits in the LOW 16 bits of the high word, w = 1 + u * 2^-20 -- a byte-aligned position, so
// each high word is ONE byte permute of the packed pair and the constant 0x3ff00000 (no shift) -- with A = scale * 2^20 and
// C = fl(offset - scale * (voxel_min + 2^20)).  |C| grows to 16 value ranges, its rounding to 2^-49 of the range: far below
// the 2^-16 of the range one stored count stands for.
template <int VPT>
__device__ __forceinline__ void load_unit20_smem(const uint16_t *p, double (&w)[VPT], uint32_t kexp)
{
    uint32_t r[VPT / 2];
    if constexpr (VPT == 8) {
        const uint4 q = *reinterpret_cast<const uint4 *>(p);
        r[0] = q.x; r[1] = q.y; r[2] = q.z; r[3] = q.w;
    } else if constexpr (VPT == 4) {
        const uint2 q = *reinterpret_cast<const uint2 *>(p);
        r[0] = q.x; r[1] = q.y;
    } else {
        r[0] = *reinterpret_cast<const uint32_t *>(p);
    }
    // Only the HIGH words of w[] are replaced: the low words (zero, but a kernel parameter the compiler cannot fold) stay
    // where they are, so a sample costs one logic op -- no shift, no move -- before its fused multiply-add.
#pragma unroll
    for (int i = 0; i < VPT / 2; ++i) {
        uint32_t h0, h1;
        asm("lop3.b32 %0, %1, 0xffff, %2, 0xEA;" : "=r"(h0) : "r"(r[i]), "r"(kexp));       // (r & 0xffff) | 0x3ff00000
        asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(h1) : "r"(r[i]), "r"(kexp));             // (r >> 16)    | 0x3ff00000
        asm("{ .reg .b32 lo, hi; mov.b64 {lo, hi}, %0; mov.b64 %0, {lo, %1}; }" : "+d"(w[2 * i]) : "r"(h0));
        asm("{ .reg .b32 lo, hi; mov.b64 {lo, hi}, %0; mov.b64 %0, {lo, %1}; }" : "+d"(w[2 * i + 1]) : "r"(h1));
    }
}

// value of one sample as the reference's doubles: separate multiply and add (never an FMA)
__device__ __forceinline__ double descale(double t, double sc, double of) { return __dadd_rn(__dmul_rn(t, sc), of); }

// Stored samples -> the doubles of the reference's reader, written to a resident V x n matrix (the permutation test
// re-reads Y for every batch of permutations, so it expands once instead of in every pass).  One sample per thread:
// the kernel moves 2-4 + 8 bytes per sample once per call and is nowhere near the critical path.
template <typename T>
__global__ void __launch_bounds__(256)
expand_stored_kernel(const T *__restrict__ U, int64_t V, int64_t ldU, const double *__restrict__ scale,
                     const double *__restrict__ offset, double cmagic, int64_t slice_len, int64_t nslices, int64_t v_base,
                     double *__restrict__ Y, int64_t ldY)
{
    const int64_t v = (int64_t)blockIdx.x * 256 + threadIdx.x;
    const int j = blockIdx.y;
    if (v >= V) return;
    double t = Stored<T>::load1(U + (int64_t)j * ldU + v, cmagic);
    if (Stored<T>::kScaled) {
        const int64_t sl = min((v_base + v) / slice_len, nslices - 1);
        t = descale(t, __ldg(scale + sl + (int64_t)j * nslices), __ldg(offset + sl + (int64_t)j * nslices));
    }
    Y[(int64_t)j * ldY + v] = t;
}

template <int KP, typename T>
__device__ __noinline__ double2 exact_pass_packed(const T *__restrict__ u, int64_t ld, int n, const double *__restrict__ Qt,
                                                  double y0, double cmagic, const double *__restrict__ sc,
                                                  const double *__restrict__ of, int64_t sstride)
{
    auto sample = [&](int j) {
        double t = Stored<T>::load1(u + (int64_t)j * ld, cmagic);
        if (Stored<T>::kScaled) t = descale(t, __ldg(sc + (int64_t)j * sstride), __ldg(of + (int64_t)j * sstride));
        return t - y0;
    };
    double e[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) e[k] = 0.0;
    for (int j = 0; j < n; ++j) {
        const double yv = sample(j);
#pragma unroll
        for (int k = 0; k < KP; ++k) e[k] = fma(__ldg(Qt + (size_t)j * KP + k), yv, e[k]);
    }
    double r2 = 0.0, sf = 0.0;
    for (int j = 0; j < n; ++j) {
        double fit = 0.0;
#pragma unroll
        for (int k = 0; k < KP; ++k) fit = fma(__ldg(Qt + (size_t)j * KP + k), e[k], fit);
        const double r = sample(j) - fit;
        r2 = fma(r, r, r2);
        sf += fit;
    }
    const double mean = sf / n;
    double m2 = 0.0;
    for (int j = 0; j < n; ++j) {
        double fit = 0.0;
#pragma unroll
        for (int k = 0; k < KP; ++k) fit = fma(__ldg(Qt + (size_t)j * KP + k), e[k], fit);
        const double dlt = fit - mean;
        m2 = fma(dlt, dlt, m2);
    }
    return make_double2(r2, m2);
}

// The same recomputation for the TMA variant, result row included.  A call in the middle of the epilogue makes the
// compiler park the epilogue's live values on the stack for EVERY tile (84-104 bytes per thread and tile: ~90 MB of extra
// DRAM writes per config-2 launch); this one is called after the regular rows are stored, takes scalars and pointers
// only, and overwrites the voxel's row.
template <int KP, typename T>
__device__ __noinline__ void exact_voxel_packed(const T *__restrict__ u, int64_t ld, int n, const double *__restrict__ Qt,
                                                double y0, double cmagic, const double *__restrict__ sc,
                                                const double *__restrict__ of, int64_t sstride,
                                                const SmemDesign<KP> *__restrict__ sd, double *__restrict__ out, int64_t ldOut)
{
    auto sample = [&](int j) {
        double t = Stored<T>::load1(u + (int64_t)j * ld, cmagic);
        if (Stored<T>::kScaled) t = descale(t, __ldg(sc + (int64_t)j * sstride), __ldg(of + (int64_t)j * sstride));
        return t - y0;
    };
    VoxelSums<KP> s;
#pragma unroll
    for (int k = 0; k < KP; ++k) s.e[k] = 0.0;
    s.y0 = y0;
    s.yy = 0.0;
    for (int j = 0; j < n; ++j) {
        const double yv = sample(j);
#pragma unroll
        for (int k = 0; k < KP; ++k) s.e[k] = fma(__ldg(Qt + (size_t)j * KP + k), yv, s.e[k]);
    }
    double r2 = 0.0, sf = 0.0;
    for (int j = 0; j < n; ++j) {
        double fit = 0.0;
#pragma unroll
        for (int k = 0; k < KP; ++k) fit = fma(__ldg(Qt + (size_t)j * KP + k), s.e[k], fit);
        const double r = sample(j) - fit;
        r2 = fma(r, r, r2);
        sf += fit;
    }
    const double mean = sf / n;
    double m2 = 0.0;
    for (int j = 0; j < n; ++j) {
        double fit = 0.0;
#pragma unroll
        for (int k = 0; k < KP; ++k) fit = fma(__ldg(Qt + (size_t)j * KP + k), s.e[k], fit);
        const double dlt = fit - mean;
        m2 = fma(dlt, dlt, m2);
    }
    finish_voxel<KP>(s, r2, m2, *sd, out, ldOut);
}

}  // namespace

// One thread = VPT (4 or 1) adjacent voxels, all n subjects.  Shared memory holds, JT subjects at a time, one row
// [Q'(j, 0..KP) | scale_a(j) offset_a(j) | scale_b(j) offset_b(j)] per subject, where a / b are the (at most two)
// slices the block's voxels lie in: no global load besides the samples themselves remains in the inner loop.  The
// next UNR subjects' raw samples are prefetched into registers while the current ones are converted and accumulated.
template <int KP, typename T, int VPT>
__global__ void __launch_bounds__(128, VPT == 1 ? 6 : (KP <= 4 ? 4 : (KP <= 6 ? 3 : (KP <= 8 ? 2 : 1))))
lm_fit_packed_kernel(PackedArgs a, int JT)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ SmemDesign<KP> sd;
    constexpr bool SC = Stored<T>::kScaled;
    constexpr int TS = (KP + 1) & ~1;                   // first table column (even: 16-byte aligned pairs)
    constexpr int QW = TS + (SC ? 4 : 0);               // doubles per staged row
    double *sQ = reinterpret_cast<double *>(smem_raw);
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int n = a.n;
    const T *U = static_cast<const T *>(a.U);

    load_design_to_smem<KP>(sd, a.design, tid, nthreads);
    __syncthreads();

    const int64_t vb = (int64_t)blockIdx.x * nthreads * VPT;     // first voxel of the block
    const int64_t v = vb + (int64_t)tid * VPT;
    bool in[VPT], act[VPT];
    bool any = false;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        in[i] = v + i < a.V;
        act[i] = in[i] && (!a.mask || mask_selects(a.mask[v + i], a.mask_lo, a.mask_hi));
        any |= act[i];
    }
    // slices: the block's first slice is table a, the next one table b; a block reaching a third slice (tiny slices)
    // or a thread whose voxels straddle a boundary reads the tables from global memory instead
    const int64_t sl_a = SC ? min((a.v_base + vb) / a.slice_len, a.nslices - 1) : 0;
    const int64_t sl_b = min(sl_a + 1, a.nslices - 1);
    int64_t sl[VPT];
    bool same = true;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        sl[i] = SC ? min((a.v_base + v + i) / a.slice_len, a.nslices - 1) : 0;
        same &= sl[i] == sl[0];
    }
    const bool fast = !SC || (same && sl[0] <= sl_b);
    const int tsel = TS + ((SC && sl[0] != sl_a) ? 2 : 0);       // column of this thread's (scale, offset) pair
    const double cm = a.cmagic;
    VoxelSums<KP> sum[VPT];
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
#pragma unroll
        for (int k = 0; k < KP; ++k) sum[i].e[k] = 0.0;
        sum[i].yy = 0.0;
        sum[i].y0 = 0.0;
        if (sd.shift && act[i]) {
            double t = Stored<T>::load1(U + v + i, cm);
            if (SC) t = descale(t, __ldg(a.scale + sl[i]), __ldg(a.offset + sl[i]));
            sum[i].y0 = t;
        }
    }
    const T *up = U + (any ? v : 0);
    constexpr int UNR = 8;

    for (int t0 = 0; t0 < n; t0 += JT) {
        __syncthreads();
        const int cnt = min(JT, n - t0);
        for (int i = tid; i < cnt * QW; i += nthreads) {
            const int j = t0 + i / QW, c = i % QW;
            double val = 0.0;
            if (c < KP) val = a.Qt[(size_t)j * KP + c];
            else if (SC && c >= TS) {
                const int64_t s_ = (c - TS) < 2 ? sl_a : sl_b;
                val = ((c - TS) & 1) ? a.offset[s_ + (int64_t)j * a.nslices] : a.scale[s_ + (int64_t)j * a.nslices];
            }
            sQ[i] = val;
        }
        __syncthreads();
        if (!any) continue;

        // FAST: this thread's (scale, offset) pair comes from the staged row; otherwise (a thread straddling a slice
        // boundary, or tiny slices) from global memory.  Two separately compiled loops keep the common one lean.
        auto run = [&](auto fast_tag) {
            constexpr bool FAST = decltype(fast_tag)::value;
            auto accumulate = [&](int jj, int i, double t) {
                const double *row = sQ + jj * QW;
                if (SC) {
                    double sc, of;
                    if (FAST) {
                        const double2 so = *reinterpret_cast<const double2 *>(row + tsel);
                        sc = so.x;
                        of = so.y;
                    } else {
                        sc = __ldg(a.scale + sl[i] + (int64_t)(t0 + jj) * a.nslices);
                        of = __ldg(a.offset + sl[i] + (int64_t)(t0 + jj) * a.nslices);
                    }
                    t = descale(t, sc, of);
                }
                const double yv = t - sum[i].y0;
                sum[i].yy = fma(yv, yv, sum[i].yy);
#pragma unroll
                for (int k = 0; k < KP; ++k) sum[i].e[k] = fma(row[k], yv, sum[i].e[k]);
            };
            int jj = 0;
            if constexpr (VPT == 4) {
                typename Stored<T>::Raw cur[UNR], nxt[UNR];
                if (cnt >= UNR) {
#pragma unroll
                    for (int u = 0; u < UNR; ++u) nxt[u] = Stored<T>::load_raw(up + (int64_t)(t0 + u) * a.ldU);
                }
                for (; jj + UNR <= cnt; jj += UNR) {
#pragma unroll
                    for (int u = 0; u < UNR; ++u) cur[u] = nxt[u];
                    if (jj + 2 * UNR <= cnt) {
#pragma unroll
                        for (int u = 0; u < UNR; ++u)
                            nxt[u] = Stored<T>::load_raw(up + (int64_t)(t0 + jj + UNR + u) * a.ldU);
                    }
#pragma unroll
                    for (int u = 0; u < UNR; ++u) {
                        double y[4];
                        Stored<T>::convert(cur[u], y, cm);
#pragma unroll
                        for (int i = 0; i < 4; ++i) accumulate(jj + u, i, y[i]);
                    }
                }
            } else {
                for (; jj + UNR <= cnt; jj += UNR) {
                    double y1[UNR];
#pragma unroll
                    for (int u = 0; u < UNR; ++u) y1[u] = Stored<T>::load1(up + (int64_t)(t0 + jj + u) * a.ldU, cm);
#pragma unroll
                    for (int u = 0; u < UNR; ++u) accumulate(jj + u, 0, y1[u]);
                }
            }
            for (; jj < cnt; ++jj) {
                const T *src = up + (int64_t)(t0 + jj) * a.ldU;
#pragma unroll
                for (int i = 0; i < VPT; ++i) accumulate(jj, i, Stored<T>::load1(src + i, cm));
            }
        };
        if (fast) run(std::true_type{});
        else run(std::false_type{});
    }

    // ---- fused epilogue (as lm_fit_ldg_kernel)
    double rss[VPT], mss[VPT];
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        rss[i] = sum[i].yy - sumsq_e(sum[i]);
        mss[i] = mss_from_e(sum[i], sd);
        if (act[i] && rss[i] <= kExactThreshold * sum[i].yy && sum[i].yy > 0.0) {
            const double2 ex = exact_pass_packed<KP, T>(U + v + i, a.ldU, n, a.Qt, sum[i].y0, cm, a.scale + sl[i],
                                                        a.offset + sl[i], a.nslices);
            rss[i] = ex.x;
            mss[i] = ex.y;
        }
    }
    if (VPT == 4 && a.out_vec2 && in[VPT - 1]) {
#pragma unroll
        for (int i = 0; i + 1 < VPT; i += 2)
            finish_pair<KP>(sum[i], rss[i], mss[i], act[i], sum[i + 1], rss[i + 1], mss[i + 1], act[i + 1], sd,
                            a.out + v + i, a.ldOut);
        return;
    }
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        if (!in[i]) continue;
        double *o = a.out + v + i;
        if (act[i]) finish_voxel<KP>(sum[i], rss[i], mss[i], sd, o, a.ldOut);
        else zero_row<KP>(sd.ncol_out, o, a.ldOut);
    }
}


// ------------------------------------------------------------------ masked fits on stored volumes: compacted quads
// The everyday call -- uint16 files AND a mask -- used to cost as much as the unmasked fit: the direct-load kernel above
// visits every voxel and only skips the loads of masked ones.  Here three small kernels turn the mask into an ascending list of
// voxel QUADS (4 adjacent voxels = one 8-byte load of uint16 samples) that contain a selected voxel and write the +0.0 rows of all
// other quads (the double path's pair compaction, lm_kernels.cu); the fit kernel then gives every thread one LISTED quad, so
// all lanes of the FP64-bound loop do useful work, and its loads carry a 64-byte L2 prefetch size (runs of selected voxels are
// ~190 bytes per subject: the default 128 bytes around each load would fetch half as much again).  Measured and dropped: a
// per-thread ring of asynchronous 8-byte copies (cp.async, 32 subjects deep) instead of the register prefetch -- 1.55 ms
// against 1.27 ms on the 38 % ellipsoid of config 2 (one LDGSTS + one LDS per quad and subject cost more than they hide).  Arithmetic as the TMA
// variant: w = 1 + u 2^-20 and one fused multiply-add per sample, ONE shift per quad folded into the offsets (FOLD), or the
// reader's three roundings (RMG_STORED_EXACT=1).
constexpr int kQuadBlock = 1024;   // quads per compaction block

__device__ __forceinline__ bool quad_selected(const double *__restrict__ mask, int64_t q, int64_t V, double lo, double hi)
{
    const int64_t v = q * 4;
    bool sel = false;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (v + i < V) sel |= mask_selects(__ldg(mask + v + i), lo, hi);
    return sel;
}

__global__ void __launch_bounds__(kQuadBlock)
quad_count_kernel(const double *__restrict__ mask, int64_t V, double lo, double hi, unsigned int *__restrict__ counts)
{
    const int64_t q = (int64_t)blockIdx.x * kQuadBlock + threadIdx.x;
    const int c = __syncthreads_count(quad_selected(mask, q, V, lo, hi));
    if (threadIdx.x == 0) counts[blockIdx.x] = (unsigned int)c;
}

// exclusive scan of the per-block counts by one block; total -> *n_quads
__global__ void __launch_bounds__(1024)
quad_scan_kernel(unsigned int *__restrict__ counts, int nblocks, unsigned int *__restrict__ n_quads)
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
            warp_sums[threadIdx.x] = wi - w;
        }
        __syncthreads();
        const unsigned int excl = carry + warp_sums[threadIdx.x >> 5] + incl - x;
        if (i < nblocks) counts[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_quads = carry;
}

__global__ void __launch_bounds__(kQuadBlock)
quad_compact_kernel(const double *__restrict__ mask, int64_t V, double lo, double hi, const unsigned int *__restrict__ offsets,
                    uint32_t *__restrict__ list, double *__restrict__ out, int64_t ldOut, int ncol, int vec2)
{
    __shared__ unsigned int warp_sums[32];
    const int64_t q = (int64_t)blockIdx.x * kQuadBlock + threadIdx.x;
    const bool sel = quad_selected(mask, q, V, lo, hi);
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
        list[offsets[blockIdx.x] + warp_sums[warp] + __popc(ballot & ((1u << lane) - 1u))] = (uint32_t)q;
    } else if (q * 4 < V) {
        // V % 4 == 0 (host): the quad is whole; with out 16-byte aligned and ldOut even its rows take two 16-byte stores a column
        if (vec2) {
            double2 *o = reinterpret_cast<double2 *>(out + q * 4);
            const int64_t ld2 = ldOut / 2;
            const double2 z = make_double2(0.0, 0.0);
#pragma unroll 4
            for (int c = 0; c < ncol; ++c) {
                o[(int64_t)c * ld2] = z;
                o[(int64_t)c * ld2 + 1] = z;
            }
        } else {
            for (int c = 0; c < ncol; ++c)
#pragma unroll
                for (int i = 0; i < 4; ++i) out[q * 4 + i + (int64_t)c * ldOut] = 0.0;
        }
    }
}

template <typename T>
__device__ __forceinline__ typename Stored<T>::Raw load_raw_64b(const T *p);
template <>
__device__ __forceinline__ uint2 load_raw_64b<uint16_t>(const uint16_t *p)
{
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
template <>
__device__ __forceinline__ float4 load_raw_64b<float>(const float *p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

// the four packed uint16 samples of a quad as w = 1 + u 2^-20 (see load_unit20_smem)
__device__ __forceinline__ void unit20_from_raw(const uint2 &r, double (&w)[4])
{
    w[0] = __hiloint2double((int)__byte_perm(r.x, 0x3ff00000u, 0x7610), 0);
    w[1] = __hiloint2double((int)__byte_perm(r.x, 0x3ff00000u, 0x7632), 0);
    w[2] = __hiloint2double((int)__byte_perm(r.y, 0x3ff00000u, 0x7610), 0);
    w[3] = __hiloint2double((int)__byte_perm(r.y, 0x3ff00000u, 0x7632), 0);
}

template <int KP, typename T, bool FOLD>
__global__ void __launch_bounds__(128, KP <= 4 ? 4 : (KP <= 6 ? 3 : (KP <= 8 ? 2 : 1)))
lm_fit_packed_quads_kernel(PackedArgs a, const uint32_t *__restrict__ list, const unsigned int *__restrict__ n_quads, int JT)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ SmemDesign<KP> sd;
    constexpr bool SC = Stored<T>::kScaled;
    constexpr bool FD = SC && FOLD;
    constexpr int TS = (KP + 1) & ~1;
    constexpr int QW = TS + (SC ? 4 : 0);
    constexpr double kOne = 1048576.0;                  // the value of w's leading 1 in stored counts (2^20)
    const unsigned int nq = *n_quads;
    if ((int64_t)blockIdx.x * 128 >= (int64_t)nq) return;      // whole block past the list
    double *sQ = reinterpret_cast<double *>(smem_raw);
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int n = a.n;
    const T *U = static_cast<const T *>(a.U);
    load_design_to_smem<KP>(sd, a.design, tid, nthreads);
    __syncthreads();

    const int64_t slot = (int64_t)blockIdx.x * 128 + tid;
    const bool have = slot < (int64_t)nq;
    const int64_t v = have ? (int64_t)list[slot] * 4 : 0;
    const int64_t vb = (int64_t)list[(int64_t)blockIdx.x * 128] * 4;     // first voxel of the block's first quad
    bool act[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) act[i] = have && mask_selects(__ldg(a.mask + v + i), a.mask_lo, a.mask_hi);   // V % 4 == 0 (host)
    const int64_t sl_a = SC ? min((a.v_base + vb) / a.slice_len, a.nslices - 1) : 0;
    const int64_t sl_b = min(sl_a + 1, a.nslices - 1);
    int64_t sl[4];
    bool same = true;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        sl[i] = SC ? min((a.v_base + v + i) / a.slice_len, a.nslices - 1) : 0;
        same &= sl[i] == sl[0];
    }
    const bool fast = !SC || (same && sl[0] <= sl_b);
    const int tsel = TS + ((SC && sl[0] != sl_a) ? 2 : 0);
    const double cm = a.cmagic;
    const double vmin1 = (cm - 4503599627370496.0) + kOne;      // voxel_min + 2^20, an exact integer
    VoxelSums<KP> sum[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int k = 0; k < KP; ++k) sum[i].e[k] = 0.0;
        sum[i].yy = 0.0;
        sum[i].y0 = 0.0;
    }
    // ONE shift per quad: the value of its first voxel in the first subject -- unless a voxel of the quad lies far from it
    // next to its spread over the first subjects (an edge inside the quad: see the TMA variant); then the WARP takes the
    // per-voxel form of the loop (one more subtraction per sample), so that such voxels do not all end in the explicit-residual
    // path.
    const T *up = U + v;
    double shift = 0.0, dlt[4] = {0.0, 0.0, 0.0, 0.0};
    bool far = false;
    // Every value below is formed EXACTLY as the loop forms it, so that a voxel with one value in every subject gives y' = 0 bit
    // for bit (its rss = 0 classes, without the exact path).  The three-rounding and float32 forms subtract per sample anyway:
    // they always take per-voxel shifts (shift = 0, dlt[i] = the voxel's first sample).
    if (sd.shift && have) {
        const int nv0 = min(8, n);
        double var[4] = {0.0, 0.0, 0.0, 0.0};
        for (int j = 0; j < nv0; ++j) {
            const typename Stored<T>::Raw raw = load_raw_64b<T>(up + (int64_t)j * a.ldU);
            double y[4];
            if constexpr (FD) unit20_from_raw(raw, y);
            else Stored<T>::convert(raw, y, cm);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                double t = y[i];
                if constexpr (FD) {
                    const double sc = __ldg(a.scale + sl[i] + (int64_t)j * a.nslices), of = __ldg(a.offset + sl[i] + (int64_t)j * a.nslices);
                    const double cc = fma(-sc, vmin1, of);
                    if (j == 0 && i == 0) shift = fma(sc * kOne, t, cc);
                    t = fma(sc * kOne, t, cc - shift);
                } else if (SC) {
                    t = descale(t, __ldg(a.scale + sl[i] + (int64_t)j * a.nslices), __ldg(a.offset + sl[i] + (int64_t)j * a.nslices));
                }
                if (j == 0) dlt[i] = t;
                const double d = t - dlt[i];
                var[i] = fma(d, d, var[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) far |= dlt[i] * dlt[i] * (double)nv0 > 50.0 * var[i];
    }
    const bool per_voxel = !FD || a.shift_mode == 2 || (a.shift_mode == 0 && __any_sync(0xffffffffu, far));
#pragma unroll
    for (int i = 0; i < 4; ++i) dlt[i] = per_voxel ? dlt[i] : 0.0;         // voxel i's shift is shift + dlt[i] from here on
    constexpr int UNR = 8;

    for (int t0 = 0; t0 < n; t0 += JT) {
        __syncthreads();
        const int cnt = min(JT, n - t0);
        for (int i = tid; i < cnt * QW; i += nthreads) {
            const int j = t0 + i / QW, c = i % QW;
            double val = 0.0;
            if (c < KP) val = a.Qt[(size_t)j * KP + c];
            else if (SC && c >= TS) {
                const int64_t s_ = (c - TS) < 2 ? sl_a : sl_b;
                const double sc = a.scale[s_ + (int64_t)j * a.nslices], of = a.offset[s_ + (int64_t)j * a.nslices];
                if (FD) val = ((c - TS) & 1) ? fma(-sc, vmin1, of) : sc * kOne;      // (A, C) of the one-FMA form
                else val = ((c - TS) & 1) ? of : sc;
            }
            sQ[i] = val;
        }
        __syncthreads();
        if (!have) continue;

        auto run = [&](auto fast_tag, auto pv_tag) {
            constexpr bool FAST = decltype(fast_tag)::value;
            constexpr bool PV = decltype(pv_tag)::value;
            auto accumulate = [&](int jj, const typename Stored<T>::Raw &raw) {
                const double *row = sQ + jj * QW;
                double yv[4];
                if constexpr (FD) {
                    double w[4];
                    unit20_from_raw(raw, w);
                    if (FAST) {
                        const double2 so = *reinterpret_cast<const double2 *>(row + tsel);
                        const double cs = so.y - shift;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            yv[i] = fma(so.x, w[i], cs);
                            if (PV) yv[i] -= dlt[i];
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const double sc = __ldg(a.scale + sl[i] + (int64_t)(t0 + jj) * a.nslices);
                            const double of = __ldg(a.offset + sl[i] + (int64_t)(t0 + jj) * a.nslices);
                            yv[i] = fma(sc * kOne, w[i], fma(-sc, vmin1, of) - shift) - dlt[i];
                        }
                    }
                } else {
                    double y[4];
                    Stored<T>::convert(raw, y, cm);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        double t = y[i];
                        if (SC) {
                            double sc, of;
                            if (FAST) {
                                const double2 so = *reinterpret_cast<const double2 *>(row + tsel);
                                sc = so.x;
                                of = so.y;
                            } else {
                                sc = __ldg(a.scale + sl[i] + (int64_t)(t0 + jj) * a.nslices);
                                of = __ldg(a.offset + sl[i] + (int64_t)(t0 + jj) * a.nslices);
                            }
                            t = descale(t, sc, of);
                        }
                        yv[i] = t - (shift + dlt[i]);
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    sum[i].yy = fma(yv[i], yv[i], sum[i].yy);
#pragma unroll
                    for (int k = 0; k < KP; ++k) sum[i].e[k] = fma(row[k], yv[i], sum[i].e[k]);
                }
            };
            int jj = 0;
            typename Stored<T>::Raw cur[UNR], nxt[UNR];
            if (cnt >= UNR) {
#pragma unroll
                for (int u = 0; u < UNR; ++u) nxt[u] = load_raw_64b<T>(up + (int64_t)(t0 + u) * a.ldU);
            }
            for (; jj + UNR <= cnt; jj += UNR) {
#pragma unroll
                for (int u = 0; u < UNR; ++u) cur[u] = nxt[u];
                if (jj + 2 * UNR <= cnt) {
#pragma unroll
                    for (int u = 0; u < UNR; ++u) nxt[u] = load_raw_64b<T>(up + (int64_t)(t0 + jj + UNR + u) * a.ldU);
                }
#pragma unroll
                for (int u = 0; u < UNR; ++u) accumulate(jj + u, cur[u]);
            }
            for (; jj < cnt; ++jj) accumulate(jj, load_raw_64b<T>(up + (int64_t)(t0 + jj) * a.ldU));
        };
        if (fast && !(FD && per_voxel)) run(std::true_type{}, std::false_type{});
        else if (fast) run(std::true_type{}, std::true_type{});
        else run(std::false_type{}, std::false_type{});
    }
    if (!have) return;

    // ---- fused epilogue (as lm_fit_packed_kernel)
    double rss[4], mss[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        sum[i].y0 = shift + dlt[i];
        rss[i] = sum[i].yy - sumsq_e(sum[i]);
        mss[i] = mss_from_e(sum[i], sd);
        if (act[i] && rss[i] <= kExactThreshold * sum[i].yy && sum[i].yy > 0.0) {
            const double2 ex = exact_pass_packed<KP, T>(U + v + i, a.ldU, n, a.Qt, sum[i].y0, cm, a.scale + sl[i], a.offset + sl[i],
                                                        a.nslices);
            rss[i] = ex.x;
            mss[i] = ex.y;
        }
    }
    if (a.out_vec2) {
#pragma unroll
        for (int i = 0; i < 4; i += 2)
            finish_pair<KP>(sum[i], rss[i], mss[i], act[i], sum[i + 1], rss[i + 1], mss[i + 1], act[i + 1], sd, a.out + v + i,
                            a.ldOut);
        return;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double *o = a.out + v + i;
        if (act[i]) finish_voxel<KP>(sum[i], rss[i], mss[i], sd, o, a.ldOut);
        else zero_row<KP>(sd.ncol_out, o, a.ldOut);
    }
}

// ------------------------------------------------------------------ TMA-staged variant (default when TMA can describe U)
// Same contract, K1's structure (lm_kernels.cu): persistent CTAs, one producer lane issues 2-D TMA boxes of the
// STORED samples (NJ subjects x up to 256 voxels, uint16 or float32 tensor map) plus, for scaled types, the NJ
// (scale, offset) pairs of the tile's two possible slices (1-D bulk copies out of `tab`, the tables re-packed as
// [slice][subject][2]) into a STAGES-deep ring guarded by full / empty mbarriers.  CW consumer warps own 4 adjacent
// voxels per thread, expand them in registers and accumulate; Q' stays resident in shared memory.  Bytes in flight
// are set by the ring (32 KB per CTA), not by registers, so the FP64 pipe rather than load latency limits it.
__global__ void pack_tables_kernel(const double *__restrict__ scale, const double *__restrict__ offset, int64_t nslices,
                                   int n, int n_pad, double2 *__restrict__ tab, int fold, double voxel_min)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nslices * n_pad) return;
    const int64_t s_ = i / n_pad;
    const int j = (int)(i % n_pad);
    double2 t = make_double2(0.0, 0.0);
    if (j < n) {
        const double sc = scale[s_ + (int64_t)j * nslices], of = offset[s_ + (int64_t)j * nslices];
        // FOLD: (A, C) of load_unit_smem's form; voxel_min + 65536 is an exact integer, the fma rounds C once
        const double one = fold == 2 ? 1048576.0 : 65536.0;        // the value of w's leading 1 in stored counts
        t = fold ? make_double2(sc * one, fma(-sc, voxel_min + one, of)) : make_double2(sc, of);
    }
    tab[i] = t;
}

template <int KP, typename T, int CW, int NJ, int STAGES, int MINB, int VPT, int FOLD>
__global__ void __launch_bounds__((CW + 1) * 32, MINB)
lm_fit_packed_tma_kernel(const __grid_constant__ CUtensorMap umap, PackedArgs a, const double2 *__restrict__ tab, int n_pad)
{
    constexpr bool SC = Stored<T>::kScaled;
    constexpr int TV = CW * 32 * VPT;
    constexpr int BOXW = TV < 256 ? TV : 256;
    constexpr int NBOX = TV / BOXW;
    constexpr int DATA_B = NJ * TV * (int)sizeof(T);
    constexpr int TAB_B = SC ? 2 * NJ * 16 : 0;
    constexpr int STAGE_B = (DATA_B + TAB_B + 127) & ~127;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *ring = smem_raw;
    const int n = a.n;
    const int nit = (n + NJ - 1) / NJ;
    double *sQ = reinterpret_cast<double *>(ring + (size_t)STAGES * STAGE_B);
    SmemDesign<KP> &sd = *reinterpret_cast<SmemDesign<KP> *>(sQ + (size_t)nit * NJ * KP);
    uint64_t *full = reinterpret_cast<uint64_t *>(&sd + 1);
    uint64_t *empty = full + STAGES;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    for (int i = tid; i < nit * NJ * KP; i += blockDim.x) sQ[i] = a.Qt[i];   // Qt rows are zero-padded to 16 (make_qt)
    load_design_to_smem<KP>(sd, a.design, tid, blockDim.x);
    if (tid == 0) {
        for (int s_ = 0; s_ < STAGES; ++s_) {
            mbar_init(&full[s_], 1);
            mbar_init(&empty[s_], CW);
        }
        fence_barrier_init();
    }
    __syncthreads();

    const int64_t ntiles = (a.V + TV - 1) / TV;

    if (warp == CW) {
        if ((tid & 31) == 0) {
            uint32_t stage = 0, phase = 0;
            for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const int c0 = (int)(t * TV);
                const int64_t sl_a = SC ? min((a.v_base + t * TV) / a.slice_len, a.nslices - 1) : 0;
                const int64_t sl_b = SC ? min(sl_a + 1, a.nslices - 1) : 0;
                for (int it = 0; it < nit; ++it) {
                    unsigned char *dst = ring + (size_t)stage * STAGE_B;
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], DATA_B + TAB_B);
#pragma unroll
                    for (int b = 0; b < NBOX; ++b)
                        tma_load_2d(dst + (size_t)b * NJ * BOXW * sizeof(T), &umap, &full[stage], c0 + b * BOXW, it * NJ);
                    if (SC) {
                        bulk_load_1d(dst + DATA_B, tab + sl_a * n_pad + it * NJ, NJ * 16, &full[stage]);
                        bulk_load_1d(dst + DATA_B + NJ * 16, tab + sl_b * n_pad + it * NJ, NJ * 16, &full[stage]);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        return;
    }

    uint32_t stage = 0, phase = 0;
    const int lane = tid & 31;
    const double cm = a.cmagic;
    const int boff = ((VPT * tid) / BOXW) * NJ * BOXW + (VPT * tid) % BOXW;     // this thread's samples inside a stage
    // a.V < 2^31 here (the launcher's condition), so tile and voxel indices are 32-bit: nothing 64-bit stays live across
    // the subject loop (the loop takes every register; what is live across it is spilled once per tile)
    const uint32_t kexp = 0x3ff00000u | (uint32_t)a.zero;
    double wp[VPT];           // FOLD = 2: the samples of one subject as doubles; the low words are written here, once
#pragma unroll
    for (int i = 0; i < VPT; ++i) wp[i] = __hiloint2double((int)kexp, a.zero);
    for (int t = blockIdx.x; t < (int)ntiles; t += gridDim.x) {
        // Which of the tile's two slices each voxel lies in: ONE division per tile.  `bnd` = offset inside the tile of the
        // first voxel of the tile's second slice (>= TV: the tile lies in one slice, or its first slice is the last one).
        int bnd = TV;
        if (SC) {
            const int64_t g0 = a.v_base + (int64_t)t * TV;
            const int64_t sl_a = min(g0 / a.slice_len, a.nslices - 1);
            if (sl_a + 1 < a.nslices) bnd = (int)min((sl_a + 1) * a.slice_len - g0, (int64_t)TV);
        }
        int sel[VPT];
        bool same = true;
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            sel[i] = (VPT * tid + i >= bnd) ? 1 : 0;
            same &= sel[i] == sel[0];
        }
        VoxelSums<KP> sum[VPT];
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
#pragma unroll
            for (int k = 0; k < KP; ++k) sum[i].e[k] = 0.0;
            sum[i].yy = 0.0;
            sum[i].y0 = 0.0;
        }
        // FOLD = 2: ONE shift per thread (the first sample of its first voxel) is folded into the offsets, C_j - shift being
        // formed once per subject and thread instead of one subtraction per sample; adjacent voxels differ little next to
        // their spread, and a voxel whose shifted |y|^2 still dwarfs its rss is redone from explicit residuals below.
        // ... WHEN adjacent voxels differ little next to their spread.  Where they do not -- a tissue / background edge
        // inside a thread's 4 voxels of an intensity image -- the one-pass sums of the far voxels would lose their digits and
        // those voxels would all take the explicit-residual path (measured: 26 ms instead of 2.1 ms with an edge in every 16th
        // quad).  So the first stage of the tile is looked at before it is consumed: per voxel, the offset dlt of its first
        // sample from the thread's shift against its spread over the stage's subjects; if any voxel of the tile is far out
        // (dlt^2 > 100 x variance: ten times inside the 1000 x rss rule of the epilogue) the consumers of the CTA vote the
        // tile into the per-voxel form of the loop (one more subtraction per sample: 28 instead of 25 FP64 per 4 samples).
        double shift = 0.0, dlt[VPT];
        bool per_voxel = false;
#pragma unroll
        for (int i = 0; i < VPT; ++i) dlt[i] = 0.0;
        if constexpr (FOLD == 2) {
            if (sd.shift) {
                mbar_wait(&full[stage], phase);          // the loop below waits for the same phase again: it passes at once
                const unsigned char *st = ring + (size_t)stage * STAGE_B;
                const uint16_t *mine0 = reinterpret_cast<const uint16_t *>(st) + boff;
                const double2 *tabs0 = reinterpret_cast<const double2 *>(st + DATA_B);
                const int nv0 = min(NJ, n);
                double var[VPT];
#pragma unroll
                for (int i = 0; i < VPT; ++i) var[i] = 0.0;
                // dlt[i] = voxel i's first sample EXACTLY as the loop will form it (the same fused multiply-add with the same
                // shifted offset): a voxel with one value in every subject then gives y' = 0 bit for bit in the per-voxel form
                // (or already here, when dlt is 0), its |y'|^2 is 0 and it keeps its rss = 0 classes without the exact path
#pragma unroll
                for (int jj = 0; jj < NJ; ++jj) {
                    if (jj < nv0) {
                        load_unit20_smem<VPT>(mine0 + jj * BOXW, wp, kexp);
                        if (jj == 0) {
                            const double2 so = tabs0[sel[0] * NJ];
                            shift = fma(so.x, wp[0], so.y);
                        }
#pragma unroll
                        for (int i = 0; i < VPT; ++i) {
                            const double2 so = tabs0[sel[i] * NJ + jj];
                            const double tv = fma(so.x, wp[i], so.y - shift);
                            if (jj == 0) dlt[i] = tv;
                            const double d = tv - dlt[i];
                            var[i] = fma(d, d, var[i]);
                        }
                    }
                }
                bool far = false;
#pragma unroll
                for (int i = 0; i < VPT; ++i) far |= dlt[i] * dlt[i] * (double)nv0 > 50.0 * var[i];       // var ~ 2 nv0 sigma^2
                int voted;
                asm volatile("{ .reg .pred p, q; setp.ne.s32 q, %1, 0; bar.red.or.pred p, 1, %2, q; selp.s32 %0, 1, 0, p; }"
                             : "=r"(voted) : "r"((int)far), "r"(CW * 32) : "memory");
                per_voxel = a.shift_mode == 2 || (a.shift_mode == 0 && voted != 0);
            }
        }
        for (int it = 0; it < nit; ++it) {
            mbar_wait(&full[stage], phase);
            const unsigned char *st = ring + (size_t)stage * STAGE_B;
            const T *mine = reinterpret_cast<const T *>(st) + boff;
            const double2 *tabs = reinterpret_cast<const double2 *>(st + DATA_B);
            const double *q = sQ + (size_t)it * NJ * KP;
            const int nvalid = min(NJ, n - it * NJ);      // rows >= n of the last box are TMA zero-fill: skipped
            auto run = [&](auto same_tag, auto full_tag, auto pv_tag) {
                constexpr bool SAME = decltype(same_tag)::value;
                constexpr bool FULL = decltype(full_tag)::value;     // all NJ rows real: no per-row test in the hot loop
                constexpr bool PV = decltype(pv_tag)::value;         // per-voxel shifts (FOLD = 2, a tile with far-out voxels)
#pragma unroll
                for (int jj = 0; jj < NJ; ++jj) {
                    if (FULL || jj < nvalid) {
                        double y[VPT];
                        if constexpr (FOLD == 2) load_unit20_smem<VPT>(reinterpret_cast<const uint16_t *>(mine) + jj * BOXW, wp, kexp);
                        else if constexpr (FOLD == 1) load_unit_smem<VPT>(mine + jj * BOXW, y);
                        else load_convert_smem<T, VPT>(mine + jj * BOXW, y, cm);
                        double2 so0 = make_double2(1.0, 0.0), so1 = so0;
                        if (SC) {
                            so0 = tabs[(SAME ? sel[0] : 0) * NJ + jj];
                            if (!SAME) so1 = tabs[NJ + jj];
                            if constexpr (FOLD == 2) {
                                so0.y -= shift;
                                if (!SAME) so1.y -= shift;
                            }
                        }
#pragma unroll
                        for (int i = 0; i < VPT; ++i) {
                            double tv = FOLD == 2 ? wp[i] : y[i];
                            if (SC) {
                                const double2 so = (SAME || sel[i] == 0) ? so0 : so1;
                                tv = FOLD ? fma(so.x, tv, so.y) : descale(tv, so.x, so.y);
                            }
                            double yv = tv;
                            if constexpr (FOLD != 2) {
                                if (it == 0 && jj == 0 && sd.shift) sum[i].y0 = tv;
                                yv = tv - sum[i].y0;
                            } else if constexpr (PV) {
                                yv = tv - dlt[i];
                            }
                            sum[i].yy = fma(yv, yv, sum[i].yy);
#pragma unroll
                            for (int k = 0; k < KP; ++k) sum[i].e[k] = fma(q[jj * KP + k], yv, sum[i].e[k]);
                        }
                    }
                }
            };
            if (FOLD == 2 && per_voxel) {
                if (same && nvalid == NJ) run(std::true_type{}, std::true_type{}, std::true_type{});
                else run(std::false_type{}, std::false_type{}, std::true_type{});
            } else if (same) {
                if (nvalid == NJ) run(std::true_type{}, std::true_type{}, std::false_type{});
                else run(std::true_type{}, std::false_type{}, std::false_type{});
            } else {
                run(std::false_type{}, std::false_type{}, std::false_type{});
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if constexpr (FOLD == 2) {
#pragma unroll
            for (int i = 0; i < VPT; ++i) sum[i].y0 = per_voxel ? shift + dlt[i] : shift;
        }
        // ---- fused epilogue
        const int64_t v = (int64_t)t * TV + VPT * tid;
        bool in[VPT];
#pragma unroll
        for (int i = 0; i < VPT; ++i) in[i] = v + i < a.V;
        double rss[VPT], mss[VPT];
        unsigned redo = 0;       // voxels whose one-pass rss lost too many digits: redone from explicit residuals below
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            rss[i] = sum[i].yy - sumsq_e(sum[i]);
            mss[i] = mss_from_e(sum[i], sd);
            if (in[i] && rss[i] <= kExactThreshold * sum[i].yy && sum[i].yy > 0.0) redo |= 1u << i;
        }
        if (a.out_vec2 && in[VPT - 1]) {
#pragma unroll
            for (int i = 0; i + 1 < VPT; i += 2)
                finish_pair<KP>(sum[i], rss[i], mss[i], true, sum[i + 1], rss[i + 1], mss[i + 1], true, sd, a.out + v + i, a.ldOut);
        } else {
#pragma unroll
            for (int i = 0; i < VPT; ++i)
                if (in[i]) finish_voxel<KP>(sum[i], rss[i], mss[i], sd, a.out + v + i, a.ldOut);
        }
        if (redo) {
#pragma unroll
            for (int i = 0; i < VPT; ++i) {
                if (redo >> i & 1) {
                    const int64_t sli = SC ? min((a.v_base + v + i) / a.slice_len, a.nslices - 1) : 0;
                    exact_voxel_packed<KP, T>(static_cast<const T *>(a.U) + v + i, a.ldU, n, a.Qt, sum[i].y0, cm, a.scale + sli,
                                              a.offset + sli, a.nslices, &sd, a.out + v + i, a.ldOut);
                }
            }
        }
    }
}

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn packed_encode_tiled()
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

// returns cudaErrorNotSupported when this layout / size is not for the TMA variant (caller falls back)
template <int KP, typename T, int MINB, int FOLD, int CW = 4, int VPT = 4, int NJ = 8, int STAGES = 4>
cudaError_t launch_packed_tma(const PackedArgs &a, const LmPlan &plan, LmScratch &scr, cudaStream_t st, LmLaunchInfo *info)
{
    constexpr bool SC = Stored<T>::kScaled;
    constexpr int TV = CW * 32 * VPT, BOXW = TV < 256 ? TV : 256;
    constexpr int STAGE_B = (NJ * TV * (int)sizeof(T) + (SC ? 2 * NJ * 16 : 0) + 127) & ~127;
    if (a.mask || KP > 8) return cudaErrorNotSupported;
    if ((reinterpret_cast<uintptr_t>(a.U) & 15) || (a.ldU * sizeof(T)) % 16 || a.V > 0x7fffffffLL) return cudaErrorNotSupported;
    if (SC && a.slice_len < TV) return cudaErrorNotSupported;          // a tile may then touch more than two slices
    if (a.V < (int64_t)plan.sm_count * TV) return cudaErrorNotSupported;   // too few tiles to fill the GPU
    const int nit = (a.n + NJ - 1) / NJ;
    const size_t smem = (size_t)STAGES * STAGE_B + (size_t)nit * NJ * KP * 8 + sizeof(SmemDesign<KP>) + 2 * STAGES * 8 + 128;
    if (smem > (size_t)(227 * 1024) / MINB - 1024) return cudaErrorNotSupported;   // MINB CTAs must fit an SM (1 KB reserved each)
    EncodeTiledFn enc = packed_encode_tiled();
    if (!enc) return cudaErrorNotSupported;
    CUtensorMap map;
    const cuuint64_t dims[2] = {(cuuint64_t)a.V, (cuuint64_t)a.n};
    const cuuint64_t strides[1] = {(cuuint64_t)a.ldU * sizeof(T)};
    const cuuint32_t box[2] = {BOXW, NJ};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt = sizeof(T) == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    if (enc(&map, dt, 2, const_cast<void *>(a.U), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return cudaErrorNotSupported;
    const int n_pad = nit * NJ;
    const double2 *tab = nullptr;
    if (SC) {
        cudaError_t e = scr.ensure_bytes((size_t)a.nslices * n_pad * sizeof(double2), st);
        if (e != cudaSuccess) return e;
        const int64_t cnt = a.nslices * n_pad;
        pack_tables_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(a.scale, a.offset, a.nslices, a.n, n_pad,
                                                                        reinterpret_cast<double2 *>(scr.flags), FOLD,
                                                                        a.cmagic - 4503599627370496.0);
        if (info) info->launches++;
        tab = reinterpret_cast<const double2 *>(scr.flags);
    }
    auto kern = lm_fit_packed_tma_kernel<KP, T, CW, NJ, STAGES, MINB, VPT, FOLD>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int64_t ntiles = (a.V + TV - 1) / TV;
    const int64_t slots = (int64_t)plan.sm_count * MINB;
    kern<<<(int)(ntiles < slots ? ntiles : slots), (CW + 1) * 32, smem, st>>>(map, a, tab, n_pad);
    if (info) { info->launches++; info->used_tma = 1; }
    return cudaGetLastError();
}

template <int KP, typename T>
cudaError_t launch_packed_t(const PackedArgs &a, const LmPlan &plan, LmScratch &scr, cudaStream_t st, LmLaunchInfo *info)
{
    if (plan.variant != LmVariant::Ldg) {
        if constexpr (KP <= 8) {
            // CTAs per SM: as many as the accumulators leave registers for
            constexpr int MINB = KP <= 2 ? 4 : (KP <= 4 ? 3 : 2);
            // Shape: 4 consumer warps x 4 voxels per thread, MINB CTAs/SM.  uint16 samples take one fused multiply-add
            // each (FOLD) unless the reader's three roundings are asked for, and 16 subjects per stage in a 3-deep ring:
            // the barrier round trip per stage is then paid half as often (config 2: 2.30 ms against 2.45 ms for 8
            // subjects x 4 stages; the ring's bytes in flight are the same).  Measured alternatives: profiles/README.md.
            cudaError_t e;
            if constexpr (Stored<T>::kScaled) {
                if (plan.stored_exact) {
                    e = launch_packed_tma<KP, T, MINB, 0>(a, plan, scr, st, info);
                } else {
                    // measured alternatives: one more CTA per SM at 96 registers
                    if constexpr (KP == 4) {
                        cudaError_t ex = cudaErrorNotSupported;
                        if (plan.tma_cfg == 2) ex = launch_packed_tma<KP, T, 4, 2, 4, 4, 16, 2>(a, plan, scr, st, info);
                        if (plan.tma_cfg == 3) ex = launch_packed_tma<KP, T, 4, 2, 4, 4, 8, 4>(a, plan, scr, st, info);
                        // (32 or 24 subjects per stage in a 2-deep ring: 2.096 / 2.065 ms against 2.10 ms -- the barrier round trips
                        // are no longer what the kernel waits on; not kept)
                        if (ex != cudaErrorNotSupported) return ex;
                    }
                    // RMG_TMA_CFG: 1 = 8 subjects x 4 stages, 4 = the round-2a form (w = 1 + u 2^-16, one shift per voxel)
                    if (plan.tma_cfg == 4) e = launch_packed_tma<KP, T, MINB, 1, 4, 4, 16, 3>(a, plan, scr, st, info);
                    else
                        e = plan.tma_cfg == 1 ? cudaErrorNotSupported
                                              : launch_packed_tma<KP, T, MINB, 2, 4, 4, 16, 3>(a, plan, scr, st, info);
                    if (e == cudaErrorNotSupported) e = launch_packed_tma<KP, T, MINB, 2>(a, plan, scr, st, info);
                }
            } else {
                e = launch_packed_tma<KP, T, MINB, 0>(a, plan, scr, st, info);
            }
            if (e != cudaErrorNotSupported) return e;
        }
    }
    // vector path: 4 samples per load need an 8-byte (u16) / 16-byte (f32) aligned base and row pitch
    const bool vec = (reinterpret_cast<uintptr_t>(a.U) % (4 * sizeof(T)) == 0) && (a.ldU % 4 == 0);
    constexpr int QW = ((KP + 1) & ~1) + (Stored<T>::kScaled ? 4 : 0);
    int JT = a.n;
    while ((size_t)JT * QW * 8 > 40 * 1024) JT = ((JT + 1) / 2 + 7) & ~7;   // multiple of the unroll depth
    const size_t smem = (size_t)JT * QW * 8;
    // masked fits: compact the mask into a list of voxel quads and fit only those (RMG_PACKED_NOCOMPACT=1: the plain kernel)
    static const bool env_no_compact = [] { const char *s_ = std::getenv("RMG_PACKED_NOCOMPACT"); return s_ && std::atoi(s_) != 0; }();
    if (a.mask && vec && a.V % 4 == 0 && a.V / 4 < 0xffffffffLL && !plan.no_compact && !env_no_compact) {
        const int64_t nquads = a.V / 4;
        const int nb = (int)((nquads + kQuadBlock - 1) / kQuadBlock);
        const size_t off_list = ((size_t)nb * 4 + 4 + 255) & ~size_t(255);
        cudaError_t e = scr.ensure_bytes(off_list + (size_t)nquads * 4, st);
        if (e != cudaSuccess) return e;
        unsigned int *counts = reinterpret_cast<unsigned int *>(scr.flags);
        unsigned int *n_quads = counts + nb;
        uint32_t *list = reinterpret_cast<uint32_t *>(scr.flags + off_list);
        const int ncol = 2 * a.p + 3;
        quad_count_kernel<<<nb, kQuadBlock, 0, st>>>(a.mask, a.V, a.mask_lo, a.mask_hi, counts);
        quad_scan_kernel<<<1, 1024, 0, st>>>(counts, nb, n_quads);
        quad_compact_kernel<<<nb, kQuadBlock, 0, st>>>(a.mask, a.V, a.mask_lo, a.mask_hi, counts, list, a.out, a.ldOut, ncol, a.out_vec2);
        const int64_t qblocks = (nquads + 127) / 128;
        auto goq = [&](auto kern) -> cudaError_t {
            cudaError_t e2 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e2 != cudaSuccess) return e2;
            kern<<<(unsigned)qblocks, 128, smem, st>>>(a, list, n_quads, JT);
            return cudaGetLastError();
        };
        e = plan.stored_exact ? goq(lm_fit_packed_quads_kernel<KP, T, false>) : goq(lm_fit_packed_quads_kernel<KP, T, true>);
        if (info) { info->launches += 4; info->used_tma = 0; info->split = 1; }
        return e;
    }
    const int VPT = vec ? 4 : 1;
    const int64_t per_block = (int64_t)128 * VPT;
    const int64_t blocks = (a.V + per_block - 1) / per_block;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidValue;
    auto go = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<(unsigned)blocks, 128, smem, st>>>(a, JT);
        return cudaGetLastError();
    };
    const cudaError_t e = vec ? go(lm_fit_packed_kernel<KP, T, 4>) : go(lm_fit_packed_kernel<KP, T, 1>);
    if (info) { info->launches++; info->used_tma = 0; info->split = 1; }
    return e;
}

template <int KP>
cudaError_t launch_packed_kp(const PackedArgs &a, const LmPlan &plan, LmScratch &scr, cudaStream_t st, LmLaunchInfo *info)
{
    if (a.dtype == kStoredU16) return launch_packed_t<KP, uint16_t>(a, plan, scr, st, info);
    if (a.dtype == kStoredF32) return launch_packed_t<KP, float>(a, plan, scr, st, info);
    return cudaErrorInvalidValue;
}

}  // namespace

cudaError_t launch_expand_stored(const void *U, int dtype, int64_t V, int64_t ldU, int n, const double *scale,
                                 const double *offset, double cmagic, int64_t slice_len, int64_t nslices, int64_t v_base,
                                 double *Y, int64_t ldY, cudaStream_t st)
{
    if (V <= 0 || n <= 0) return cudaSuccess;
    const int64_t blocks = (V + 255) / 256;
    if (blocks > 0x7fffffffLL || n > 65535) return cudaErrorInvalidValue;
    const dim3 grid((unsigned)blocks, (unsigned)n);
    if (dtype == kStoredU16)
        expand_stored_kernel<uint16_t><<<grid, 256, 0, st>>>(static_cast<const uint16_t *>(U), V, ldU, scale, offset, cmagic,
                                                             std::max<int64_t>(slice_len, 1), std::max<int64_t>(nslices, 1),
                                                             v_base, Y, ldY);
    else if (dtype == kStoredF32)
        expand_stored_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float *>(U), V, ldU, nullptr, nullptr, 0.0, 1, 1, 0, Y, ldY);
    else
        return cudaErrorInvalidValue;
    return cudaGetLastError();
}

cudaError_t launch_lm_packed(const PackedArgs &a_in, const LmPlan &plan, LmScratch &scr, cudaStream_t st, LmLaunchInfo *info)
{
    if (a_in.V <= 0) return cudaSuccess;
    PackedArgs a = a_in;
    a.out_vec2 = ((reinterpret_cast<uintptr_t>(a.out) & 15) == 0 && a.ldOut % 2 == 0) ? 1 : 0;
    static const int env_shift_mode = [] { const char *s_ = std::getenv("RMG_STORED_SHIFT"); return s_ ? std::atoi(s_) : 0; }();
    a.shift_mode = env_shift_mode;
    switch (lm_padded_p(a.p)) {
#define RMG_CASE(K) case K: return launch_packed_kp<K>(a, plan, scr, st, info);
#ifdef RMG_QUICK_KP4      // developer switch: compile the p = 4 instances only (SASS inspection in seconds)
        RMG_CASE(4)
#else
        RMG_CASE(1) RMG_CASE(2) RMG_CASE(3) RMG_CASE(4) RMG_CASE(5) RMG_CASE(6) RMG_CASE(7) RMG_CASE(8)
        RMG_CASE(12) RMG_CASE(16)
#endif
#undef RMG_CASE
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace rmg
