// lm_common.cuh -- device-side constants and the fused per-voxel statistics epilogue
// shared by the fit kernels (K1/K2) and the permutation kernel (K3).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace rmg {

constexpr int kMaxPDev = 32;

// Everything the epilogue needs that depends only on the design.  Lives in global
// memory (one per design; the permutation kernel indexes an array of them).
struct DevDesign {
    int n, p, k, shift;     // shift = 1: 1 in span(X), voxels are centred on their first sample
    double gap;             // n - |Q'1|^2  (0 when shift)
    double inv_n;           // 1/n  (1/sum(w) for weighted fits)
    double rdf;             // n - p         (NOT n - rank: src/modelling_functions.c:534)
    double pm1;             // p - 1         (:542)
    double loglik_c;        // log(2*pi) + 1 - log(n)   (:527)
    double mhalf_n;         // -0.5 * n
    double loglik_add;      // 0.5 * sum(log w) for weighted fits (src/weighted_least_squares.c:112), else 0
    double q1[kMaxPDev];    // Q'1
    double ct[kMaxPDev];    // diag((R'R)^-1), un-pivoted as the reference un-pivots se (:560-564)
    double Rinv[kMaxPDev * kMaxPDev];  // row-major: Rinv[i*kMaxPDev + j], upper triangle, k x k
    // anova epilogue (voxel_anova, src/minc_anova.c:4-80): mode 1 writes nterms F-statistics
    int mode, nterms, ncol_out, pad_;   // ncol_out = 2p+3 (lm) or nterms (anova)
    int term[kMaxPDev];     // asgn[i]: term of the effect at (pivoted) position i; 0 = intercept
    double inv_df[kMaxPDev + 1];  // 1 / (number of columns of term t)
};

// Per-CTA shared-memory copy of the k-dependent part, padded to the kernel's KP.
template <int KP>
struct SmemDesign {
    double q1[KP];
    double ct[KP];
    double isct[KP];        // 1 / sqrt(ct): t_i = beta_i * isct_i * sqrt(rdf / rss)
    double Rinv[KP * KP];   // row-major KP x KP
    double gap, inv_n, rdf, pm1, inv_pm1, loglik_c, mhalf_n, loglik_add;
    int n, p, k, shift;
    int mode, nterms, ncol_out, pad_;
    int term[KP];
    double inv_df[KP + 1];
};

template <int KP>
__device__ __forceinline__ void load_design_to_smem(SmemDesign<KP> &s, const DevDesign *__restrict__ d,
                                                    int tid, int nthreads)
{
    for (int i = tid; i < KP; i += nthreads) {
        s.q1[i] = d->q1[i];
        s.ct[i] = d->ct[i];
        s.isct[i] = 1.0 / sqrt(d->ct[i]);
        s.term[i] = d->term[i];
    }
    for (int i = tid; i < KP + 1; i += nthreads) s.inv_df[i] = d->inv_df[i];
    for (int i = tid; i < KP * KP; i += nthreads) s.Rinv[i] = d->Rinv[(i / KP) * kMaxPDev + (i % KP)];
    if (tid == 0) {
        s.gap = d->gap; s.inv_n = d->inv_n; s.rdf = d->rdf; s.pm1 = d->pm1; s.inv_pm1 = 1.0 / d->pm1;
        s.loglik_c = d->loglik_c; s.mhalf_n = d->mhalf_n; s.loglik_add = d->loglik_add;
        s.n = d->n; s.p = d->p; s.k = d->k; s.shift = d->shift;
        s.mode = d->mode; s.nterms = d->nterms; s.ncol_out = d->ncol_out;
    }
}

// One-pass sums of one voxel: e = Q'(y - y0), yy = |y - y0|^2.
// One-pass rss = yy - |e|^2 loses digits when the model explains almost all of yy; callers
// test needs_exact() and then recompute rss/mss from explicit residuals (exact_pass()).
template <int KP>
struct VoxelSums {
    double e[KP];
    double yy;
    double y0;
};

template <int KP>
__device__ __forceinline__ double sumsq_e(const VoxelSums<KP> &s)
{
    double ee = 0.0;
#pragma unroll
    for (int k = 0; k < KP; ++k) ee = fma(s.e[k], s.e[k], ee);
    return ee;
}

// mss = sum((fitted - mean(fitted))^2)   (:529-532), from e alone:
//   fitted = Q e, sum(fitted) = q1.e =: s, |fitted|^2 = |e|^2
//   mss = |e - q1 s/n|^2 + (n - |q1|^2)(s/n)^2        (both terms >= 0: no cancellation)
template <int KP>
__device__ __forceinline__ double mss_from_e(const VoxelSums<KP> &v, const SmemDesign<KP> &d)
{
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < KP; ++k) s = fma(d.q1[k], v.e[k], s);
    const double m = s * d.inv_n;
    double mss = 0.0;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        const double dk = fma(-d.q1[k], m, v.e[k]);
        mss = fma(dk, dk, mss);
    }
    return fma(d.gap * m, m, mss);
}

// Relative threshold below which the one-pass rss is replaced by explicit residuals.
// One-pass error ~ n*eps*yy/rss worst case; 1e-3 keeps it under ~1e-10 for n <= 4000.
constexpr double kExactThreshold = 1e-3;

// Write one result row.  Column layout: src/modelling_functions.c:1159-1174.
//   out[v + c*ld]: c=0 F, 1 R2, 2..p+1 beta, p+2..2p+1 t, 2p+2 logLik
template <int KP>
__device__ __forceinline__ void finish_voxel(const VoxelSums<KP> &v, double rss, double mss,
                                             const SmemDesign<KP> &d, double *__restrict__ out,
                                             int64_t ld)
{
    const int p = d.p;
    if (d.mode == 1) {
        const double resvar = rss / d.rdf;                   // :537
        // voxel_anova: F_t = (sum of effects^2 over the term's columns / df_t) / (ssr / (n-p))
        double e2[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            const double ek = fma(v.y0, d.q1[k], v.e[k]);
            e2[k] = ek * ek;
        }
        for (int t = 1; t <= d.nterms; ++t) {
            double ss = 0.0;
#pragma unroll
            for (int k = 0; k < KP; ++k) ss += (d.term[k] == t) ? e2[k] : 0.0;
            out[(int64_t)(t - 1) * ld] = (ss * d.inv_df[t]) / resvar;
        }
        return;
    }
    // One division and one square root per voxel instead of p + 2 and p:  with r = rdf / rss = 1 / resvar,
    //   F = (mss / (p-1)) / resvar = mss * (1/(p-1)) * r            (:542)
    //   t_i = beta_i / sqrt(ct_i * resvar) = beta_i * (1/sqrt(ct_i)) * sqrt(r)   (:563, :568)
    // The IEEE classes of the degenerate cases are those of the reference's expressions: rss = 0 gives r = Inf, hence
    // F = Inf or NaN (mss = 0) and t = +-Inf or NaN (beta = 0); rdf = 0 gives r = 0, hence F = 0 and t = 0.
    const double r = d.rdf / rss;
    const double sr = sqrt(r);
    out[0] = (mss * d.inv_pm1) * r;
    out[ld] = mss / (mss + rss);                             // :572
    out[(int64_t)(2 * p + 2) * ld] = d.mhalf_n * (d.loglik_c + log(rss)) + d.loglik_add;  // :527
    // un-shift: e(y) = e(y - y0) + y0 * Q'1
    double e[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) e[k] = fma(v.y0, d.q1[k], v.e[k]);
#pragma unroll
    for (int i = 0; i < KP; ++i) {
        if (i < p) {
            double b = 0.0;
#pragma unroll
            for (int j = i; j < KP; ++j) b = fma(d.Rinv[i * KP + j], e[j], b);
            out[(int64_t)(2 + i) * ld] = b;                            // beta, pivoted order
            out[(int64_t)(2 + p + i) * ld] = (b * d.isct[i]) * sr;
        }
    }
}

// Two adjacent voxels (v, v+1) at once with 16-byte stores: one st.global.v2.f64 per result
// column instead of two half-used sectors.  `ok*` false -> that row is +0.0 (outside the mask).
// Requires out + v 16-byte aligned and ld even (checked on the host).
template <int KP>
__device__ __forceinline__ void finish_pair(const VoxelSums<KP> &a, double rssa, double mssa, bool oka,
                                            const VoxelSums<KP> &b, double rssb, double mssb, bool okb,
                                            const SmemDesign<KP> &d, double *__restrict__ out, int64_t ld)
{
    const int p = d.p;
    auto put = [&](int c, double x, double y) {
        *reinterpret_cast<double2 *>(out + (int64_t)c * ld) = make_double2(oka ? x : 0.0, okb ? y : 0.0);
    };
    if (d.mode == 1) {
        const double rva = rssa / d.rdf, rvb = rssb / d.rdf;
        double ea2[KP], eb2[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            const double x = fma(a.y0, d.q1[k], a.e[k]), y = fma(b.y0, d.q1[k], b.e[k]);
            ea2[k] = x * x;
            eb2[k] = y * y;
        }
        for (int t = 1; t <= d.nterms; ++t) {
            double sa = 0.0, sb = 0.0;
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                sa += (d.term[k] == t) ? ea2[k] : 0.0;
                sb += (d.term[k] == t) ? eb2[k] : 0.0;
            }
            put(t - 1, (sa * d.inv_df[t]) / rva, (sb * d.inv_df[t]) / rvb);
        }
        return;
    }
    const double ra = d.rdf / rssa, rb = d.rdf / rssb;       // see finish_voxel
    const double sra = sqrt(ra), srb = sqrt(rb);
    put(0, (mssa * d.inv_pm1) * ra, (mssb * d.inv_pm1) * rb);
    put(1, mssa / (mssa + rssa), mssb / (mssb + rssb));
    put(2 * p + 2, d.mhalf_n * (d.loglik_c + log(rssa)) + d.loglik_add, d.mhalf_n * (d.loglik_c + log(rssb)) + d.loglik_add);
    double ea[KP], eb[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        ea[k] = fma(a.y0, d.q1[k], a.e[k]);
        eb[k] = fma(b.y0, d.q1[k], b.e[k]);
    }
#pragma unroll
    for (int i = 0; i < KP; ++i) {
        if (i < p) {
            double ba = 0.0, bb = 0.0;
#pragma unroll
            for (int j = i; j < KP; ++j) {
                ba = fma(d.Rinv[i * KP + j], ea[j], ba);
                bb = fma(d.Rinv[i * KP + j], eb[j], bb);
            }
            put(2 + i, ba, bb);
            put(2 + p + i, (ba * d.isct[i]) * sra, (bb * d.isct[i]) * srb);
        }
    }
}

template <int KP>
__device__ __forceinline__ void zero_row(int ncol, double *__restrict__ out, int64_t ld)
{
    for (int c = 0; c < ncol; ++c) out[(int64_t)c * ld] = 0.0;
}

__device__ __forceinline__ bool mask_selects(double m, double lo, double hi)
{
    return m > lo - 0.5 && m < hi + 0.5;  // src/modelling_functions.c:1063-1066
}

// Streaming 8/16-byte loads that do not pollute L1 (Y is read exactly once).
__device__ __forceinline__ double ldg_stream(const double *p)
{
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double2 ldg_stream2(const double *p)
{
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

// The same 16-byte load with an explicit L2 prefetch size.  A compacted masked fit reads runs of ~90 selected voxels per
// subject; what the memory system fetches beyond a run's ends (measured: ~80 bytes per run end with the default policy,
// 21 % of the traffic on the 38 % ellipsoid) is set by this size, not by the 32-byte sectors the loads touch.
__device__ __forceinline__ double2 ldg_stream2_hint(const double *p, int l2_bytes)
{
    double2 v;
    if (l2_bytes == 64)
        asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    else if (l2_bytes == 128)
        asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    else if (l2_bytes == 256)
        asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    else
        asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

// Explicit-residual recomputation for one voxel (rare path).  Qt is the row-major [n][KP] copy
// of Q in global memory; y is read again with stride ld.  Deliberately takes only scalars and
// pointers and recomputes e itself: passing the caller's accumulator struct by reference to a
// non-inlined function forces it onto the local-memory stack for EVERY voxel (measured: 0.46 GB of
// extra DRAM writes per config-2 launch).  Returns (rss, mss).
// rowscale (sqrt of the weights) may be null.
template <int KP>
__device__ __noinline__ double2 exact_pass(const double *__restrict__ y, int64_t ld, int n,
                                           const double *__restrict__ Qt, double y0,
                                           const double *__restrict__ rowscale = nullptr)
{
    double e[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) e[k] = 0.0;
    for (int j = 0; j < n; ++j) {
        const double yv = (y[(int64_t)j * ld] - y0) * (rowscale ? rowscale[j] : 1.0);
#pragma unroll
        for (int k = 0; k < KP; ++k) e[k] = fma(__ldg(Qt + (size_t)j * KP + k), yv, e[k]);
    }
    // in the scaled problem fit_j = sqrt(w_j) * fitted_j: weighted mean m = sum(s_j fit_j) / sum(s_j^2),
    // mss = sum (fit_j - s_j m)^2   (s_j = 1: the ordinary mean / mss)
    double r2 = 0.0, sf = 0.0, sw = 0.0;
    for (int j = 0; j < n; ++j) {
        const double sj = rowscale ? rowscale[j] : 1.0;
        double fit = 0.0;
#pragma unroll
        for (int k = 0; k < KP; ++k) fit = fma(__ldg(Qt + (size_t)j * KP + k), e[k], fit);
        const double r = (y[(int64_t)j * ld] - y0) * sj - fit;
        r2 = fma(r, r, r2);
        sf = fma(sj, fit, sf);
        sw = fma(sj, sj, sw);
    }
    const double mean = sf / sw;
    double m2 = 0.0;
    for (int j = 0; j < n; ++j) {
        const double sj = rowscale ? rowscale[j] : 1.0;
        double fit = 0.0;
#pragma unroll
        for (int k = 0; k < KP; ++k) fit = fma(__ldg(Qt + (size_t)j * KP + k), e[k], fit);
        const double dlt = fma(-sj, mean, fit);
        m2 = fma(dlt, dlt, m2);
    }
    return make_double2(r2, m2);
}

}  // namespace rmg
