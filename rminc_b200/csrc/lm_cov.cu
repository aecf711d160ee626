// lm_cov.cu -- voxel-wise covariate designs, hand-written for sm_100a.
//
// Reference: a MINC file term on the right-hand side of the formula makes the design of voxel v
//   X_v = [Xs | z_v]   (p = ncol(Xs) + 1),   or   [1 | z_v]   (p = 2) without a static part
// (src/vertex_modelling.c:112-140, src/modelling_functions.c:1118-1144, R/minc_interface.R:1266-1300),
// and voxel_lm (src/modelling_functions.c:448-576) factorises that n x p matrix from scratch per voxel.
//
// Here the static part Xs = Q R is factored once on the host (design.cpp).  Per voxel one pass over the
// two streamed operands y and z accumulates
//     e = Q'y, f = Q'z, |y|^2, |z|^2, z'y, sum(z)
// (all after the first-sample shift y - y0, z - z0 when 1 is in span(Xs), see lm_common.cuh H1) and the
// last Householder column follows in closed form:
//     z_perp = z - Q f,  d^2 = |z|^2 - |f|^2,  g = z_perp'y = z'y - f.e
//     beta_z = g / d^2,  beta_s = R^-1 (e - f beta_z),  rss = |y|^2 - |e|^2 - g^2/d^2
//     diag((R_v'R_v)^-1) = [ct_s + (R^-1 f)^2 / d^2,  1/d^2]
// Voxels where d^2 or rss lost digits to cancellation recompute them from explicit residuals.
// dqrdc2's limited pivoting for the last column (the only per-voxel pivot decision when Xs has full
// rank): |z_perp| < 1e-7 |z| -> rank p-1, beta_z = 0, t_z = 0, residuals of the static fit; an all-zero
// z is an exact zero on R's diagonal, i.e. the reference's dpotri error -- reported through *status.
//
// Roofline: HBM, 16 n + 8 (2p+3) bytes per voxel (two streamed operands).
#include <cuda_runtime.h>

#include <cstdint>

#include "lm_common.cuh"
#include "lm_launch.hpp"

namespace rmg {

template <int KS>
struct CovSums {
    double e[KS], f[KS];
    double yy, zz, zy, sz;
    double y0, z0;
};

struct CovExact {
    double d2, g, rss, mss, rss_s, mss_s;   // *_s: the fit on the static part alone (negligible covariate)
};

// Explicit-residual recomputation for one voxel (rare path); scalars and pointers only, see exact_pass().
template <int KS>
__device__ __noinline__ CovExact exact_cov(const double *__restrict__ y, const double *__restrict__ z, int64_t ld, int n,
                                           const double *__restrict__ Qt, double y0, double z0, const int32_t *__restrict__ zrow)
{
    double e[KS], f[KS];
#pragma unroll
    for (int k = 0; k < KS; ++k) e[k] = f[k] = 0.0;
    for (int j = 0; j < n; ++j) {
        const double yv = y[(int64_t)j * ld] - y0, zv = z[(int64_t)(zrow ? zrow[j] : j) * ld] - z0;
#pragma unroll
        for (int k = 0; k < KS; ++k) {
            const double q = __ldg(Qt + (size_t)j * KS + k);
            e[k] = fma(q, yv, e[k]);
            f[k] = fma(q, zv, f[k]);
        }
    }
    double d2 = 0.0, g = 0.0, rr = 0.0, sfit = 0.0;
    for (int j = 0; j < n; ++j) {
        double fy = 0.0, fz = 0.0;
#pragma unroll
        for (int k = 0; k < KS; ++k) {
            const double q = __ldg(Qt + (size_t)j * KS + k);
            fy = fma(q, e[k], fy);
            fz = fma(q, f[k], fz);
        }
        const double r = (y[(int64_t)j * ld] - y0) - fy, zp = (z[(int64_t)(zrow ? zrow[j] : j) * ld] - z0) - fz;
        d2 = fma(zp, zp, d2);
        g = fma(zp, r, g);
        rr = fma(r, r, rr);
        sfit += fy;
    }
    const double bz = d2 > 0.0 ? g / d2 : 0.0;
    double rss = 0.0, sfull = 0.0;
    for (int j = 0; j < n; ++j) {
        double fy = 0.0, fz = 0.0;
#pragma unroll
        for (int k = 0; k < KS; ++k) {
            const double q = __ldg(Qt + (size_t)j * KS + k);
            fy = fma(q, e[k], fy);
            fz = fma(q, f[k], fz);
        }
        const double zp = (z[(int64_t)(zrow ? zrow[j] : j) * ld] - z0) - fz;
        const double r = ((y[(int64_t)j * ld] - y0) - fy) - bz * zp;
        rss = fma(r, r, rss);
        sfull += fma(bz, zp, fy);
    }
    const double ms = sfit / n, mf = sfull / n;
    double mss = 0.0, mss_s = 0.0;
    for (int j = 0; j < n; ++j) {
        double fy = 0.0, fz = 0.0;
#pragma unroll
        for (int k = 0; k < KS; ++k) {
            const double q = __ldg(Qt + (size_t)j * KS + k);
            fy = fma(q, e[k], fy);
            fz = fma(q, f[k], fz);
        }
        const double zp = (z[(int64_t)(zrow ? zrow[j] : j) * ld] - z0) - fz;
        const double a = fy - ms, b = fma(bz, zp, fy) - mf;
        mss_s = fma(a, a, mss_s);
        mss = fma(b, b, mss);
    }
    return CovExact{d2, g, rss, mss, rr, mss_s};
}

// Result row of one voxel.  Column layout as finish_voxel() with p = ps + 1.
template <int KS>
__device__ __forceinline__ void finish_cov(const CovSums<KS> &v, const SmemDesign<KS> &d, const double *__restrict__ y,
                                           const double *__restrict__ z, int64_t ld, const double *__restrict__ Qt,
                                           double *__restrict__ out, int64_t ldo, int *__restrict__ status,
                                           const int32_t *__restrict__ zrow)
{
    const int ps = d.p, p = ps + 1, n = d.n;
    const double nn = (double)n;
    // |z|^2 of the un-shifted covariate: dqrdc2 compares the remaining norm with the ORIGINAL column norm
    const double zn2 = fma(nn * v.z0, v.z0, fma(2.0 * v.z0, v.sz, v.zz));
    const bool all_zero = (v.zz == 0.0 && v.z0 == 0.0);
    if (all_zero) {
        // exact zero on R's diagonal: dpotri info = p -> the reference aborts the call (:551-557)
        if (status) atomicMax(status, p);
        const double qnan = __longlong_as_double(0x7ff8000000000000ULL);
        for (int c = 0; c < 2 * p + 3; ++c) out[(int64_t)c * ldo] = qnan;
        return;
    }
    double ee = 0.0, ff = 0.0, fe = 0.0;
#pragma unroll
    for (int k = 0; k < KS; ++k) {
        ee = fma(v.e[k], v.e[k], ee);
        ff = fma(v.f[k], v.f[k], ff);
        fe = fma(v.f[k], v.e[k], fe);
    }
    double d2 = v.zz - ff, g = v.zy - fe;
    double rss_s = v.yy - ee;                 // residual sum of squares of the static fit
    double rss = rss_s - (d2 > 0.0 ? g * g / d2 : 0.0);
    // un-shifted projections
    double e[KS], f[KS];
    double s_e = 0.0, s_f = 0.0;
#pragma unroll
    for (int k = 0; k < KS; ++k) {
        e[k] = fma(v.y0, d.q1[k], v.e[k]);
        f[k] = fma(v.z0, d.q1[k], v.f[k]);
        s_e = fma(d.q1[k], e[k], s_e);
        s_f = fma(d.q1[k], f[k], s_f);
    }
    bool exact = false;
    CovExact ex;
    if (v.zz > 0.0 && (d2 <= kExactThreshold * v.zz || (v.yy > 0.0 && rss <= kExactThreshold * v.yy))) {
        ex = exact_cov<KS>(y, z, ld, n, Qt, v.y0, v.z0, zrow);
        exact = true;
        d2 = ex.d2;
        g = ex.g;
        rss = ex.rss;
        rss_s = ex.rss_s;
    }
    // dqrdc2: the last column is negligible when its remaining norm < tol * original norm (tol = 1e-7)
    const bool negl = !(d2 > 1e-14 * zn2);
    double bz = 0.0, mss;
    if (negl) {
        rss = rss_s;
        if (exact) {
            mss = ex.mss_s;
        } else {
            const double m = s_e * d.inv_n;
            mss = 0.0;
#pragma unroll
            for (int k = 0; k < KS; ++k) {
                const double dk = fma(-d.q1[k], m, e[k]);
                mss = fma(dk, dk, mss);
            }
            mss = fma(d.gap * m, m, mss);
        }
        // R's diagonal is rounding noise there; its size in the reference is ~eps |z|
        d2 = fmax(d2, 1e-32 * zn2);
    } else {
        bz = g / d2;
        if (exact) {
            mss = ex.mss;
        } else {
            // fitted = Q e + u ep, u = z_perp / d: sum(fitted) = q1.e + (1'u) ep
            const double dd = sqrt(d2);
            const double ep = g / dd;
            const double q1p = d.shift ? 0.0 : ((v.sz - s_f) / dd);
            const double m = (s_e + q1p * ep) * d.inv_n;
            mss = 0.0;
#pragma unroll
            for (int k = 0; k < KS; ++k) {
                const double dk = fma(-d.q1[k], m, e[k]);
                mss = fma(dk, dk, mss);
            }
            const double dp = fma(-q1p, m, ep);
            mss = fma(dp, dp, mss);
            mss = fma(fmax(d.gap - q1p * q1p, 0.0) * m, m, mss);
        }
    }
    const double rdf = (double)(n - p);
    const double resvar = rss / rdf;
    out[0] = (mss / (double)(p - 1)) / resvar;
    out[ldo] = mss / (mss + rss);
    out[(int64_t)(2 * p + 2) * ldo] = d.mhalf_n * (d.loglik_c + log(rss));
    const double inv_d2 = 1.0 / d2;
#pragma unroll
    for (int i = 0; i < KS; ++i) {
        if (i < ps) {
            double be = 0.0, h = 0.0;
#pragma unroll
            for (int j = i; j < KS; ++j) {
                be = fma(d.Rinv[i * KS + j], e[j], be);
                h = fma(d.Rinv[i * KS + j], f[j], h);
            }
            const double b = fma(-h, bz, be);
            const double c = fma(h * h, inv_d2, d.ct[i]);
            out[(int64_t)(2 + i) * ldo] = b;
            out[(int64_t)(2 + p + i) * ldo] = b / sqrt(c * resvar);
        }
    }
    out[(int64_t)(2 + ps) * ldo] = bz;
    out[(int64_t)(2 + p + ps) * ldo] = negl ? 0.0 : bz / sqrt(inv_d2 * resvar);
}

// blockDim = (TX, S): thread (x, s) accumulates subjects [s n/S, (s+1) n/S) of VPT adjacent voxels; slices are
// reduced through shared memory and slice 0 runs the epilogue (same scheme as lm_fit_ldg_kernel).
template <int KS, int VPT>
__global__ void __launch_bounds__(256)
lm_cov_kernel(CovArgs a, int JT)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ SmemDesign<KS> sd;
    double *sQ = reinterpret_cast<double *>(smem_raw);
    const int TX = blockDim.x, S = blockDim.y;
    const int x = threadIdx.x, s = threadIdx.y;
    const int tid = s * TX + x, nthreads = TX * S;
    const int n = a.n;

    load_design_to_smem<KS>(sd, a.design, tid, nthreads);
    __syncthreads();

    const int64_t v = ((int64_t)blockIdx.x * TX + x) * VPT;
    bool in[VPT], act[VPT];
    bool any = false;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        in[i] = v + i < a.V;
        act[i] = in[i] && (!a.mask || mask_selects(a.mask[v + i], a.mask_lo, a.mask_hi));
        any |= act[i];
    }
    CovSums<KS> sum[VPT];
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
#pragma unroll
        for (int k = 0; k < KS; ++k) sum[i].e[k] = sum[i].f[k] = 0.0;
        sum[i].yy = sum[i].zz = sum[i].zy = sum[i].sz = 0.0;
        sum[i].y0 = (sd.shift && act[i]) ? a.Y[v + i] : 0.0;
        sum[i].z0 = (sd.shift && act[i]) ? a.Z[(a.zrow ? (int64_t)__ldg(a.zrow) * a.ldY : 0) + v + i] : 0.0;   // first (sampled) subject
    }
    const int jb = (int)(((int64_t)s * n) / S), je = (int)(((int64_t)(s + 1) * n) / S);
    const int span = (n + S - 1) / S;
    const double *yp = a.Y + (any ? v : 0);
    const double *zp = a.Z + (any ? v : 0);

    for (int t0 = 0; t0 < span; t0 += JT) {
        __syncthreads();
        for (int sl = 0; sl < S; ++sl) {
            const int b = (int)(((int64_t)sl * n) / S) + t0;
            const int e = (int)(((int64_t)(sl + 1) * n) / S);
            for (int i = tid; i < JT * KS; i += nthreads) {
                const int j = b + i / KS;
                sQ[(size_t)sl * JT * KS + i] = j < e ? a.Qt[(size_t)j * KS + i % KS] : 0.0;
            }
        }
        __syncthreads();
        if (!any) continue;
        const double *q = sQ + (size_t)s * JT * KS;
        const int cnt = min(JT, je - (jb + t0));
        int jj = 0;
        constexpr int U = 4;
        for (; jj + U <= cnt; jj += U) {
            double y[U][VPT], z[U][VPT];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int j = jb + t0 + jj + u;
                const int64_t off = (int64_t)j * a.ldY, offz = a.zrow ? (int64_t)__ldg(a.zrow + j) * a.ldY : off;
                if constexpr (VPT == 2) {
                    const double2 ty = ldg_stream2(yp + off), tz = ldg_stream2(zp + offz);
                    y[u][0] = ty.x; y[u][1] = ty.y;
                    z[u][0] = tz.x; z[u][1] = tz.y;
                } else {
                    y[u][0] = ldg_stream(yp + off);
                    z[u][0] = ldg_stream(zp + offz);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
#pragma unroll
                for (int i = 0; i < VPT; ++i) {
                    const double yv = y[u][i] - sum[i].y0, zv = z[u][i] - sum[i].z0;
                    sum[i].yy = fma(yv, yv, sum[i].yy);
                    sum[i].zz = fma(zv, zv, sum[i].zz);
                    sum[i].zy = fma(zv, yv, sum[i].zy);
                    sum[i].sz += zv;
#pragma unroll
                    for (int k = 0; k < KS; ++k) {
                        const double qk = q[(jj + u) * KS + k];
                        sum[i].e[k] = fma(qk, yv, sum[i].e[k]);
                        sum[i].f[k] = fma(qk, zv, sum[i].f[k]);
                    }
                }
            }
        }
        for (; jj < cnt; ++jj) {
            const int j = jb + t0 + jj;
            const int64_t off = (int64_t)j * a.ldY, offz = a.zrow ? (int64_t)__ldg(a.zrow + j) * a.ldY : off;
#pragma unroll
            for (int i = 0; i < VPT; ++i) {
                const double yv = yp[off + i] - sum[i].y0, zv = zp[offz + i] - sum[i].z0;
                sum[i].yy = fma(yv, yv, sum[i].yy);
                sum[i].zz = fma(zv, zv, sum[i].zz);
                sum[i].zy = fma(zv, yv, sum[i].zy);
                sum[i].sz += zv;
#pragma unroll
                for (int k = 0; k < KS; ++k) {
                    const double qk = q[jj * KS + k];
                    sum[i].e[k] = fma(qk, yv, sum[i].e[k]);
                    sum[i].f[k] = fma(qk, zv, sum[i].f[k]);
                }
            }
        }
    }

    if (S > 1) {
        // cross-slice reduction; partial layout [slice][component][voxel-in-block]
        __syncthreads();
        double *red = reinterpret_cast<double *>(smem_raw);
        constexpr int NC = 2 * KS + 4;
        const int W = TX * VPT;
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            double *r = red + (size_t)s * NC * W + x * VPT + i;
#pragma unroll
            for (int k = 0; k < KS; ++k) {
                r[(size_t)k * W] = sum[i].e[k];
                r[(size_t)(KS + k) * W] = sum[i].f[k];
            }
            r[(size_t)(2 * KS) * W] = sum[i].yy;
            r[(size_t)(2 * KS + 1) * W] = sum[i].zz;
            r[(size_t)(2 * KS + 2) * W] = sum[i].zy;
            r[(size_t)(2 * KS + 3) * W] = sum[i].sz;
        }
        __syncthreads();
        if (s != 0) return;
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            for (int sl = 1; sl < S; ++sl) {
                const double *r = red + (size_t)sl * NC * W + x * VPT + i;
#pragma unroll
                for (int k = 0; k < KS; ++k) {
                    sum[i].e[k] += r[(size_t)k * W];
                    sum[i].f[k] += r[(size_t)(KS + k) * W];
                }
                sum[i].yy += r[(size_t)(2 * KS) * W];
                sum[i].zz += r[(size_t)(2 * KS + 1) * W];
                sum[i].zy += r[(size_t)(2 * KS + 2) * W];
                sum[i].sz += r[(size_t)(2 * KS + 3) * W];
            }
        }
    }

    const int ncol = 2 * (sd.p + 1) + 3;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        if (!in[i]) continue;
        double *o = a.out + v + i;
        if (!act[i]) {
            zero_row<KS>(ncol, o, a.ldOut);
            continue;
        }
        finish_cov<KS>(sum[i], sd, a.Y + v + i, a.Z + v + i, a.ldY, a.Qt, o, a.ldOut, a.status, a.zrow);
    }
}

namespace {

template <int KS>
cudaError_t launch_cov_ks(const CovArgs &a, const LmPlan &plan, cudaStream_t st, LmLaunchInfo *info)
{
    const bool vec2 = ((reinterpret_cast<uintptr_t>(a.Y) & 15) == 0) && ((reinterpret_cast<uintptr_t>(a.Z) & 15) == 0) &&
                      (a.ldY % 2 == 0) && (a.V % 2 == 0);
    int S = plan.split;
    if (S <= 0) {
        S = 1;
        const int64_t want = (int64_t)plan.sm_count * 2048;
        while (S < 32 && (a.V / (vec2 ? 2 : 1)) * S < want && a.n / (S * 2) >= 32) S *= 2;
    }
    if (S > 32) S = 32;
    const int TX = S == 1 ? 128 : (256 / S < 8 ? 8 : 256 / S);
    const int VPT = vec2 ? 2 : 1;
    const int span = (a.n + S - 1) / S;
    int JT = span;
    while ((size_t)S * JT * KS * 8 > 64 * 1024) JT = (JT + 1) / 2;
    size_t smem = (size_t)S * JT * KS * 8;
    const size_t red = (size_t)S * (2 * KS + 4) * VPT * TX * 8;
    if (S > 1 && red > smem) smem = red;
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    const int64_t per_block = (int64_t)TX * VPT;
    const int64_t blocks = (a.V + per_block - 1) / per_block;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidValue;
    dim3 bd(TX, S);
    auto go = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<(unsigned)blocks, bd, smem, st>>>(a, JT);
        return cudaGetLastError();
    };
    const cudaError_t e = VPT == 2 ? go(lm_cov_kernel<KS, 2>) : go(lm_cov_kernel<KS, 1>);
    if (info) { info->launches++; info->used_tma = 0; info->split = S; }
    return e;
}

}  // namespace

cudaError_t launch_lm_cov(const CovArgs &a, const LmPlan &plan, cudaStream_t st, LmLaunchInfo *info)
{
    if (a.V <= 0) return cudaSuccess;
    switch (lm_padded_p(a.ps)) {
#define RMG_CASE(K) case K: return launch_cov_ks<K>(a, plan, st, info);
        RMG_CASE(1) RMG_CASE(2) RMG_CASE(3) RMG_CASE(4) RMG_CASE(5) RMG_CASE(6) RMG_CASE(7) RMG_CASE(8)
        RMG_CASE(12) RMG_CASE(16)
#undef RMG_CASE
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace rmg
