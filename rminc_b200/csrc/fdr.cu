// fdr.cu -- mincFDR / vertexFDR / anatFDR on the device, sm_100a.
//
// Reference: mincFDR.mincMultiDim, R/minc_FDR.R:254-520 -- per result column
//   t columns:  pvals = pt2(t, df) = 2 * pt(-|t|, df)                       (R/minc_FDR.R:392-397)
//   F columns:  pvals = pf(F, df1, df2, lower.tail = FALSE)                  (:398-412)
//   qvalue = p.adjust(pvals, "fdr")  (Benjamini-Hochberg over the voxels with mask > 0.5; NA p-values stay NA
//            and do not count)                                               (:432-434)
//   thresholds at q <= 0.01, 0.05, 0.10, 0.15, 0.20: the statistic whose p-value is the largest one still
//            accepted: qt(max p / 2, df, lower.tail = FALSE) / qf(max p, df1, df2, lower.tail = FALSE), NA if none
//                                                                            (:444-475)
//   output: 1 outside the mask                                               (:369, :505)
// pt / pf are R's (nmath, TOMS 708 incomplete beta -- in libR, not in the reference tree); restated here through
// the regularised incomplete beta function I_x(a, b) evaluated by its continued fraction (modified Lentz).
//
// Device pipeline per column: p-value kernel (one thread per voxel, compacting selected non-NaN voxels) ->
// radix sort of (p, voxel) pairs (CUB, library code) -> m/j * p_(j), reverse running minimum -> scatter q-values.
#include <cuda_runtime.h>

#include <cmath>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "capi_common.hpp"

namespace rmg {

namespace {

// Continued fraction of the incomplete beta function (modified Lentz).
__host__ __device__ inline double beta_cf(double a, double b, double x)
{
    const double tiny = 1e-300, eps = 1e-16;
    const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (fabs(d) < tiny) d = tiny;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m <= 5000; ++m) {
        const double m2 = 2.0 * m;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c;
        if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < tiny) d = tiny;
        c = 1.0 + aa / c;
        if (fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < eps) break;
    }
    return h;
}

// I_x(a, b) with y = 1 - x supplied by the caller (computed without cancellation).
__host__ __device__ inline double reg_inc_beta(double a, double b, double x, double y)
{
    if (!(x > 0.0)) return 0.0;
    if (!(y > 0.0)) return 1.0;
    const double lbeta = lgamma(a) + lgamma(b) - lgamma(a + b);
    const double front = exp(a * log(x) + b * log(y) - lbeta);
    if (x < (a + 1.0) / (a + b + 2.0)) return front * beta_cf(a, b, x) / a;
    return 1.0 - front * beta_cf(b, a, y) / b;
}

// pt2(t, df) = 2 * pt(-|t|, df) = I_{df / (df + t^2)}(df / 2, 1 / 2)
__host__ __device__ inline double p_two_sided_t(double t, double df)
{
    if (t != t) return t;
    if (isinf(t)) return 0.0;
    const double t2 = t * t, den = df + t2;
    return reg_inc_beta(0.5 * df, 0.5, df / den, t2 / den);
}

// pf(F, df1, df2, lower.tail = FALSE) = I_{df2 / (df2 + df1 F)}(df2 / 2, df1 / 2)
__host__ __device__ inline double p_upper_f(double f, double df1, double df2)
{
    if (f != f) return f;
    if (f <= 0.0) return 1.0;
    if (isinf(f)) return 0.0;
    const double den = df2 + df1 * f;
    return reg_inc_beta(0.5 * df2, 0.5 * df1, df2 / den, df1 * f / den);
}

__host__ __device__ inline double p_of_stat(double s, int type, double df1, double df2)
{
    return type == RMG_STAT_T ? p_two_sided_t(s, df1) : p_upper_f(s, df1, df2);
}

constexpr int kBlock = 256;

// p-value of every selected voxel; selected voxels with a non-NaN p are counted per block (compaction pass 1)
__global__ void __launch_bounds__(kBlock)
fdr_pvalue_kernel(const double *__restrict__ stat, int64_t V, int type, double df1, double df2, const double *__restrict__ mask,
                  double *__restrict__ p, unsigned int *__restrict__ counts)
{
    const int64_t v = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    bool keep = false;
    if (v < V) {
        const bool sel = !mask || mask[v] > 0.5;                 // R/minc_FDR.R:394 (mask > 0.5, not the lm mask rule)
        double pv = -1.0;                                         // marks "outside the mask"
        if (sel) {
            pv = p_of_stat(stat[v], type, df1, df2);
            keep = pv == pv;
        }
        p[v] = pv;
    }
    const int c = __syncthreads_count(keep);
    if (threadIdx.x == 0) counts[blockIdx.x] = (unsigned int)c;
}

__global__ void __launch_bounds__(kBlock)
fdr_compact_kernel(const double *__restrict__ p, int64_t V, const unsigned int *__restrict__ offsets, double *__restrict__ keys,
                   unsigned int *__restrict__ vals)
{
    __shared__ unsigned int warp_sums[kBlock / 32];
    const int64_t v = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const double pv = v < V ? p[v] : -1.0;
    const bool keep = pv >= 0.0;                                  // false for NaN and for the -1 marker
    const unsigned int ballot = __ballot_sync(0xffffffffu, keep);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_sums[warp] = __popc(ballot);
    __syncthreads();
    unsigned int before = 0;
    for (int w = 0; w < warp; ++w) before += warp_sums[w];
    if (keep) {
        const unsigned int pos = offsets[blockIdx.x] + before + __popc(ballot & ((1u << lane) - 1u));
        keys[pos] = pv;
        vals[pos] = (unsigned int)v;
    }
}

// w[m-1-j] = m / (j+1) * p_(j)   (sorted ascending, j 0-based): reversed so that a forward running minimum is
// p.adjust's cummin over decreasing p
__global__ void fdr_weight_kernel(const double *__restrict__ sorted_p, unsigned int m, double *__restrict__ w)
{
    const unsigned int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m) w[m - 1 - j] = ((double)m / (double)(j + 1)) * sorted_p[j];
}

struct MinOp {
    __host__ __device__ double operator()(double a, double b) const { return a < b ? a : b; }
};

// q of every voxel: 1 outside the mask, NaN where p is NaN, else min(1, running minimum)
__global__ void fdr_fill_kernel(const double *__restrict__ p, int64_t V, double *__restrict__ q)
{
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v < V) q[v] = p[v] < 0.0 ? 1.0 : p[v];                   // NaN stays NaN; selected entries are overwritten below
}

__global__ void fdr_scatter_kernel(const double *__restrict__ cummin_rev, const unsigned int *__restrict__ sorted_v, unsigned int m,
                                   double *__restrict__ q, double *__restrict__ q_sorted)
{
    const unsigned int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m) {
        const double qq = fmin(1.0, cummin_rev[m - 1 - j]);
        q[sorted_v[j]] = qq;
        q_sorted[j] = qq;
    }
}

// largest sorted position whose q is <= level (q_sorted is non-decreasing): one thread per level, bisection
__global__ void fdr_threshold_kernel(const double *__restrict__ q_sorted, const double *__restrict__ sorted_p, unsigned int m,
                                     const double *__restrict__ levels, int nlevels, double *__restrict__ pmax)
{
    const int l = threadIdx.x;
    if (l >= nlevels) return;
    unsigned int lo = 0, hi = m;                                  // first position with q > level
    while (lo < hi) {
        const unsigned int mid = lo + (hi - lo) / 2;
        if (q_sorted[mid] <= levels[l]) lo = mid + 1;
        else hi = mid;
    }
    pmax[l] = lo == 0 ? -1.0 : sorted_p[lo - 1];
}

// exclusive scan of the block counts (one block; V / 256 entries)
__global__ void __launch_bounds__(1024)
fdr_scan_kernel(unsigned int *__restrict__ counts, int nblocks, unsigned int *__restrict__ total)
{
    __shared__ unsigned int part[1024];
    __shared__ unsigned int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned int x = i < nblocks ? counts[i] : 0u;
        part[threadIdx.x] = x;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const unsigned int y = threadIdx.x >= o ? part[threadIdx.x - o] : 0u;
            __syncthreads();
            part[threadIdx.x] += y;
            __syncthreads();
        }
        if (i < nblocks) counts[i] = carry + part[threadIdx.x] - x;
        __syncthreads();
        if (threadIdx.x == 1023) carry += part[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

// statistic whose p-value is `p`: qt(p/2, df, lower = FALSE) / qf(p, df1, df2, lower = FALSE), by bisection on the
// monotone p-value function (host; five scalars per column)
double stat_of_p(double p, int type, double df1, double df2)
{
    if (!(p >= 0.0)) return std::nan("");
    if (p >= 1.0) return 0.0;
    if (p == 0.0) return INFINITY;
    double lo = 0.0, hi = 1.0;
    while (p_of_stat(hi, type, df1, df2) > p && hi < 1e300) hi *= 2.0;
    for (int it = 0; it < 400 && hi - lo > 1e-15 * hi; ++it) {
        const double mid = 0.5 * (lo + hi);
        if (p_of_stat(mid, type, df1, df2) > p) lo = mid;
        else hi = mid;
    }
    return 0.5 * (lo + hi);
}

constexpr int kLevels = 5;
const double kLevelValues[kLevels] = {0.01, 0.05, 0.10, 0.15, 0.20};   // R/minc_FDR.R:370

struct FdrWorkspace {
    double *p = nullptr, *keys = nullptr, *keys2 = nullptr, *w = nullptr, *qs = nullptr, *levels = nullptr, *pmax = nullptr;
    unsigned int *vals = nullptr, *vals2 = nullptr, *counts = nullptr, *total = nullptr;
    void *tmp = nullptr;
    cudaStream_t st = nullptr;
    template <class P>
    cudaError_t get(P *&ptr, size_t bytes) { return cudaMallocAsync(reinterpret_cast<void **>(&ptr), bytes ? bytes : 16, st); }
    void release()
    {
        void *all[] = {p, keys, keys2, w, qs, levels, pmax, vals, vals2, counts, total, tmp};
        for (void *x : all)
            if (x) cudaFreeAsync(x, st);
    }
};

}  // namespace

// dstat / dmask / dq on the device; thresholds (host, 5 doubles) written after a stream synchronisation.
int fdr_device_impl(const double *dstat, int64_t V, int stat_type, double df1, double df2, const double *dmask, double *dq,
                    double *thresholds, cudaStream_t st, char *err, size_t errlen)
{
    int rc = RMG_OK;
    FdrWorkspace ws;
    ws.st = st;
    if (V > 0xfffffff0LL) return fail(err, errlen, RMG_ERR_INVALID, "FDR supports at most 2^32 voxels per call");
    if (V == 0) {
        if (thresholds)
            for (int l = 0; l < kLevels; ++l) thresholds[l] = std::nan("");
        return RMG_OK;
    }
    cudaGetLastError();   // CUB reports any stale (non-sticky) runtime error as its own: start clean
    const int nblocks = (int)((V + kBlock - 1) / kBlock);
    unsigned int m = 0;
    double pmax_h[kLevels];
    size_t tmp_bytes = 0, tmp_scan = 0;
    {
        RMG_CUDA(ws.get(ws.p, (size_t)V * 8));
        RMG_CUDA(ws.get(ws.counts, (size_t)nblocks * 4));
        RMG_CUDA(ws.get(ws.total, 4));
        RMG_CUDA(ws.get(ws.levels, kLevels * 8));
        RMG_CUDA(ws.get(ws.pmax, kLevels * 8));
        fdr_pvalue_kernel<<<nblocks, kBlock, 0, st>>>(dstat, V, stat_type, df1, df2, dmask, ws.p, ws.counts);
        fdr_scan_kernel<<<1, 1024, 0, st>>>(ws.counts, nblocks, ws.total);
        fdr_fill_kernel<<<nblocks, kBlock, 0, st>>>(ws.p, V, dq);
        g_launch_count += 3;
        RMG_CUDA(cudaMemcpyAsync(&m, ws.total, 4, cudaMemcpyDeviceToHost, st));
        RMG_CUDA(cudaStreamSynchronize(st));
        for (int l = 0; l < kLevels; ++l) pmax_h[l] = -1.0;
        if (m > 0) {
            RMG_CUDA(ws.get(ws.keys, (size_t)m * 8));
            RMG_CUDA(ws.get(ws.keys2, (size_t)m * 8));
            RMG_CUDA(ws.get(ws.vals, (size_t)m * 4));
            RMG_CUDA(ws.get(ws.vals2, (size_t)m * 4));
            RMG_CUDA(ws.get(ws.w, (size_t)m * 8));
            RMG_CUDA(ws.get(ws.qs, (size_t)m * 8));
            fdr_compact_kernel<<<nblocks, kBlock, 0, st>>>(ws.p, V, ws.counts, ws.keys, ws.vals);
            RMG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, ws.keys, ws.keys2, ws.vals, ws.vals2, (int)m, 0, 64, st));
            RMG_CUDA(cub::DeviceScan::InclusiveScan(nullptr, tmp_scan, ws.w, ws.w, MinOp(), (int)m, st));
            if (tmp_scan > tmp_bytes) tmp_bytes = tmp_scan;
            RMG_CUDA(ws.get(ws.tmp, tmp_bytes));
            RMG_CUDA(cub::DeviceRadixSort::SortPairs(ws.tmp, tmp_bytes, ws.keys, ws.keys2, ws.vals, ws.vals2, (int)m, 0, 64, st));
            const unsigned int mb = (m + 255) / 256;
            fdr_weight_kernel<<<mb, 256, 0, st>>>(ws.keys2, m, ws.w);
            RMG_CUDA(cub::DeviceScan::InclusiveScan(ws.tmp, tmp_bytes, ws.w, ws.w, MinOp(), (int)m, st));
            fdr_scatter_kernel<<<mb, 256, 0, st>>>(ws.w, ws.vals2, m, dq, ws.qs);
            RMG_CUDA(cudaMemcpyAsync(ws.levels, kLevelValues, kLevels * 8, cudaMemcpyHostToDevice, st));
            fdr_threshold_kernel<<<1, 32, 0, st>>>(ws.qs, ws.keys2, m, ws.levels, kLevels, ws.pmax);
            g_launch_count += 6;
            RMG_CUDA(cudaMemcpyAsync(pmax_h, ws.pmax, kLevels * 8, cudaMemcpyDeviceToHost, st));
            RMG_CUDA(cudaStreamSynchronize(st));
        }
        if (thresholds)
            for (int l = 0; l < kLevels; ++l)
                thresholds[l] = pmax_h[l] < 0.0 ? std::nan("")
                                                : stat_of_p(stat_type == RMG_STAT_T ? pmax_h[l] : pmax_h[l], stat_type, df1, df2);
    }
cleanup:
    ws.release();
    if (rc != RMG_OK) cudaGetLastError();
    return rc;
}

}  // namespace rmg

using namespace rmg;

extern "C" {

static int check_fdr_args(const void *stat, int64_t V, int stat_type, double df1, double df2, const void *q, char *err, size_t errlen)
{
    if (V < 0 || (V > 0 && (!stat || !q))) return fail(err, errlen, RMG_ERR_INVALID, "stat or qvalues is NULL");
    if (stat_type != RMG_STAT_T && stat_type != RMG_STAT_F)
        return fail(err, errlen, RMG_ERR_INVALID, "stat_type must be RMG_STAT_T or RMG_STAT_F (got %d)", stat_type);
    if (!(df1 > 0.0) || (stat_type == RMG_STAT_F && !(df2 > 0.0)))
        return fail(err, errlen, RMG_ERR_INVALID, "Error: need to specify the degrees of freedom");   // R/minc_FDR.R:344
    return RMG_OK;
}

int rmg_fdr_device(const double *dstat, int64_t V, int stat_type, double df1, double df2, const double *dmask, double *dq,
                   double *thresholds, int device, void *stream, char *err, size_t errlen)
{
    int rc = check_fdr_args(dstat, V, stat_type, df1, df2, dq, err, errlen);
    if (rc) return rc;
    if ((rc = select_device(device, err, errlen))) return rc;
    return fdr_device_impl(dstat, V, stat_type, df1, df2, dmask, dq, thresholds, static_cast<cudaStream_t>(stream), err, errlen);
}

int rmg_fdr(const double *stat, int64_t V, int stat_type, double df1, double df2, const double *mask, double *qvalues,
            double *thresholds, int device, char *err, size_t errlen)
{
    int rc = check_fdr_args(stat, V, stat_type, df1, df2, qvalues, err, errlen);
    if (rc) return rc;
    if ((rc = select_device(device, err, errlen))) return rc;
    cudaStream_t st = nullptr;
    double *ds = nullptr, *dm = nullptr, *dq = nullptr;
    {
        RMG_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&ds), (size_t)std::max<int64_t>(V, 1) * 8, st));
        RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&dq), (size_t)std::max<int64_t>(V, 1) * 8, st));
        RMG_CUDA(cudaMemcpyAsync(ds, stat, (size_t)V * 8, cudaMemcpyHostToDevice, st));
        if (mask) {
            RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&dm), (size_t)std::max<int64_t>(V, 1) * 8, st));
            RMG_CUDA(cudaMemcpyAsync(dm, mask, (size_t)V * 8, cudaMemcpyHostToDevice, st));
        }
        rc = fdr_device_impl(ds, V, stat_type, df1, df2, dm, dq, thresholds, st, err, errlen);
        if (rc == RMG_OK) {
            RMG_CUDA(cudaMemcpyAsync(qvalues, dq, (size_t)V * 8, cudaMemcpyDeviceToHost, st));
            RMG_CUDA(cudaStreamSynchronize(st));
        }
    }
cleanup:
    if (ds) cudaFreeAsync(ds, st);
    if (dm) cudaFreeAsync(dm, st);
    if (dq) cudaFreeAsync(dq, st);
    if (st) {
        cudaStreamSynchronize(st);
        cudaStreamDestroy(st);
    }
    if (rc != RMG_OK) cudaGetLastError();
    return rc;
}

/* p-value of one statistic (host; the same code the device runs) -- exposed for tests and for R-side consumers
 * that want pt2 / pf of a threshold. */
double rmg_pvalue(double stat, int stat_type, double df1, double df2) { return p_of_stat(stat, stat_type, df1, df2); }

}  // extern "C"
