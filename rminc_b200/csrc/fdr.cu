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
//   method = "BY":  p.adjust(pvals, "BY") = pmin(1, cummin(c * n / i * p)), c = sum(1 / (1:n))   (:435-437)
//   method = "pFDR" / "fastqvalue":  fast.qvalue (R/minc_misc.R:366-560): q = pi0 * m * p / rank (rank with ties =
//            number of p <= p_i), then the reference's own qvalue_min (src/modelling_functions.c:15-46), which -- as
//            written -- is a pairwise minimum with the next-ranked RAW value, not a running minimum, and is
//            restated literally.  pi0 comes from the caller: its default estimate is stats::smooth.spline (R, third
//            party) over mean(p >= lambda), which rmg_fdr_pi0_fractions supplies.
// pt / pf are R's (nmath, TOMS 708 incomplete beta -- in libR, not in the reference tree); restated here through
// the regularised incomplete beta function I_x(a, b) evaluated by its continued fraction (modified Lentz).
//
// Device pipeline per column: p-value kernel (one thread per voxel, compacting selected non-NaN voxels) ->
// radix sort of (p, voxel) pairs (CUB, library code) -> m/j * p_(j), reverse running minimum -> scatter q-values.
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
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

// one predicate for "this p-value takes part" in both compaction passes: false for NaN and for the -1 marker
__device__ __forceinline__ bool fdr_keeps(double pv) { return pv >= 0.0; }

// p-value of every selected voxel; selected voxels with a non-NaN p are counted per block (compaction pass 1)
__global__ void __launch_bounds__(kBlock)
fdr_pvalue_kernel(const double *__restrict__ stat, int64_t V, int type, double df1, double df2, const double *__restrict__ mask,
                  double *__restrict__ p, unsigned int *__restrict__ counts, unsigned int *__restrict__ nan_count)
{
    const int64_t v = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    bool keep = false;
    if (v < V) {
        const bool sel = !mask || mask[v] > 0.5;                 // R/minc_FDR.R:394 (mask > 0.5, not the lm mask rule)
        double pv = -1.0;                                         // marks "outside the mask"
        if (sel) {
            pv = p_of_stat(stat[v], type, df1, df2);
            if (pv == pv) pv = fmin(fmax(pv, 0.0), 1.0);          // [0, 1] exactly (no -0.0, no 1 + ulp)
            keep = fdr_keeps(pv);
            if (!keep) atomicAdd(nan_count, 1u);                  // a selected voxel with a NaN statistic (rare)
        }
        p[v] = pv;
    }
    const int c = __syncthreads_count(keep);
    if (threadIdx.x == 0) counts[blockIdx.x] = (unsigned int)c;
}

__global__ void __launch_bounds__(kBlock)
fdr_compact_kernel(const double *__restrict__ p, int64_t V, const unsigned int *__restrict__ offsets, double *__restrict__ keys,
                   unsigned int *__restrict__ vals, unsigned int *__restrict__ vox)
{
    __shared__ unsigned int warp_sums[kBlock / 32];
    const int64_t v = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const double pv = v < V ? p[v] : -1.0;
    const bool keep = fdr_keeps(pv);
    const unsigned int ballot = __ballot_sync(0xffffffffu, keep);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_sums[warp] = __popc(ballot);
    __syncthreads();
    unsigned int before = 0;
    for (int w = 0; w < warp; ++w) before += warp_sums[w];
    if (keep) {
        const unsigned int pos = offsets[blockIdx.x] + before + __popc(ballot & ((1u << lane) - 1u));
        keys[pos] = pv;
        vals[pos] = pos;                   // compacted position (order of the masked p vector R sees)
        vox[pos] = (unsigned int)v;
    }
}

// w[m-1-j] = num / (j+1) * p_(j)   (sorted ascending, j 0-based): reversed so that a forward running minimum is
// p.adjust's cummin over decreasing p.  num = n for "fdr" (n / i * p), c * n for "BY" (q * n / i * p, left to right).
__global__ void fdr_weight_kernel(const double *__restrict__ sorted_p, unsigned int m, double num, double *__restrict__ w)
{
    const unsigned int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m) w[m - 1 - j] = (num / (double)(j + 1)) * sorted_p[j];
}

// fast.qvalue's raw q of sorted position j: pi0 * m * p / v, v = number of p <= p_j (qvalue.rank, R/minc_misc.R:505-516)
__device__ __forceinline__ double storey_raw_q(const double *__restrict__ sorted_p, unsigned int m, double pi0, double pj)
{
    unsigned int lo = 0, hi = m;                                  // first position with p > pj
    while (lo < hi) {
        const unsigned int mid = lo + (hi - lo) / 2;
        if (sorted_p[mid] <= pj) lo = mid + 1;
        else hi = mid;
    }
    return pi0 * (double)m * pj / (double)lo;
}

__global__ void fdr_storey_raw_kernel(const double *__restrict__ sorted_p, unsigned int m, double pi0, double *__restrict__ qraw)
{
    const unsigned int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m) qraw[j] = storey_raw_q(sorted_p, m, pi0, sorted_p[j]);
}

// qvalue_min as the reference wrote it (src/modelling_functions.c:15-46), in sorted coordinates:
//   j < m-1:  both q_j and q_{j+1} above 1 -> 1, else the smaller of the two RAW values
//   j = m-1:  the reference tests qvalue[u[m-1]] with the 1-based index used 0-based, i.e. the raw q of the element that
//             FOLLOWS the largest-p element in the (masked) input order; above 1 -> 1, else the raw value.  When the
//             largest p is the last element the reference reads past the end of the vector: treated as "not above 1".
__global__ void fdr_storey_min_kernel(const double *__restrict__ qraw, const double *__restrict__ sorted_p,
                                      const unsigned int *__restrict__ sorted_k, const double *__restrict__ compact_p, unsigned int m,
                                      double pi0, const unsigned int *__restrict__ vox, double *__restrict__ q, double *__restrict__ q_sorted)
{
    const unsigned int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    double out;
    if (j + 1 < m) {
        const double a = qraw[j], b = qraw[j + 1];
        out = (a > 1.0 && b > 1.0) ? 1.0 : (a < b ? a : b);
    } else {
        const unsigned int k = sorted_k[j];
        bool above = false;
        if (k + 1 < m) above = storey_raw_q(sorted_p, m, pi0, compact_p[k + 1]) > 1.0;
        out = above ? 1.0 : qraw[j];
    }
    q[vox[sorted_k[j]]] = out;
    q_sorted[j] = out;
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

__global__ void fdr_scatter_kernel(const double *__restrict__ cummin_rev, const unsigned int *__restrict__ sorted_k,
                                   const unsigned int *__restrict__ vox, unsigned int m, double mult, double *__restrict__ q,
                                   double *__restrict__ q_sorted)
{
    const unsigned int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < m) {
        const double qq = mult * fmin(1.0, cummin_rev[m - 1 - j]);   // mult: 1, or pi0 for the qvalue package's form
        q[vox[sorted_k[j]]] = qq;
        q_sorted[j] = qq;
    }
}

// max(p[q <= level]) for every level (R/minc_FDR.R:444-475).  A reduction, not a bisection: fast.qvalue's q-values are
// not monotone in p.  pmax_bits[l] holds the bit pattern of the largest accepted p plus one (0 = none; p >= 0 so the
// patterns order like the values).
__global__ void __launch_bounds__(256)
fdr_threshold_kernel(const double *__restrict__ q_sorted, const double *__restrict__ sorted_p, unsigned int m,
                     const double *__restrict__ levels, int nlevels, unsigned long long *__restrict__ pmax_bits)
{
    __shared__ unsigned long long red[8];
    for (int l = 0; l < nlevels; ++l) {
        const double level = levels[l];
        unsigned long long best = 0;
        for (unsigned int j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x)
            if (q_sorted[j] <= level) best = max(best, (unsigned long long)__double_as_longlong(sorted_p[j]) + 1ull);
        for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < 8; ++w) best = max(best, red[w]);
            if (best) atomicMax(&pmax_bits[l], best);
        }
        __syncthreads();
    }
}

// fast.qvalue's pi0 grid: count[i] = #{p >= lambda_i} over the participating voxels (R/minc_misc.R:439-441)
__global__ void __launch_bounds__(256)
fdr_lambda_count_kernel(const double *__restrict__ p, int64_t V, const double *__restrict__ lambda, int nlambda,
                        unsigned long long *__restrict__ counts)
{
    __shared__ unsigned int sc[64];
    if (threadIdx.x < 64) sc[threadIdx.x] = 0;
    __syncthreads();
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (int64_t)gridDim.x * blockDim.x) {
        const double pv = p[v];
        if (!fdr_keeps(pv)) continue;
        for (int i = 0; i < nlambda; ++i)
            if (pv >= lambda[i]) atomicAdd(&sc[i], 1u);
    }
    __syncthreads();
    if (threadIdx.x < nlambda && sc[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)sc[threadIdx.x]);
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
    unsigned int *vals = nullptr, *vals2 = nullptr, *counts = nullptr, *total = nullptr, *vox = nullptr;
    unsigned long long *pmax_bits = nullptr;
    void *tmp = nullptr;
    cudaStream_t st = nullptr;
    template <class P>
    cudaError_t get(P *&ptr, size_t bytes) { return cudaMallocAsync(reinterpret_cast<void **>(&ptr), bytes ? bytes : 16, st); }
    void release()
    {
        void *all[] = {p, keys, keys2, w, qs, levels, pmax, vals, vals2, counts, total, tmp, vox, pmax_bits};
        for (void *x : all)
            if (x) cudaFreeAsync(x, st);
    }
};

}  // namespace

// c = sum(1 / (1:m)) as R's sum() forms it: accumulated in long double, rounded once (p.adjust "BY")
static double harmonic_number(unsigned int m)
{
    long double s = 0.0L;
    for (unsigned int i = 1; i <= m; ++i) s += (long double)(1.0 / (double)i);
    return (double)s;
}

// p-values of the selected voxels (ws.p; -1 outside the mask), block counts scanned, m and the number of selected
// voxels whose p is NaN read back.  Shared by the q-value pipeline and the pi0 grid.
static int fdr_pvalues(FdrWorkspace &ws, const double *dstat, int64_t V, int stat_type, double df1, double df2, const double *dmask,
                       int nblocks, unsigned int &m, unsigned int &n_nan, char *err, size_t errlen)
{
    int rc = RMG_OK;
    cudaStream_t st = ws.st;
    unsigned int h[2] = {0, 0};
    {
        RMG_CUDA(ws.get(ws.p, (size_t)V * 8));
        RMG_CUDA(ws.get(ws.counts, (size_t)nblocks * 4));
        RMG_CUDA(ws.get(ws.total, 8));
        RMG_CUDA(cudaMemsetAsync(ws.total, 0, 8, st));
        fdr_pvalue_kernel<<<nblocks, kBlock, 0, st>>>(dstat, V, stat_type, df1, df2, dmask, ws.p, ws.counts, ws.total + 1);
        fdr_scan_kernel<<<1, 1024, 0, st>>>(ws.counts, nblocks, ws.total);
        g_launch_count += 2;
        RMG_CUDA(cudaMemcpyAsync(h, ws.total, 8, cudaMemcpyDeviceToHost, st));
        RMG_CUDA(cudaStreamSynchronize(st));
        m = h[0];
        n_nan = h[1];
    }
cleanup:
    return rc;
}

// dstat / dmask / dq on the device; thresholds (host, 5 doubles) written after a stream synchronisation.
int fdr_device_impl(const double *dstat, int64_t V, int stat_type, double df1, double df2, const double *dmask, double *dq,
                    double *thresholds, int method, double pi0, cudaStream_t st, char *err, size_t errlen)
{
    int rc = RMG_OK;
    FdrWorkspace ws;
    ws.st = st;
    // the sort and the scan take int item counts, voxel numbers travel as unsigned int
    if (V > 0x7fffffffLL) return fail(err, errlen, RMG_ERR_INVALID, "FDR supports at most 2^31 - 1 voxels per call");
    const bool storey = method == RMG_FDR_STOREY, scaled = method == RMG_FDR_QVALUE;
    if ((storey || scaled) && !(pi0 > 0.0 && pi0 <= 1.0))
        return fail(err, errlen, RMG_ERR_INVALID,
                    "ERROR: The estimated pi0 <= 0. Check that you have valid p-values or use another lambda method.");
    if (V == 0) {
        if (thresholds)
            for (int l = 0; l < kLevels; ++l) thresholds[l] = std::nan("");
        return RMG_OK;
    }
    cudaGetLastError();   // CUB reports any stale (non-sticky) runtime error as its own: start clean
    const int nblocks = (int)((V + kBlock - 1) / kBlock);
    unsigned int m = 0, n_nan = 0;
    unsigned long long pmax_h[kLevels] = {0, 0, 0, 0, 0};
    size_t tmp_bytes = 0, tmp_scan = 0;
    {
        if ((rc = fdr_pvalues(ws, dstat, V, stat_type, df1, df2, dmask, nblocks, m, n_nan, err, errlen))) goto cleanup;
        // fast.qvalue / qvalue test min(p) < 0 || max(p) > 1 first: an NA p-value is R's "missing value" error
        if ((storey || scaled) && n_nan > 0) {
            rc = fail(err, errlen, RMG_ERR_INVALID, "missing value where TRUE/FALSE needed (%u selected voxels have a NaN statistic)", n_nan);
            goto cleanup;
        }
        fdr_fill_kernel<<<nblocks, kBlock, 0, st>>>(ws.p, V, dq);
        g_launch_count += 1;
        if (m > 0) {
            RMG_CUDA(ws.get(ws.keys, (size_t)m * 8));
            RMG_CUDA(ws.get(ws.keys2, (size_t)m * 8));
            RMG_CUDA(ws.get(ws.vals, (size_t)m * 4));
            RMG_CUDA(ws.get(ws.vals2, (size_t)m * 4));
            RMG_CUDA(ws.get(ws.vox, (size_t)m * 4));
            RMG_CUDA(ws.get(ws.w, (size_t)m * 8));
            RMG_CUDA(ws.get(ws.qs, (size_t)m * 8));
            RMG_CUDA(ws.get(ws.levels, kLevels * 8));
            RMG_CUDA(ws.get(ws.pmax_bits, kLevels * 8));
            RMG_CUDA(cudaMemsetAsync(ws.pmax_bits, 0, kLevels * 8, st));
            fdr_compact_kernel<<<nblocks, kBlock, 0, st>>>(ws.p, V, ws.counts, ws.keys, ws.vals, ws.vox);
            RMG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, ws.keys, ws.keys2, ws.vals, ws.vals2, (int)m, 0, 64, st));
            RMG_CUDA(cub::DeviceScan::InclusiveScan(nullptr, tmp_scan, ws.w, ws.w, MinOp(), (int)m, st));
            if (tmp_scan > tmp_bytes) tmp_bytes = tmp_scan;
            RMG_CUDA(ws.get(ws.tmp, tmp_bytes));
            // stable sort: equal p-values keep the order of the masked vector, like R's order(p)
            RMG_CUDA(cub::DeviceRadixSort::SortPairs(ws.tmp, tmp_bytes, ws.keys, ws.keys2, ws.vals, ws.vals2, (int)m, 0, 64, st));
            const unsigned int mb = (m + 255) / 256;
            if (storey) {
                fdr_storey_raw_kernel<<<mb, 256, 0, st>>>(ws.keys2, m, pi0, ws.w);
                fdr_storey_min_kernel<<<mb, 256, 0, st>>>(ws.w, ws.keys2, ws.vals2, ws.keys, m, pi0, ws.vox, dq, ws.qs);
                g_launch_count += 1;
            } else {
                const double num = method == RMG_FDR_BY ? harmonic_number(m) * (double)m : (double)m;
                fdr_weight_kernel<<<mb, 256, 0, st>>>(ws.keys2, m, num, ws.w);
                RMG_CUDA(cub::DeviceScan::InclusiveScan(ws.tmp, tmp_bytes, ws.w, ws.w, MinOp(), (int)m, st));
                fdr_scatter_kernel<<<mb, 256, 0, st>>>(ws.w, ws.vals2, ws.vox, m, scaled ? pi0 : 1.0, dq, ws.qs);
                g_launch_count += 2;
            }
            RMG_CUDA(cudaMemcpyAsync(ws.levels, kLevelValues, kLevels * 8, cudaMemcpyHostToDevice, st));
            fdr_threshold_kernel<<<(unsigned)std::min<unsigned int>(mb, 592u), 256, 0, st>>>(ws.qs, ws.keys2, m, ws.levels, kLevels, ws.pmax_bits);
            g_launch_count += 4;
            RMG_CUDA(cudaMemcpyAsync(pmax_h, ws.pmax_bits, kLevels * 8, cudaMemcpyDeviceToHost, st));
            RMG_CUDA(cudaStreamSynchronize(st));
        }
        if (thresholds)
            for (int l = 0; l < kLevels; ++l) {
                if (pmax_h[l] == 0) { thresholds[l] = std::nan(""); continue; }
                const unsigned long long bits = pmax_h[l] - 1ull;
                double pm;
                std::memcpy(&pm, &bits, 8);
                thresholds[l] = stat_of_p(pm, stat_type, df1, df2);
            }
    }
cleanup:
    ws.release();
    if (rc != RMG_OK) cudaGetLastError();
    return rc;
}

// fractions[i] = mean(p >= lambda[i]) over the participating voxels (fast.qvalue's pi0 grid, R/minc_misc.R:439-441)
int fdr_pi0_device_impl(const double *dstat, int64_t V, int stat_type, double df1, double df2, const double *dmask,
                        const double *lambda, int nlambda, double *fractions, cudaStream_t st, char *err, size_t errlen)
{
    int rc = RMG_OK;
    FdrWorkspace ws;
    ws.st = st;
    if (V > 0x7fffffffLL) return fail(err, errlen, RMG_ERR_INVALID, "FDR supports at most 2^31 - 1 voxels per call");
    if (nlambda < 1 || nlambda > 64 || !lambda || !fractions)
        return fail(err, errlen, RMG_ERR_INVALID, "lambda must hold 1..64 values (got %d)", nlambda);
    for (int i = 0; i < nlambda; ++i)
        if (!(lambda[i] >= 0.0 && lambda[i] < 1.0)) return fail(err, errlen, RMG_ERR_INVALID, "ERROR: Lambda must be within [0, 1).");
    cudaGetLastError();
    const int nblocks = (int)((std::max<int64_t>(V, 1) + kBlock - 1) / kBlock);
    unsigned int m = 0, n_nan = 0;
    unsigned long long hc[64];
    double *dl = nullptr;
    unsigned long long *dc = nullptr;
    {
        if ((rc = fdr_pvalues(ws, dstat, V, stat_type, df1, df2, dmask, nblocks, m, n_nan, err, errlen))) goto cleanup;
        if (n_nan > 0) {
            rc = fail(err, errlen, RMG_ERR_INVALID, "missing value where TRUE/FALSE needed (%u selected voxels have a NaN statistic)", n_nan);
            goto cleanup;
        }
        RMG_CUDA(ws.get(dl, 64 * 8));
        RMG_CUDA(ws.get(dc, 64 * 8));
        RMG_CUDA(cudaMemsetAsync(dc, 0, 64 * 8, st));
        RMG_CUDA(cudaMemcpyAsync(dl, lambda, (size_t)nlambda * 8, cudaMemcpyHostToDevice, st));
        fdr_lambda_count_kernel<<<std::min(nblocks, 1184), 256, 0, st>>>(ws.p, V, dl, nlambda, dc);
        g_launch_count += 1;
        RMG_CUDA(cudaMemcpyAsync(hc, dc, (size_t)nlambda * 8, cudaMemcpyDeviceToHost, st));
        RMG_CUDA(cudaStreamSynchronize(st));
        for (int i = 0; i < nlambda; ++i) fractions[i] = m ? (double)hc[i] / (double)m : std::nan("");
    }
cleanup:
    if (dl) cudaFreeAsync(dl, st);
    if (dc) cudaFreeAsync(dc, st);
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

static int check_fdr_method(int method, char *err, size_t errlen)
{
    if (method < RMG_FDR_BH || method > RMG_FDR_QVALUE)
        return fail(err, errlen, RMG_ERR_INVALID, "method must be RMG_FDR_BH, _BY, _STOREY or _QVALUE (got %d)", method);
    return RMG_OK;
}

int rmg_fdr_method_device(const double *dstat, int64_t V, int stat_type, double df1, double df2, const double *dmask, double *dq,
                          double *thresholds, int method, double pi0, int device, void *stream, char *err, size_t errlen)
{
    int rc = check_fdr_args(dstat, V, stat_type, df1, df2, dq, err, errlen);
    if (rc || (rc = check_fdr_method(method, err, errlen))) return rc;
    if ((rc = select_device(device, err, errlen))) return rc;
    return fdr_device_impl(dstat, V, stat_type, df1, df2, dmask, dq, thresholds, method, pi0, static_cast<cudaStream_t>(stream), err, errlen);
}

int rmg_fdr_device(const double *dstat, int64_t V, int stat_type, double df1, double df2, const double *dmask, double *dq,
                   double *thresholds, int device, void *stream, char *err, size_t errlen)
{
    return rmg_fdr_method_device(dstat, V, stat_type, df1, df2, dmask, dq, thresholds, RMG_FDR_BH, 1.0, device, stream, err, errlen);
}

}  // extern "C"

// stat (and mask) to the device on a fresh stream; `run` does the work there
template <class F>
static int fdr_with_host_input(const double *stat, int64_t V, const double *mask, double *qvalues, int device, char *err, size_t errlen, F run)
{
    int rc = select_device(device, err, errlen);
    if (rc) return rc;
    cudaStream_t st = nullptr;
    double *ds = nullptr, *dm = nullptr, *dq = nullptr;
    {
        RMG_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&ds), (size_t)std::max<int64_t>(V, 1) * 8, st));
        RMG_CUDA(cudaMemcpyAsync(ds, stat, (size_t)V * 8, cudaMemcpyHostToDevice, st));
        if (qvalues) RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&dq), (size_t)std::max<int64_t>(V, 1) * 8, st));
        if (mask) {
            RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&dm), (size_t)std::max<int64_t>(V, 1) * 8, st));
            RMG_CUDA(cudaMemcpyAsync(dm, mask, (size_t)V * 8, cudaMemcpyHostToDevice, st));
        }
        rc = run(ds, dm, dq, st);
        if (rc == RMG_OK && qvalues) {
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

extern "C" {

int rmg_fdr_method(const double *stat, int64_t V, int stat_type, double df1, double df2, const double *mask, double *qvalues,
                   double *thresholds, int method, double pi0, int device, char *err, size_t errlen)
{
    int rc = check_fdr_args(stat, V, stat_type, df1, df2, qvalues, err, errlen);
    if (rc || (rc = check_fdr_method(method, err, errlen))) return rc;
    return fdr_with_host_input(stat, V, mask, qvalues, device, err, errlen, [&](double *ds, double *dm, double *dq, cudaStream_t st) {
        return fdr_device_impl(ds, V, stat_type, df1, df2, dm, dq, thresholds, method, pi0, st, err, errlen);
    });
}

int rmg_fdr(const double *stat, int64_t V, int stat_type, double df1, double df2, const double *mask, double *qvalues,
            double *thresholds, int device, char *err, size_t errlen)
{
    return rmg_fdr_method(stat, V, stat_type, df1, df2, mask, qvalues, thresholds, RMG_FDR_BH, 1.0, device, err, errlen);
}

int rmg_fdr_pi0_fractions_device(const double *dstat, int64_t V, int stat_type, double df1, double df2, const double *dmask,
                                 const double *lambda, int nlambda, double *fractions, int device, void *stream, char *err, size_t errlen)
{
    int rc = check_fdr_args(dstat, V, stat_type, df1, df2, dstat, err, errlen);
    if (rc) return rc;
    if ((rc = select_device(device, err, errlen))) return rc;
    return fdr_pi0_device_impl(dstat, V, stat_type, df1, df2, dmask, lambda, nlambda, fractions, static_cast<cudaStream_t>(stream), err, errlen);
}

int rmg_fdr_pi0_fractions(const double *stat, int64_t V, int stat_type, double df1, double df2, const double *mask,
                          const double *lambda, int nlambda, double *fractions, int device, char *err, size_t errlen)
{
    int rc = check_fdr_args(stat, V, stat_type, df1, df2, stat, err, errlen);
    if (rc) return rc;
    return fdr_with_host_input(stat, V, mask, nullptr, device, err, errlen, [&](double *ds, double *dm, double *, cudaStream_t st) {
        return fdr_pi0_device_impl(ds, V, stat_type, df1, df2, dm, lambda, nlambda, fractions, st, err, errlen);
    });
}

/* p-value of one statistic (host; the same code the device runs) -- exposed for tests and for R-side consumers
 * that want pt2 / pf of a threshold. */
double rmg_pvalue(double stat, int stat_type, double df1, double df2) { return p_of_stat(stat, stat_type, df1, df2); }

}  // extern "C"
