// randomize.cu -- mincRandomize on the device (SURVEY.md pieces K3/K4/K5).
//
// Reference: R/minc_voxel_statistics.R:1465-1497.  Each permutation r refits every voxel
// with X_pi = X[idx[, r], ] and the same Y and mask, sets non-finite statistics to 0
// (:1478), takes abs() when two-sided (:1484-1486) and records the column maxima over ALL
// V rows (:1488; masked rows are 0).  The reference re-runs the whole mincLm call,
// re-reading every file, once per permutation.
//
// Here every permuted design is factored on the host (microseconds each), Y stays resident
// in HBM, and the per-permutation maxima are reduced on the device into an R x ncols
// buffer with ordered-integer atomicMax.  Voxel shards on several GPUs are combined by the
// caller with an element-wise max (NCCL all-reduce(max) in rminc_b200/parallel.py, a host
// max in rmg_randomize_multi).
//
// Two execution paths produce the same numbers:
//   "refit": one streaming fit (lm_kernels.cu) per permutation into a scratch result
//            matrix + a column-max reduction.  HBM-bound: R passes over Y.
//   "gemm":  K permutations share each pass over Y -- stacked (K*p) x n by n x V FP64
//            tensor-core GEMM (mma.sync DMMA) with the statistics + max fused in the
//            epilogue (see perm_gemm.cuh).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <thread>

#include "capi_common.hpp"
#include "perm_gemm.cuh"
#include "tfce.hpp"

namespace rmg {
namespace {

__global__ void keys_init_kernel(unsigned long long *keys, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = ordered_key(-INFINITY);
}

__global__ void keys_decode_kernel(const unsigned long long *keys, double *dist, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dist[i] = ordered_value(keys[i]);
}

// Column maxima of one fitted result matrix (V x ncol_total, ld) for the selected columns,
// merged into keys[c * R + r].
__global__ void __launch_bounds__(256)
colmax_kernel(const double *__restrict__ out, int64_t V, int64_t ld, const int32_t *__restrict__ cols, int ncols,
              int two_sided, unsigned long long *__restrict__ keys, int R, int r)
{
    __shared__ double red[8];
    for (int c = 0; c < ncols; ++c) {
        const double *col = out + (int64_t)cols[c] * ld;
        double best = -INFINITY;
        for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (int64_t)gridDim.x * blockDim.x)
            best = fmax(best, perm_stat_filter(col[v], two_sided));
        for (int o = 16; o > 0; o >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, o));
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
        __syncthreads();
        if (threadIdx.x < 32) {
            double b = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -INFINITY;
            for (int o = 4; o > 0; o >>= 1) b = fmax(b, __shfl_xor_sync(0xffffffffu, b, o));
            if (threadIdx.x == 0) atomicMax(&keys[(size_t)c * R + r], ordered_key(b));
        }
        __syncthreads();
    }
}

// Factor every permuted design; translate the reference's error contract.
int factor_permutations(const double *X, int n, int p, const int32_t *idx, int R, PermSet &pd, char *err,
                        size_t errlen)
{
    for (size_t i = 0; i < (size_t)n * p; ++i)
        if (!std::isfinite(X[i])) return fail(err, errlen, RMG_ERR_INVALID, "model matrix contains non-finite values");
    pd.own.assign(R, Design{});
    pd.idx = idx;
    pd.R = R;
    for (int r = 0; r < R; ++r)
        for (int i = 0; i < n; ++i) {
            const int32_t src = idx[i + (size_t)r * n];
            if (src < 0 || src >= n) return fail(err, errlen, RMG_ERR_INVALID, "idx[%d, %d] = %d out of range", i, r, src);
        }
    // A column of idx that is a true permutation (replace = FALSE) only re-orders the rows of X:
    // X_pi = P X  =>  Q_pi = P Q with the same R, pivoting, rank and (R'R)^-1, so its factorisation is a
    // row gather of the base one.  Columns with repeated indices (replace = TRUE) change the column
    // space and are factored from scratch.  The work is independent per permutation: host threads.
    const int base_info = factor_design(X, n, p, pd.base);
    std::vector<int> infos(R, 0);
    parallel_for(R, [&](int r0, int r1) {
        std::vector<double> Xp((size_t)n * p);
        std::vector<unsigned char> seen(n);
        for (int r = r0; r < r1; ++r) {
            const int32_t *ix = idx + (size_t)r * n;
            std::fill(seen.begin(), seen.end(), 0);
            bool is_perm = base_info == 0;
            for (int i = 0; i < n && is_perm; ++i) {
                if (seen[ix[i]]) is_perm = false;
                seen[ix[i]] = 1;
            }
            if (is_perm) continue;   // own[r] stays empty: represented by the base factorisation + idx[, r]
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < p; ++j) Xp[i + (size_t)j * n] = X[ix[i] + (size_t)j * n];
            infos[r] = factor_design(Xp.data(), n, p, pd.own[r]);
        }
    });
    for (int r = 0; r < R; ++r) {   // report the FIRST failing permutation, as a sequential loop would
        if (infos[r] > 0)
            return fail(err, errlen, RMG_ERR_SINGULAR, "permutation %d: %s", r + 1, singular_message(infos[r]).c_str());
        pd.all_intercept = pd.all_intercept && pd.design(r).has_intercept;
    }
    return RMG_OK;
}

// t columns of one fitted result matrix -> maps[(slot * ncols + c) * V + v], non-finite -> 0
// (new_mod[!is.finite(new_mod)] <- 0, R/minc_voxel_statistics.R:1478)
__global__ void __launch_bounds__(256)
extract_maps_kernel(const double *__restrict__ out, int64_t V, int64_t ld, const int32_t *__restrict__ cols, int ncols,
                    int slot, double *__restrict__ maps)
{
    const int c = blockIdx.y;
    const int64_t v = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (v >= V) return;
    const double x = out[(int64_t)cols[c] * ld + v];
    maps[((size_t)slot * ncols + c) * V + v] = isfinite(x) ? x : 0.0;
}

// per-map maximum of the enhanced maps (abs first when two-sided, :1484-1488) -> keys[c * R + r]
__global__ void __launch_bounds__(256)
tfce_colmax_kernel(const double *__restrict__ tf, int64_t V, int ncols, int two_sided, unsigned long long *__restrict__ keys,
                   int R, int r0)
{
    __shared__ double red[8];
    const int j = blockIdx.y;                   // map index inside the batch: (r - r0) * ncols + c
    double best = -INFINITY;
    for (int64_t v = (int64_t)blockIdx.x * 256 + threadIdx.x; v < V; v += (int64_t)gridDim.x * 256) {
        const double x = tf[(size_t)j * V + v];
        best = fmax(best, two_sided ? fabs(x) : x);
    }
    for (int o = 16; o > 0; o >>= 1) best = fmax(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) best = fmax(best, red[w]);
        atomicMax(&keys[(size_t)(j % ncols) * R + (r0 + j / ncols)], ordered_key(best));
    }
}

enum class PermPath { Auto, Refit, Gemm };

PermPath perm_path_from_env()
{
    if (const char *v = std::getenv("RMG_PERM_PATH")) {
        if (!std::strcmp(v, "refit")) return PermPath::Refit;
        if (!std::strcmp(v, "gemm")) return PermPath::Gemm;
    }
    return PermPath::Auto;
}

// keys[s][i], s < nsets: element-wise max over the sets, decoded to doubles (the voxel blocks of one call each reduce
// into a set of their own because the GEMM path finalises its keys in place at the end of a block)
__global__ void keys_merge_decode_kernel(const unsigned long long *keys, int nsets, double *dist, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long k = keys[i];
    for (int s = 1; s < nsets; ++s) k = max(k, keys[(size_t)s * n + i]);
    dist[i] = ordered_value(k);
}

int check_perm_args(int p, const int32_t *idx, int R, const int32_t *cols, int ncols, const void *dist, char *err, size_t errlen)
{
    if (R < 0 || ncols < 0 || (R > 0 && !idx) || (ncols > 0 && !cols) || (R > 0 && ncols > 0 && !dist))
        return fail(err, errlen, RMG_ERR_INVALID, "bad permutation arguments");
    for (int c = 0; c < ncols; ++c)
        if (cols[c] < 0 || cols[c] >= 2 * p + 3) return fail(err, errlen, RMG_ERR_INVALID, "cols[%d] out of range", c);
    return RMG_OK;
}

// All R permutations of one RESIDENT block of voxels, reduced into keys[ncols][R] (which must hold key(-inf)).
// Everything is queued on st; the host needs nothing of the block afterwards.
int randomize_block(const PermSet &pd, const double *dY, int64_t V, int64_t ldY, int n, int p, const double *dmask,
                    double mask_lo, double mask_hi, const int32_t *cols, const int32_t *dcols, int ncols, int two_sided,
                    unsigned long long *keys, int R, int sm_count, cudaStream_t st, int64_t &launches, char *err,
                    size_t errlen)
{
    int rc = RMG_OK;
    if (V <= 0) return rc;                     // nothing to reduce: the keys keep key(-inf), max over no voxels
    double *scratch = nullptr;
    void *gemm_ws = nullptr;
    LmScratch scr;
    LmPlan plan;
    plan.sm_count = sm_count;
    const PermPath path = perm_path_from_env();
    {
        bool done = false;
        if (path != PermPath::Refit && V > 0 && perm_gemm_supported(n, p, V, ldY, dY, cols, ncols)) {
            PermGemmArgs ga{dY, V, ldY, n, p, dmask, mask_lo, mask_hi, cols, ncols, two_sided, keys, R};
            int gl = 0;
            cudaError_t ge = perm_gemm_run(ga, pd, plan.sm_count, st, &gl, &gemm_ws);
            launches += gl;
            if (ge != cudaSuccess) {
                rc = fail(err, errlen, RMG_ERR_CUDA, "permutation GEMM failed: %s", cudaGetErrorString(ge));
                goto cleanup;
            }
            done = true;
        } else if (path == PermPath::Gemm) {
            rc = fail(err, errlen, RMG_ERR_INVALID, "RMG_PERM_PATH=gemm but the GEMM path does not support this shape");
            goto cleanup;
        }
        if (!done) {
            // refit path: one streaming fit per permutation + column maxima
            const int ncol_total = 2 * p + 3;
            const int64_t ldo = (V + 1) & ~int64_t(1);
            RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&scratch), (size_t)ldo * ncol_total * sizeof(double), st));
            const int KP = lm_padded_p(p);
            DevDesign *dd = nullptr;
            double *qt = nullptr;
            RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&dd), sizeof(DevDesign) * (size_t)R, st));
            const size_t qstride = make_qt(pd.base, KP).size();   // rows padded for the TMA kernel
            RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&qt), sizeof(double) * (size_t)R * qstride, st));
            {
                std::vector<DevDesign> hdd(R);
                std::vector<double> hqt((size_t)R * qstride);
                for (int r = 0; r < R; ++r) {
                    const Design Dr = pd.materialise(r);
                    fill_dev_design(Dr, hdd[r]);
                    const std::vector<double> q = make_qt(Dr, KP);
                    std::copy(q.begin(), q.end(), hqt.begin() + (size_t)r * qstride);
                }
                cudaError_t e1 = cudaMemcpyAsync(dd, hdd.data(), sizeof(DevDesign) * (size_t)R, cudaMemcpyHostToDevice, st);
                cudaError_t e2 = cudaMemcpyAsync(qt, hqt.data(), sizeof(double) * hqt.size(), cudaMemcpyHostToDevice, st);
                // large pageable sources are not staged: wait before the host vectors die
                cudaError_t e3 = cudaStreamSynchronize(st);
                if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
                    cudaFreeAsync(dd, st);
                    cudaFreeAsync(qt, st);
                    rc = fail(err, errlen, RMG_ERR_CUDA, "uploading permuted designs failed");
                    goto cleanup;
                }
            }
            const int cm_grid = (int)std::min<int64_t>((V + 255) / 256, (int64_t)plan.sm_count * 8);
            cudaError_t le = cudaSuccess;
            for (int r = 0; r < R && le == cudaSuccess; ++r) {
                LmArgs a{dY, V, ldY, n, p, qt + (size_t)r * qstride, dd + r, dmask, mask_lo, mask_hi, scratch, ldo};
                LmLaunchInfo info;
                le = launch_lm_fit(a, plan, scr, st, &info);
                launches += info.launches;
                if (le != cudaSuccess) break;
                colmax_kernel<<<cm_grid, 256, 0, st>>>(scratch, V, ldo, dcols, ncols, two_sided, keys, R, r);
                ++launches;
                le = cudaGetLastError();
            }
            cudaFreeAsync(dd, st);
            cudaFreeAsync(qt, st);
            if (le != cudaSuccess) {
                rc = fail(err, errlen, RMG_ERR_CUDA, "permutation refit failed: %s", cudaGetErrorString(le));
                goto cleanup;
            }
        }
    }
cleanup:
    if (scratch) cudaFreeAsync(scratch, st);
    if (gemm_ws) cudaFreeAsync(gemm_ws, st);
    scr.release(st);
    if (rc != RMG_OK) cudaGetLastError();
    return rc;
}

}  // namespace
}  // namespace rmg

using namespace rmg;

extern "C" {

int rmg_randomize_device(const double *dY, int64_t V, int64_t ldY, int n, const double *X, int p,
                         const double *dmask, double mask_lo, double mask_hi, const int32_t *idx, int R,
                         const int32_t *cols, int ncols, int two_sided, double *ddist, int device, void *stream,
                         char *err, size_t errlen)
{
    int rc = check_fit_args(dY, V, ldY, n, X, p, dY, ldY, err, errlen);
    if (rc) return rc;
    if ((rc = check_perm_args(p, idx, R, cols, ncols, ddist, err, errlen))) return rc;
    PermSet pd;
    const auto t_f0 = std::chrono::steady_clock::now();
    if ((rc = factor_permutations(X, n, p, idx, R, pd, err, errlen))) return rc;
    if (std::getenv("RMG_TIMING"))
        std::fprintf(stderr, "[rmg perm] %-28s %8.3f ms\n", "factor R designs",
                     std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_f0).count());
    if ((rc = select_device(device, err, errlen))) return rc;
    if (R == 0 || ncols == 0) return RMG_OK;

    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int nk = R * ncols;
    unsigned long long *keys = nullptr;
    int32_t *dcols = nullptr;
    int64_t launches = 0;
    {
        RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&keys), (size_t)nk * sizeof(unsigned long long), st));
        RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&dcols), (size_t)ncols * sizeof(int32_t), st));
        RMG_CUDA(cudaMemcpyAsync(dcols, cols, (size_t)ncols * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        keys_init_kernel<<<(nk + 255) / 256, 256, 0, st>>>(keys, nk);
        ++launches;
        if ((rc = randomize_block(pd, dY, V, ldY, n, p, dmask, mask_lo, mask_hi, cols, dcols, ncols, two_sided, keys, R,
                                  device_sm_count(device), st, launches, err, errlen)))
            goto cleanup;
        keys_merge_decode_kernel<<<(nk + 255) / 256, 256, 0, st>>>(keys, 1, ddist, nk);
        ++launches;
        RMG_CUDA(cudaGetLastError());
    }
cleanup:
    g_launch_count += launches;
    if (keys) cudaFreeAsync(keys, st);
    if (dcols) cudaFreeAsync(dcols, st);
    if (rc != RMG_OK) cudaGetLastError();
    return rc;
}

// vertexTFCE.vertexLm's randomisation (R/minc_vertex_statistics.R:892-948): mincRandomize_core with
// post_proc = vertexTFCE.matrix -- per permutation refit, non-finite -> 0, graph TFCE of every requested column, abs when
// two-sided, column maximum.  Batches of permutations are fitted into a map stack and enhanced together.
// mask (nullable; volumes): masked voxels are 0 in every statistic map, as in mincLm's result.  step > 0: thresholds
// from a fixed height increment (mincTFCE's d) instead of nsteps.
static int randomize_tfce_impl(const double *Y, int64_t V, int64_t ldY, int n, const double *X, int p, const double *mask,
                               double mask_lo, double mask_hi, const int32_t *idx, int R, const int32_t *cols, int ncols,
                               int two_sided, const HostTfceGraph &hg, double E, double H, int nsteps, double step, int side,
                               double *dist, int device, char *err, size_t errlen)
{
    if (step > 0.0) nsteps = kTfceStepLevels;
    int rc = check_fit_args(Y, V, ldY, n, X, p, Y, ldY, err, errlen);
    if (rc) return rc;
    if (R < 0 || ncols < 0 || (R > 0 && !idx) || (ncols > 0 && !cols) || (R > 0 && ncols > 0 && !dist))
        return fail(err, errlen, RMG_ERR_INVALID, "bad permutation arguments");
    for (int c = 0; c < ncols; ++c)
        if (cols[c] < 0 || cols[c] >= 2 * p + 3) return fail(err, errlen, RMG_ERR_INVALID, "cols[%d] out of range", c);
    if (V > 0x7fffffffLL) return fail(err, errlen, RMG_ERR_INVALID, "bad graph arguments");
    if (side < 0 || side > 2 || nsteps < 1 || nsteps > 253) return fail(err, errlen, RMG_ERR_INVALID, "bad TFCE arguments");
    if ((rc = check_csr(hg, V, err, errlen))) return rc;
    PermSet pd;
    if ((rc = factor_permutations(X, n, p, idx, R, pd, err, errlen))) return rc;
    if ((rc = select_device(device, err, errlen))) return rc;
    if (R == 0 || ncols == 0 || V == 0) return RMG_OK;

    cudaStream_t st = nullptr;
    const int64_t ldd = std::max<int64_t>(2, (V + 1) & ~int64_t(1));
    const int ncol_total = 2 * p + 3, KP = lm_padded_p(p);
    double *dY = nullptr, *scratch = nullptr, *maps = nullptr, *tf = nullptr, *dw = nullptr, *qt = nullptr, *dD = nullptr;
    double *dM = nullptr;
    int *dptr = nullptr, *didx = nullptr;
    int32_t *dcols = nullptr;
    DevDesign *dd = nullptr;
    unsigned long long *keys = nullptr;
    void *ws = nullptr;
    LmScratch scr;
    LmPlan plan;
    plan.sm_count = device_sm_count(device);
    int launches = 0;
    TfceGraph g;
    PinnedRing ring;
    {
        // permutations per batch: keep the TFCE scratch under ~8 GiB
        const size_t per_map = tfce_workspace_bytes(V, 1, nsteps, side) + 2 * (size_t)V * 8;
        const int B = (int)std::max<size_t>(1, std::min<size_t>((size_t)R, (size_t(8) << 30) / (per_map * ncols)));
        const size_t qstride = make_qt(pd.base, KP).size();
        RMG_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        if ((rc = upload_graph(hg, V, st, g, &dptr, &didx, &dw, err, errlen))) goto cleanup;
        RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&dY), (size_t)ldd * n * 8, st));
        RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&scratch), (size_t)ldd * ncol_total * 8, st));
        RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&maps), (size_t)B * ncols * V * 8, st));
        RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&tf), (size_t)B * ncols * V * 8, st));
        RMG_CUDA(cudaMallocAsync(&ws, tfce_workspace_bytes(V, B * ncols, nsteps, side), st));
        RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&dcols), (size_t)ncols * 4, st));
        RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&keys), (size_t)R * ncols * 8, st));
        RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&dD), (size_t)R * ncols * 8, st));
        RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&dd), sizeof(DevDesign) * (size_t)B, st));
        RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&qt), sizeof(double) * (size_t)B * qstride, st));
        {   // pageable Y (R's heap) goes through the pinned ring
            int rows_per_slot = 1;
            const bool staged = staging_plan(Y, (double)V * n * 8.0, (size_t)V * 8, n, rows_per_slot);
            if (staged) RMG_CUDA(ring.init((size_t)rows_per_slot * (size_t)V * 8));
            RMG_CUDA(upload_rows(staged ? &ring : nullptr, rows_per_slot, reinterpret_cast<char *>(dY), (size_t)ldd * 8,
                                 reinterpret_cast<const char *>(Y), (size_t)ldY * 8, (size_t)V * 8, n, st));
        }
        RMG_CUDA(cudaMemcpyAsync(dcols, cols, (size_t)ncols * 4, cudaMemcpyHostToDevice, st));
        if (mask) {
            RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&dM), (size_t)V * 8, st));
            RMG_CUDA(cudaMemcpyAsync(dM, mask, (size_t)V * 8, cudaMemcpyHostToDevice, st));
        }
        keys_init_kernel<<<(R * ncols + 255) / 256, 256, 0, st>>>(keys, R * ncols);
        ++launches;
        std::vector<DevDesign> hdd(B);
        std::vector<double> hqt((size_t)B * qstride);
        const unsigned vb = (unsigned)((V + 255) / 256);
        for (int r0 = 0; r0 < R; r0 += B) {
            const int rb = std::min(B, R - r0);
            for (int i = 0; i < rb; ++i) {
                const Design Dr = pd.materialise(r0 + i);
                fill_dev_design(Dr, hdd[i]);
                const std::vector<double> q = make_qt(Dr, KP);
                std::copy(q.begin(), q.end(), hqt.begin() + (size_t)i * qstride);
            }
            RMG_CUDA(cudaMemcpyAsync(dd, hdd.data(), sizeof(DevDesign) * (size_t)rb, cudaMemcpyHostToDevice, st));
            RMG_CUDA(cudaMemcpyAsync(qt, hqt.data(), sizeof(double) * (size_t)rb * qstride, cudaMemcpyHostToDevice, st));
            RMG_CUDA(cudaStreamSynchronize(st));          // the host images are reused by the next batch
            for (int i = 0; i < rb; ++i) {
                LmArgs a{dY, V, ldd, n, p, qt + (size_t)i * qstride, dd + i, dM, mask_lo, mask_hi, scratch, ldd};
                LmLaunchInfo info;
                RMG_CUDA(launch_lm_fit(a, plan, scr, st, &info));
                launches += info.launches;
                extract_maps_kernel<<<dim3(vb, (unsigned)ncols), 256, 0, st>>>(scratch, V, ldd, dcols, ncols, i, maps);
                ++launches;
            }
            const cudaError_t te = tfce_batch(maps, V, V, rb * ncols, g, E, H, nsteps, side, tf, V, ws, st, &launches, step);
            if (te == cudaErrorNotSupported) {
                rc = fail(err, errlen, RMG_ERR_INVALID, "more than %d thresholds between 0 and a map maximum: increase d", kTfceStepLevels);
                goto cleanup;
            }
            RMG_CUDA(te);
            tfce_colmax_kernel<<<dim3(std::min<unsigned>(vb, 32u), (unsigned)(rb * ncols)), 256, 0, st>>>(tf, V, ncols, two_sided, keys, R, r0);
            ++launches;
            RMG_CUDA(cudaGetLastError());
        }
        keys_decode_kernel<<<(R * ncols + 255) / 256, 256, 0, st>>>(keys, dD, R * ncols);
        ++launches;
        RMG_CUDA(cudaMemcpyAsync(dist, dD, (size_t)R * ncols * 8, cudaMemcpyDeviceToHost, st));
        RMG_CUDA(cudaStreamSynchronize(st));
    }
cleanup:
    g_launch_count += launches;
    {
        void *all[] = {dY, scratch, maps, tf, ws, dptr, didx, dw, dcols, keys, dD, dd, qt, dM};
        for (void *x : all)
            if (x) cudaFreeAsync(x, st);
    }
    scr.release(st);
    if (st) {
        cudaStreamSynchronize(st);
        cudaStreamDestroy(st);
    }
    ring.release_all();
    if (rc != RMG_OK) cudaGetLastError();
    return rc;
}

int rmg_randomize_tfce(const double *Y, int64_t V, int64_t ldY, int n, const double *X, int p, const int32_t *idx, int R,
                       const int32_t *cols, int ncols, int two_sided, const int32_t *adj_ptr, const int32_t *adj_idx,
                       const double *weights, double E, double H, int nsteps, int side, double *dist, int device, char *err,
                       size_t errlen)
{
    HostTfceGraph hg;
    hg.adj_ptr = adj_ptr;
    hg.adj_idx = adj_idx;
    hg.weights = weights;
    return randomize_tfce_impl(Y, V, ldY, n, X, p, nullptr, 0.0, 0.0, idx, R, cols, ncols, two_sided, hg, E, H, nsteps, 0.0, side,
                               dist, device, err, errlen);
}

}  // extern "C"

// The TFCE permutation test shards over PERMUTATIONS (every permutation needs the whole surface / volume): GPU g takes
// a contiguous block of the R index vectors, one host thread per GPU, no exchange -- the rows of dist are independent.
// fn(g, idx_block, R_block, dist_block, err, errlen)
template <class Fn>
static int tfce_perm_sharded(int n, const int32_t *idx, int R, int ncols, double *dist, int n_gpus, char *err, size_t errlen, Fn fn)
{
    const int have = rmg_device_count();
    if (have <= 0) return fail(err, errlen, RMG_ERR_NO_DEVICE, "no CUDA device available; rminc_b200 has no CPU fallback");
    const int G = n_gpus <= 0 ? have : n_gpus;
    if (G > have) return fail(err, errlen, RMG_ERR_NO_DEVICE, "%d GPUs requested, %d visible", G, have);
    if (R < 0 || ncols < 0 || (R > 0 && (!idx || (ncols > 0 && !dist)))) return fail(err, errlen, RMG_ERR_INVALID, "bad permutation arguments");
    std::vector<int> rcs(G, RMG_OK);
    std::vector<std::string> msgs(G, std::string(256, '\0'));
    std::vector<std::vector<double>> part(G);
    std::vector<std::thread> th;
    for (int g = 0; g < G; ++g) {
        th.emplace_back([&, g]() {
            const int r0 = (int)((int64_t)R * g / G), r1 = (int)((int64_t)R * (g + 1) / G);
            if (r1 <= r0) return;
            part[g].assign((size_t)(r1 - r0) * std::max(ncols, 1), 0.0);
            rcs[g] = fn(g, idx + (size_t)r0 * n, r1 - r0, part[g].data(), &msgs[g][0], msgs[g].size());
        });
    }
    for (auto &t : th) t.join();
    for (int g = 0; g < G; ++g)
        if (rcs[g] != RMG_OK) return fail(err, errlen, rcs[g], "gpu %d: %s", g, msgs[g].c_str());
    for (int g = 0; g < G; ++g) {
        const int r0 = (int)((int64_t)R * g / G), r1 = (int)((int64_t)R * (g + 1) / G);
        for (int c = 0; c < ncols; ++c)
            for (int r = r0; r < r1; ++r) dist[(size_t)c * R + r] = part[g][(size_t)c * (r1 - r0) + (r - r0)];
    }
    return RMG_OK;
}

extern "C" {

int rmg_randomize_tfce_multi(const double *Y, int64_t V, int64_t ldY, int n, const double *X, int p, const int32_t *idx, int R,
                             const int32_t *cols, int ncols, int two_sided, const int32_t *adj_ptr, const int32_t *adj_idx,
                             const double *weights, double E, double H, int nsteps, int side, double *dist, int n_gpus,
                             char *err, size_t errlen)
{
    return tfce_perm_sharded(n, idx, R, ncols, dist, n_gpus, err, errlen,
                             [&](int g, const int32_t *ix, int Rb, double *part, char *e, size_t el) {
        return rmg_randomize_tfce(Y, V, ldY, n, X, p, ix, Rb, cols, ncols, two_sided, adj_ptr, adj_idx, weights, E, H, nsteps, side,
                                  part, g, e, el);
    });
}

// mincTFCE.mincLm's permutation test (R/minc_voxel_statistics.R:1382-1439: mincRandomize_core with
// post_proc = mincTFCE.matrix) on the voxel grid; n_gpus <= 1 runs on device 0.
int rmg_randomize_volume_tfce(const double *Y, int64_t V, int64_t ldY, int n, const double *X, int p, const double *mask,
                              double mask_lo, double mask_hi, const int32_t *idx, int R, const int32_t *cols, int ncols,
                              int two_sided, const int64_t *dims, int connectivity, double voxel_volume, double d, double E,
                              double H, int side, double *dist, int n_gpus, char *err, size_t errlen)
{
    if (!dims) return fail(err, errlen, RMG_ERR_INVALID, "dims is NULL");
    if (dims[0] * dims[1] * dims[2] != V) return fail(err, errlen, RMG_ERR_INVALID, "dims do not multiply to V");
    if (!(d > 0.0) || !std::isfinite(d)) return fail(err, errlen, RMG_ERR_INVALID, "d must be a positive step (got %g)", d);
    if (!(voxel_volume > 0.0) || !std::isfinite(voxel_volume)) return fail(err, errlen, RMG_ERR_INVALID, "voxel_volume must be > 0");
    if (rmg_device_count() <= 0) return fail(err, errlen, RMG_ERR_NO_DEVICE, "no CUDA device available; rminc_b200 has no CPU fallback");
    HostTfceGraph hg;
    int rc = volume_graph(dims, connectivity, voxel_volume, hg, err, errlen);
    if (rc) return rc;
    if (n_gpus <= 1)
        return randomize_tfce_impl(Y, V, ldY, n, X, p, mask, mask_lo, mask_hi, idx, R, cols, ncols, two_sided, hg, E, H, 0, d, side,
                                   dist, 0, err, errlen);
    return tfce_perm_sharded(n, idx, R, ncols, dist, n_gpus, err, errlen,
                             [&](int gpu, const int32_t *ix, int Rb, double *part, char *e, size_t el) {
        return randomize_tfce_impl(Y, V, ldY, n, X, p, mask, mask_lo, mask_hi, ix, Rb, cols, ncols, two_sided, hg, E, H, 0, d, side,
                                   part, gpu, e, el);
    });
}

// mincRandomize on host data.  Y (doubles, or the volumes in their stored type) is uploaded ONCE, in voxel blocks on a
// copy stream, and stays resident as doubles; every block runs all R permutations as soon as it has landed, so the
// upload of block b+1 hides behind the tensor-core work of block b.  Pageable input (R's heap) goes through the pinned
// staging ring.  Each block reduces into a key set of its own; the sets are max-merged at the end.
static int randomize_host_impl(const void *Yv, const StoredSpec *pk, int64_t V, int64_t ldY, int n, const double *X, int p,
                               const double *mask, double mask_lo, double mask_hi, const int32_t *idx, int R,
                               const int32_t *cols, int ncols, int two_sided, double *dist, int device, char *err,
                               size_t errlen)
{
    int rc = check_fit_args(Yv, V, ldY, n, X, p, Yv, ldY, err, errlen);
    if (rc) return rc;
    if ((rc = check_perm_args(p, idx, R, cols, ncols, dist, err, errlen))) return rc;
    if ((rc = select_device(device, err, errlen))) return rc;
    if (R == 0 || ncols == 0) return RMG_OK;
    PermSet pd;
    if ((rc = factor_permutations(X, n, p, idx, R, pd, err, errlen))) return rc;

    const size_t es = !pk ? sizeof(double) : pk->dtype == kStoredU16 ? 2 : 4;   // bytes per sample of the input
    const char *Yb = static_cast<const char *>(Yv);
    const int64_t ldd = std::max<int64_t>(8, (V + 7) & ~int64_t(7));
    // voxel blocks: about an eighth of the volume, at least ~1 GiB of doubles (every block re-packs and re-uploads the
    // permuted designs, ~5 ms of host work that hides behind the previous block), multiples of 1024 voxels
    int64_t CV = std::max<int64_t>((V + 7) / 8, (int64_t)(1024.0 * 1024 * 1024 / (8.0 * n)));
    CV = std::max<int64_t>(1024, ((CV + 1023) / 1024) * 1024);
    if (const char *s = std::getenv("RMG_PERM_CHUNK_VOXELS")) CV = std::max<int64_t>(2, (std::atoll(s) / 2) * 2);
    if (CV > V) CV = std::max<int64_t>(V, 1);
    const int nchunks = (int)std::max<int64_t>(1, (V + CV - 1) / CV);
    const int64_t slice_len = pk ? std::max<int64_t>(1, pk->slice_len) : 1;
    const int64_t nslices = !pk ? 0 : pk->nslices_total > 0 ? pk->nslices_total : (V + slice_len - 1) / slice_len;
    const int nk = R * ncols;
    const int sm_count = device_sm_count(device);

    double *dY = nullptr, *dM = nullptr, *dD = nullptr, *dScale = nullptr, *dOffset = nullptr;
    char *dU = nullptr;
    unsigned long long *keys = nullptr;
    int32_t *dcols = nullptr;
    cudaStream_t st = nullptr, up = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    std::vector<cudaEvent_t> landed(nchunks, nullptr);
    PinnedRing ring;
    int64_t launches = 0;
    g_last_kernel_ms = 0.0;
    int rows_per_slot = 1;
    const bool staged = V > 0 && staging_plan(Yv, (double)V * n * (double)es, (size_t)CV * es, n, rows_per_slot);
    {
        RMG_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        RMG_CUDA(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking));
        RMG_CUDA(cudaEventCreate(&e0));
        RMG_CUDA(cudaEventCreate(&e1));
        if (staged) RMG_CUDA(ring.init((size_t)rows_per_slot * (size_t)CV * es));
        RMG_CUDA(cudaMalloc(&dY, (size_t)ldd * n * sizeof(double)));
        RMG_CUDA(cudaMalloc(&dD, (size_t)nk * sizeof(double)));
        RMG_CUDA(cudaMalloc(&keys, (size_t)nchunks * nk * sizeof(unsigned long long)));
        RMG_CUDA(cudaMalloc(&dcols, (size_t)ncols * sizeof(int32_t)));
        RMG_CUDA(cudaMemcpyAsync(dcols, cols, (size_t)ncols * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        keys_init_kernel<<<(nchunks * nk + 255) / 256, 256, 0, st>>>(keys, nchunks * nk);
        ++launches;
        if (pk) {
            RMG_CUDA(cudaMalloc(&dU, (size_t)CV * n * es));          // one block of stored samples, rows of CV
            if (pk->dtype == kStoredU16) {
                const size_t cnt = (size_t)nslices * n * sizeof(double);
                RMG_CUDA(cudaMalloc(&dScale, cnt));
                RMG_CUDA(cudaMalloc(&dOffset, cnt));
                RMG_CUDA(cudaMemcpyAsync(dScale, pk->scale, cnt, cudaMemcpyHostToDevice, up));
                RMG_CUDA(cudaMemcpyAsync(dOffset, pk->offset, cnt, cudaMemcpyHostToDevice, up));
            }
        }
        if (mask && V > 0) {
            RMG_CUDA(cudaMalloc(&dM, (size_t)V * sizeof(double)));
            RMG_CUDA(cudaMemcpyAsync(dM, mask, (size_t)V * 8, cudaMemcpyHostToDevice, up));
        }
        RMG_CUDA(cudaEventRecord(e0, st));
        for (int c = 0; c < nchunks; ++c) {
            const int64_t v0 = (int64_t)c * CV, cv = std::min(CV, V - v0);
            if (cv > 0) {
                // ---- the block's samples: host -> device on the copy stream (straight into Y for doubles)
                char *dstdev = pk ? dU : reinterpret_cast<char *>(dY + v0);
                const size_t dpitch = pk ? (size_t)CV * es : (size_t)ldd * 8;
                RMG_CUDA(upload_rows(staged ? &ring : nullptr, rows_per_slot, dstdev, dpitch, Yb + (size_t)v0 * es, (size_t)ldY * es,
                                     (size_t)cv * es, n, up));
                // stored samples become the reader's doubles on the copy stream too: stream order then protects the
                // one staging block, and the expansion (~1 ms) is noise next to the block's transfer
                if (pk) {
                    RMG_CUDA(launch_expand_stored(dU, pk->dtype, cv, CV, n, dScale, dOffset, 4503599627370496.0 + pk->voxel_min,
                                                  slice_len, nslices, pk->v_offset + v0, dY + v0, ldd, up));
                    ++launches;
                }
            }
            RMG_CUDA(cudaEventCreateWithFlags(&landed[c], cudaEventDisableTiming));
            RMG_CUDA(cudaEventRecord(landed[c], up));
            RMG_CUDA(cudaStreamWaitEvent(st, landed[c], 0));
            // ---- all R permutations on the block
            if ((rc = randomize_block(pd, dY + v0, cv, ldd, n, p, mask ? dM + v0 : nullptr, mask_lo, mask_hi, cols, dcols, ncols,
                                      two_sided, keys + (size_t)c * nk, R, sm_count, st, launches, err, errlen)))
                goto cleanup;
        }
        keys_merge_decode_kernel<<<(nk + 255) / 256, 256, 0, st>>>(keys, nchunks, dD, nk);
        ++launches;
        RMG_CUDA(cudaGetLastError());
        RMG_CUDA(cudaEventRecord(e1, st));
        RMG_CUDA(cudaMemcpyAsync(dist, dD, (size_t)nk * sizeof(double), cudaMemcpyDeviceToHost, st));
        RMG_CUDA(cudaStreamSynchronize(st));
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) g_last_kernel_ms = ms;
    }
cleanup:
    g_launch_count += launches;
    if (up) cudaStreamSynchronize(up);
    if (st) cudaStreamSynchronize(st);
    {
        void *all[] = {dY, dM, dD, dU, dScale, dOffset, keys, dcols};
        for (void *x : all)
            if (x) cudaFree(x);
    }
    for (auto e : landed)
        if (e) cudaEventDestroy(e);
    ring.release_all();
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (up) cudaStreamDestroy(up);
    if (st) cudaStreamDestroy(st);
    if (rc != RMG_OK) cudaGetLastError();
    return rc;
}

int rmg_randomize(const double *Y, int64_t V, int64_t ldY, int n, const double *X, int p, const double *mask,
                  double mask_lo, double mask_hi, const int32_t *idx, int R, const int32_t *cols, int ncols,
                  int two_sided, double *dist, int device, char *err, size_t errlen)
{
    return randomize_host_impl(Y, nullptr, V, ldY, n, X, p, mask, mask_lo, mask_hi, idx, R, cols, ncols, two_sided, dist, device,
                               err, errlen);
}

int rmg_randomize_stored(const void *U, int dtype, int64_t V, int64_t ldU, int n, const double *scale, const double *offset,
                         double voxel_min, int64_t slice_len, const double *X, int p, const double *mask, double mask_lo,
                         double mask_hi, const int32_t *idx, int R, const int32_t *cols, int ncols, int two_sided, double *dist,
                         int device, char *err, size_t errlen)
{
    int rc = check_stored_args(dtype, scale, offset, voxel_min, slice_len, p, err, errlen);
    if (rc) return rc;
    const StoredSpec pk{dtype, scale, offset, voxel_min, dtype == RMG_STORED_U16 ? slice_len : std::max<int64_t>(V, 1)};
    return randomize_host_impl(U, &pk, V, ldU, n, X, p, mask, mask_lo, mask_hi, idx, R, cols, ncols, two_sided, dist, device, err,
                               errlen);
}

}  // extern "C"

// Voxel shards over devices 0..G-1, one host thread per GPU; the only exchange step of the path is an element-wise max
// of the per-shard R x ncols maxima.  fn(g, v0, cv, part, err, errlen) runs voxels [v0, v0 + cv) on device g.
template <class Fn>
static int randomize_sharded(int64_t V, int R, int ncols, double *dist, int n_gpus, char *err, size_t errlen, Fn fn)
{
    const int have = rmg_device_count();
    if (have <= 0) return fail(err, errlen, RMG_ERR_NO_DEVICE, "no CUDA device available; rminc_b200 has no CPU fallback");
    int G = n_gpus <= 0 ? have : n_gpus;
    if (G > have) return fail(err, errlen, RMG_ERR_NO_DEVICE, "%d GPUs requested, %d visible", G, have);
    if (R <= 0 || ncols <= 0) return RMG_OK;
    std::vector<int64_t> b(G + 1);
    for (int g = 0; g <= G; ++g) b[g] = std::min<int64_t>(V, ((V * g / G) + 1) & ~int64_t(1));
    b[0] = 0;
    b[G] = V;
    std::vector<int> rcs(G, RMG_OK);
    std::vector<std::string> msgs(G, std::string(256, '\0'));
    std::vector<std::vector<double>> part(G, std::vector<double>((size_t)R * ncols, -INFINITY));
    std::vector<double> kms(G, 0.0);
    std::vector<std::thread> th;
    for (int g = 0; g < G; ++g) {
        th.emplace_back([&, g]() {
            const int64_t v0 = b[g], cv = b[g + 1] - b[g];
            if (cv <= 0) return;
            rcs[g] = fn(g, v0, cv, part[g].data(), &msgs[g][0], msgs[g].size());
            kms[g] = rmg_last_kernel_ms();
        });
    }
    for (auto &t : th) t.join();
    g_last_kernel_ms = *std::max_element(kms.begin(), kms.end());
    for (int g = 0; g < G; ++g)
        if (rcs[g] != RMG_OK) return fail(err, errlen, rcs[g], "gpu %d: %s", g, msgs[g].c_str());
    for (size_t i = 0; i < (size_t)R * ncols; ++i) {
        double m = -INFINITY;
        for (int g = 0; g < G; ++g) m = std::fmax(m, part[g][i]);
        dist[i] = m;
    }
    return RMG_OK;
}

extern "C" {

int rmg_randomize_multi(const double *Y, int64_t V, int64_t ldY, int n, const double *X, int p, const double *mask,
                        double mask_lo, double mask_hi, const int32_t *idx, int R, const int32_t *cols, int ncols,
                        int two_sided, double *dist, int n_gpus, char *err, size_t errlen)
{
    int rc = check_fit_args(Y, V, ldY, n, X, p, Y, ldY, err, errlen);
    if (rc) return rc;
    return randomize_sharded(V, R, ncols, dist, n_gpus, err, errlen,
                             [&](int g, int64_t v0, int64_t cv, double *part, char *e, size_t el) {
        return rmg_randomize(Y + v0, cv, ldY, n, X, p, mask ? mask + v0 : nullptr, mask_lo, mask_hi, idx, R, cols, ncols,
                             two_sided, part, g, e, el);
    });
}

int rmg_randomize_stored_multi(const void *U, int dtype, int64_t V, int64_t ldU, int n, const double *scale,
                               const double *offset, double voxel_min, int64_t slice_len, const double *X, int p,
                               const double *mask, double mask_lo, double mask_hi, const int32_t *idx, int R,
                               const int32_t *cols, int ncols, int two_sided, double *dist, int n_gpus, char *err,
                               size_t errlen)
{
    int rc = check_stored_args(dtype, scale, offset, voxel_min, slice_len, p, err, errlen);
    if (rc) return rc;
    if ((rc = check_fit_args(U, V, ldU, n, X, p, U, ldU, err, errlen))) return rc;
    const size_t es = dtype == RMG_STORED_U16 ? 2 : 4;
    const int64_t sl = dtype == RMG_STORED_U16 ? slice_len : std::max<int64_t>(V, 1);
    const int64_t nslices_total = (V + sl - 1) / sl;
    return randomize_sharded(V, R, ncols, dist, n_gpus, err, errlen,
                             [&](int g, int64_t v0, int64_t cv, double *part, char *e, size_t el) {
        StoredSpec pk{dtype, scale, offset, voxel_min, sl};
        pk.v_offset = v0;                    // slices keep their global index
        pk.nslices_total = nslices_total;
        return randomize_host_impl(static_cast<const char *>(U) + (size_t)v0 * es, &pk, cv, ldU, n, X, p,
                                   mask ? mask + v0 : nullptr, mask_lo, mask_hi, idx, R, cols, ncols, two_sided, part, g, e, el);
    });
}

}  // extern "C"
