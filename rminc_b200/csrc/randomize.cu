// randomize.cu -- the TFCE permutation tests (vertexTFCE.vertexLm, mincTFCE.mincLm): per permutation a streaming refit,
// the statistic maps of a batch of permutations enhanced together (tfce.cu), column maxima.
//
// Reference: mincRandomize_core with post_proc (R/minc_voxel_statistics.R:1465-1497, :1382-1439;
// R/minc_vertex_statistics.R:892-948).  The plain permutation test (no enhancement) lives in dataset.cu: it runs on
// the resident-dataset engine (stacked FP64 tensor-core GEMM, perm_gemm.cu).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <thread>

#include "capi_common.hpp"
#include "dataset.hpp"
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

}  // namespace
}  // namespace rmg

using namespace rmg;

extern "C" {

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
    NvtxPhase nv("rmg randomize + TFCE");
    nv.next("factor R designs");
    const bool timing = std::getenv("RMG_TIMING") != nullptr;
    auto t_mark = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!timing) return;
        const auto t = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[rmg tfce] %-44s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t - t_mark).count());
        t_mark = t;
    };
    PermSet pd;
    if ((rc = factor_permutations(X, n, p, idx, R, pd, err, errlen))) return rc;
    if ((rc = select_device(device, err, errlen))) return rc;
    if (R == 0 || ncols == 0 || V == 0) return RMG_OK;
    lap("argument checks + factor R designs");

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
        nv.next("graph + workspaces + upload of Y");
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
                                 reinterpret_cast<const char *>(Y), (size_t)ldY * 8, (size_t)V * 8, n, st, device));
        }
        RMG_CUDA(cudaMemcpyAsync(dcols, cols, (size_t)ncols * 4, cudaMemcpyHostToDevice, st));
        if (mask) {
            RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&dM), (size_t)V * 8, st));
            RMG_CUDA(cudaMemcpyAsync(dM, mask, (size_t)V * 8, cudaMemcpyHostToDevice, st));
        }
        keys_init_kernel<<<(R * ncols + 255) / 256, 256, 0, st>>>(keys, R * ncols);
        ++launches;
        if (timing) {
            RMG_CUDA(cudaStreamSynchronize(st));
            lap("graph, workspaces, Y on the device");
        }
        nv.next("refits + enhancement per batch of permutations");
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
        lap("refits + TFCE + maxima (all batches)");
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

}  // extern "C"
