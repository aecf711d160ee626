// dataset.cu -- the resident dataset: Y uploaded ONCE, sharded over the GPUs of the box inside one process, and every
// routine of the path (fit, anova, permutation test, FDR) run on the resident shards.
//
// Reference: mincLm keeps `data` and the call as attributes so that mincRandomize / mincAnova / mincFDR can be run on
// the same files again (R/minc_voxel_statistics.R:911-912, :1457-1477) -- and each of them re-reads every file.  On
// the GPU the expensive step is the host link (28.3 GB at ~55 GB/s = 0.5 s against a 4.4 ms fit), so the staged matrix
// stays in HBM behind an opaque handle.  Voxel shards are contiguous (create_parallel_mask +
// mclapply, R/minc_voxel_statistics.R:848-906); the only exchange of the path is the element-wise max of the
// permutation null distribution (R x ncols doubles), done on the host because R is one process.
//
// The same engine backs rmg_randomize / rmg_randomize_stored and their _multi forms (a transient dataset whose
// upload is consumed block by block while it lands), and rmg_randomize_device (a borrowed, already resident shard).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <mutex>
#include <string>
#include <thread>

#include "capi_common.hpp"
#include "dataset.hpp"
#include "perm_gemm.cuh"

namespace rmg {

thread_local int tl_host_share = 1;

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
// (the requested columns travel as a kernel argument: a host -> device copy of them would queue behind a dataset
// upload still in flight on the copy engine)
struct ColList {
    int32_t c[2 * RMG_MAX_P + 3];
};

__global__ void __launch_bounds__(256)
colmax_kernel(const double *__restrict__ out, int64_t V, int64_t ld, const ColList cols, int ncols,
              int two_sided, unsigned long long *__restrict__ keys, int R, int r)
{
    __shared__ double red[8];
    for (int c = 0; c < ncols; ++c) {
        const double *col = out + (int64_t)cols.c[c] * ld;
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

enum class PermPath { Auto, Refit, Gemm };

PermPath perm_path_from_env()
{
    if (const char *v = std::getenv("RMG_PERM_PATH")) {
        if (!std::strcmp(v, "refit")) return PermPath::Refit;
        if (!std::strcmp(v, "gemm")) return PermPath::Gemm;
    }
    return PermPath::Auto;
}

double now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Refit path for one shard: one streaming fit per permutation into a scratch result matrix + column maxima.  Used for
// result columns other than tvalue-*, p > 8 and odd layouts; also the validation twin of the GEMM.
int refit_shard(const PermSet &pd, const PermShard &s, int n, int p, double mask_lo, double mask_hi, const int32_t *cols,
                int ncols, int two_sided, unsigned long long *keys, int R, int64_t &launches, char *err, size_t errlen)
{
    int rc = RMG_OK;
    double *scratch = nullptr, *qt = nullptr;
    DevDesign *dd = nullptr;
    LmScratch scr;
    LmPlan plan;
    plan.sm_count = s.sm_count;
    cudaStream_t st = s.st;
    const int ncol_total = 2 * p + 3, KP = lm_padded_p(p);
    const int64_t ldo = (s.cv + 1) & ~int64_t(1);
    const size_t qstride = make_qt(pd.base, KP).size();   // rows padded for the TMA kernel
    ColList cl;
    std::memset(&cl, 0, sizeof cl);
    for (int c = 0; c < ncols && c < 2 * RMG_MAX_P + 3; ++c) cl.c[c] = cols[c];
    {
        RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&scratch), (size_t)ldo * ncol_total * sizeof(double), st));
        RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&dd), sizeof(DevDesign) * (size_t)R, st));
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
            RMG_CUDA(cudaMemcpyAsync(dd, hdd.data(), sizeof(DevDesign) * (size_t)R, cudaMemcpyHostToDevice, st));
            RMG_CUDA(cudaMemcpyAsync(qt, hqt.data(), sizeof(double) * hqt.size(), cudaMemcpyHostToDevice, st));
            RMG_CUDA(cudaStreamSynchronize(st));   // pageable sources die with this scope
        }
        const int cm_grid = (int)std::min<int64_t>((s.cv + 255) / 256, (int64_t)plan.sm_count * 8);
        for (int r = 0; r < R; ++r) {
            LmArgs a{s.dY, s.cv, s.ld, n, p, qt + (size_t)r * qstride, dd + r, s.dM, mask_lo, mask_hi, scratch, ldo};
            LmLaunchInfo info;
            RMG_CUDA(launch_lm_fit(a, plan, scr, st, &info));
            launches += info.launches;
            colmax_kernel<<<cm_grid, 256, 0, st>>>(scratch, s.cv, ldo, cl, ncols, two_sided, keys, R, r);
            ++launches;
            RMG_CUDA(cudaGetLastError());
        }
    }
cleanup:
    if (scratch) cudaFreeAsync(scratch, st);
    if (dd) cudaFreeAsync(dd, st);
    if (qt) cudaFreeAsync(qt, st);
    scr.release(st);
    if (rc != RMG_OK) cudaGetLastError();
    return rc;
}

}  // namespace

// Factor every permuted design; translate the reference's error contract.
int factor_permutations(const double *X, int n, int p, const int32_t *idx, int R, PermSet &pd, char *err, size_t errlen)
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

int check_perm_args(int p, const int32_t *idx, int R, const int32_t *cols, int ncols, const void *dist, char *err, size_t errlen)
{
    if (R < 0 || ncols < 0 || (R > 0 && !idx) || (ncols > 0 && !cols) || (R > 0 && ncols > 0 && !dist))
        return fail(err, errlen, RMG_ERR_INVALID, "bad permutation arguments");
    if (ncols > 2 * RMG_MAX_P + 3) return fail(err, errlen, RMG_ERR_INVALID, "more columns requested than a result has");
    for (int c = 0; c < ncols; ++c)
        if (cols[c] < 0 || cols[c] >= 2 * p + 3) return fail(err, errlen, RMG_ERR_INVALID, "cols[%d] out of range", c);
    return RMG_OK;
}

// All R permutations on a set of resident voxel shards (one per device).  Everything is queued on the shards' streams
// from this one host thread; the host work (factorisation, operand packing) is done ONCE and shared by every device.
// Per shard the result is R x ncols maxima over ITS voxels, written to s.ddist (device) and / or s.host_part.
int randomize_shards(std::vector<PermShard> &S, int n, int p, const double *X, double mask_lo, double mask_hi,
                     const int32_t *idx, int R, const int32_t *cols, int ncols, int two_sided, bool sync_host, char *err,
                     size_t errlen)
{
    int rc = RMG_OK;
    if (R == 0 || ncols == 0) return rc;
    const bool timing = std::getenv("RMG_TIMING") != nullptr;
    double t_mark = now_ms();
    NvtxPhase nv("rmg randomize");
    nv.next("factor R designs");
    auto lap = [&](const char *what, const char *next_phase) {
        nv.next(next_phase);
        if (!timing) return;
        const double t = now_ms();
        std::fprintf(stderr, "[rmg perm] %-34s %8.3f ms\n", what, t - t_mark);
        t_mark = t;
    };
    PermSet pd;
    if ((rc = factor_permutations(X, n, p, idx, R, pd, err, errlen))) return rc;
    lap("factor R designs", "plan + per-device workspaces");

    const int G = (int)S.size();
    const int nk = R * ncols;
    const PermPath path = perm_path_from_env();
    bool gemm = path != PermPath::Refit;
    for (const PermShard &s : S)
        if (s.cv > 0 && !perm_gemm_supported(n, p, s.cv, s.ld, s.dY, cols, ncols)) gemm = false;
    for (const PermShard &s : S)
        for (size_t b = 0; b < s.blk_v0.size(); ++b)
            if (s.blk_v0[b] % 2) gemm = false;          // bulk copies need 16-byte aligned block starts
    if (!gemm && path == PermPath::Gemm)
        return fail(err, errlen, RMG_ERR_INVALID, "RMG_PERM_PATH=gemm but the GEMM path does not support this shape");

    std::vector<unsigned long long *> keys(G, nullptr);
    std::vector<double *> dD(G, nullptr);
    std::vector<PermDevice> dev(G);
    std::vector<char> begun(G, 0);
    std::vector<cudaEvent_t> e0(G, nullptr), e1(G, nullptr);
    std::vector<int64_t> launches(G, 0);
    PermPack pk;
    bool planned = false;
    auto live = [&](int g) { return S[g].cv > 0; };
    {
        for (int g = 0; g < G; ++g) {
            if (!live(g)) continue;
            PermShard &s = S[g];
            RMG_CUDA(cudaSetDevice(s.device));
            RMG_CUDA(cudaEventCreate(&e0[g]));
            RMG_CUDA(cudaEventCreate(&e1[g]));
            RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&keys[g]), (size_t)nk * sizeof(unsigned long long), s.st));
            if (!s.ddist) RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&dD[g]), (size_t)nk * sizeof(double), s.st));
            keys_init_kernel<<<(nk + 255) / 256, 256, 0, s.st>>>(keys[g], nk);
            ++launches[g];
            RMG_CUDA(cudaGetLastError());
            RMG_CUDA(cudaEventRecord(e0[g], s.st));
        }
        if (gemm) {
            RMG_CUDA(perm_pack_plan(pd, n, p, cols, ncols, pk));
            planned = true;
            const int shift = pk.shift;
            std::vector<PermGemmArgs> ga(G);
            for (int g = 0; g < G; ++g) {
                if (!live(g)) continue;
                PermShard &s = S[g];
                RMG_CUDA(cudaSetDevice(s.device));
                ga[g] = PermGemmArgs{s.dY, s.cv, s.ld, n, p, s.dM, mask_lo, mask_hi, cols, ncols, two_sided, keys[g], R};
                const PermPre *resident = nullptr;
                if (s.pre) {
                    PermPre &pre = s.pre[shift];
                    if (!pre.y0) {           // first permutation test on this dataset: allocate the resident sums
                        RMG_CUDA(cudaMalloc(reinterpret_cast<void **>(&pre.y0), (size_t)s.cv * 8));
                        RMG_CUDA(cudaMalloc(reinterpret_cast<void **>(&pre.yy), (size_t)s.cv * 8));
                        RMG_CUDA(cudaMalloc(reinterpret_cast<void **>(&pre.ys), (size_t)s.cv * 8));
                        RMG_CUDA(cudaMalloc(reinterpret_cast<void **>(&pre.any_zero), 256));
                        RMG_CUDA(cudaMemsetAsync(pre.any_zero, 0, 256, s.st));
                    }
                    resident = &pre;
                }
                int l = 0;
                RMG_CUDA(perm_dev_begin(dev[g], pk, s.device, s.sm_count, ga[g], s.st, resident, &l));
                begun[g] = 1;
                launches[g] += l;
            }
            lap("plan + per-device workspaces", "prepass; pack + upload + launch per batch");
            // per-voxel sums of block `blk` of shard g, once it has landed (skipped when the dataset already holds them)
            auto prepass = [&](int g, size_t blk) -> cudaError_t {
                PermShard &s = S[g];
                if (s.landed.size() > blk && s.landed[blk]) {
                    cudaError_t e = cudaStreamWaitEvent(s.st, s.landed[blk], 0);
                    if (e != cudaSuccess) return e;
                }
                if (s.pre_done && s.pre_done[shift].size() > blk && s.pre_done[shift][blk]) return cudaSuccess;
                int l = 0;
                cudaError_t e = perm_dev_prepass(dev[g], ga[g], s.blk_v0[blk], s.blk_cv[blk], &l);
                launches[g] += l;
                if (e == cudaSuccess && s.pre_done && s.pre_done[shift].size() > blk) s.pre_done[shift][blk] = 1;
                return e;
            };
            for (int g = 0; g < G; ++g) {
                if (!live(g)) continue;
                RMG_CUDA(cudaSetDevice(S[g].device));
                RMG_CUDA(prepass(g, 0));        // needs only Y: runs while the host packs
            }
            // block 0: pack a batch of permutations on the host (all cores), hand it to every device, launch it; the
            // host packs batch b+1 while the GPUs multiply batch b
            for (int b = 0; b < pk.nbatch; ++b) {
                perm_pack_batch(pd, pk, b);
                for (int g = 0; g < G; ++g) {
                    if (!live(g)) continue;
                    RMG_CUDA(cudaSetDevice(S[g].device));
                    RMG_CUDA(perm_dev_upload(dev[g], pk, b));
                    int l = 0;
                    RMG_CUDA(perm_dev_launch(dev[g], pk, b, b + 1, S[g].blk_v0[0], S[g].blk_cv[0], &l));
                    launches[g] += l;
                }
            }
            lap("pack + upload + block-0 launches", "launches of the later voxel blocks; finalize");
            // later blocks: the whole operand is on the device, one launch per block over every permutation tile
            size_t max_blocks = 0;
            for (int g = 0; g < G; ++g) max_blocks = std::max(max_blocks, S[g].blk_v0.size());
            for (size_t blk = 1; blk < max_blocks; ++blk)
                for (int g = 0; g < G; ++g) {
                    if (!live(g) || blk >= S[g].blk_v0.size()) continue;
                    RMG_CUDA(cudaSetDevice(S[g].device));
                    RMG_CUDA(prepass(g, blk));
                    int l = 0;
                    RMG_CUDA(perm_dev_launch(dev[g], pk, 0, pk.nbatch, S[g].blk_v0[blk], S[g].blk_cv[blk], &l));
                    launches[g] += l;
                }
            for (int g = 0; g < G; ++g) {
                if (!live(g)) continue;
                RMG_CUDA(cudaSetDevice(S[g].device));
                int l = 0;
                RMG_CUDA(perm_dev_end(dev[g], pk, &l));
                launches[g] += l;
            }
        } else {
            // refit path: the launch queues are deep (2 R launches per device), so one host thread per shard
            std::vector<int> rcs(G, RMG_OK);
            std::vector<std::string> msgs(G, std::string(256, '\0'));
            std::vector<std::thread> th;
            for (int g = 0; g < G; ++g) {
                if (!live(g)) continue;
                th.emplace_back([&, g]() {
                    PermShard &s = S[g];
                    cudaSetDevice(s.device);
                    for (cudaEvent_t ev : s.landed)
                        if (ev) cudaStreamWaitEvent(s.st, ev, 0);
                    tl_host_share = G;
                    rcs[g] = refit_shard(pd, s, n, p, mask_lo, mask_hi, cols, ncols, two_sided, keys[g], R, launches[g],
                                         &msgs[g][0], msgs[g].size());
                });
            }
            for (auto &t : th) t.join();
            for (int g = 0; g < G; ++g)
                if (rcs[g] != RMG_OK) {
                    rc = fail(err, errlen, rcs[g], "%s", msgs[g].c_str());
                    goto cleanup;
                }
        }
        for (int g = 0; g < G; ++g) {
            if (!live(g)) continue;
            PermShard &s = S[g];
            RMG_CUDA(cudaSetDevice(s.device));
            double *dst = s.ddist ? s.ddist : dD[g];
            keys_decode_kernel<<<(nk + 255) / 256, 256, 0, s.st>>>(keys[g], dst, nk);
            ++launches[g];
            RMG_CUDA(cudaGetLastError());
            RMG_CUDA(cudaEventRecord(e1[g], s.st));
            if (s.host_part) RMG_CUDA(cudaMemcpyAsync(s.host_part, dst, (size_t)nk * sizeof(double), cudaMemcpyDeviceToHost, s.st));
        }
        lap("remaining launches queued", "wait for the devices");
        if (sync_host) {
            double ms_max = 0.0;
            for (int g = 0; g < G; ++g) {
                if (!live(g)) continue;
                RMG_CUDA(cudaSetDevice(S[g].device));
                RMG_CUDA(cudaStreamSynchronize(S[g].st));
                float ms = 0.f;
                if (cudaEventElapsedTime(&ms, e0[g], e1[g]) == cudaSuccess) ms_max = std::max(ms_max, (double)ms);
            }
            g_last_kernel_ms = ms_max;
            lap("GPU work (sync)", "release");
        }
    }
cleanup:
    for (int g = 0; g < G; ++g) {
        if (!live(g)) continue;
        cudaSetDevice(S[g].device);
        g_launch_count += launches[g];
        if (begun[g]) perm_dev_release(dev[g]);
        if (keys[g]) cudaFreeAsync(keys[g], S[g].st);
        if (dD[g]) cudaFreeAsync(dD[g], S[g].st);
        if (e0[g]) cudaEventDestroy(e0[g]);
        if (e1[g]) cudaEventDestroy(e1[g]);
    }
    if (planned) perm_pack_release(pk);
    if (rc != RMG_OK) cudaGetLastError();
    return rc;
}

// ---------------------------------------------------------------------------------------------------- the dataset
struct DsShard {
    int device = 0, sm_count = 148;
    int64_t v0 = 0, cv = 0, ld = 0;          // first voxel (global), voxels, leading dimension of dY
    double *dY = nullptr, *dM = nullptr;
    char *dU = nullptr;                       // one block of stored samples (stored-type uploads)
    double *dScale = nullptr, *dOffset = nullptr;
    cudaStream_t st = nullptr, up = nullptr;
    std::vector<int64_t> blk_v0, blk_cv;
    std::vector<cudaEvent_t> landed;
    double *dOut = nullptr;                   // resident result of the last fit on this handle
    int64_t ldo = 0;
    size_t out_cap = 0;
    PermPre pre[2];
    std::vector<char> pre_done[2];
    LmScratch scr;
};

}  // namespace rmg

using namespace rmg;

struct rmg_dataset {
    int64_t V = 0;
    int n = 0, G = 0;
    bool has_mask = false;
    double mask_lo = 0.0, mask_hi = 0.0;
    std::vector<DsShard> sh;
    int out_cols = 0, out_p = 0, out_kind = 0;     // last fit: columns, design columns, 1 = lm, 2 = anova
    double *dMaskFull = nullptr;                    // whole mask on the first device (FDR over all shards), G > 1
    std::mutex mu;
    double upload_ms = 0.0;
};

namespace rmg {

void dataset_destroy(rmg_dataset *ds)
{
    if (!ds) return;
    for (DsShard &s : ds->sh) {
        if (cudaSetDevice(s.device) != cudaSuccess) continue;
        if (s.up) cudaStreamSynchronize(s.up);
        if (s.st) cudaStreamSynchronize(s.st);
        s.scr.release(s.st);
        void *all[] = {s.dY, s.dM, s.dU, s.dScale, s.dOffset, s.dOut, s.pre[0].y0, s.pre[0].yy, s.pre[0].ys, s.pre[0].any_zero,
                       s.pre[1].y0, s.pre[1].yy, s.pre[1].ys, s.pre[1].any_zero};
        for (void *x : all)
            if (x) cudaFree(x);
        for (cudaEvent_t e : s.landed)
            if (e) cudaEventDestroy(e);
        if (s.up) cudaStreamDestroy(s.up);
        if (s.st) cudaStreamDestroy(s.st);
    }
    if (ds->dMaskFull && !ds->sh.empty() && cudaSetDevice(ds->sh[0].device) == cudaSuccess) cudaFree(ds->dMaskFull);
    cudaGetLastError();
    delete ds;
}

// Creates the handle and QUEUES the upload: voxel blocks travel on each shard's copy stream (straight into Y for
// doubles; through one block-sized buffer + the expansion kernel for stored samples) and every block records an
// event, so a consumer can start on block b while block b+1 is still on the link.  Pageable input is staged through
// the pinned ring by all host threads (the call then returns when the last block has been handed to the DMA engine).
int dataset_create(const void *Yv, const StoredSpec *pk, int64_t V, int64_t ldY, int n, const double *mask, double mask_lo,
                   double mask_hi, const int *devices, int G, rmg_dataset **out, char *err, size_t errlen)
{
    int rc = RMG_OK;
    NvtxPhase nv("rmg dataset upload (queued per voxel block)");
    *out = nullptr;
    rmg_dataset *ds = new rmg_dataset();
    ds->V = V;
    ds->n = n;
    ds->G = G;
    ds->has_mask = mask != nullptr;
    ds->mask_lo = mask_lo;
    ds->mask_hi = mask_hi;
    ds->sh.resize(G);
    const size_t es = !pk ? sizeof(double) : pk->dtype == kStoredU16 ? 2 : 4;   // bytes per sample of the input
    const char *Yb = static_cast<const char *>(Yv);
    const int64_t slice_len = pk ? std::max<int64_t>(1, pk->slice_len) : 1;
    const int64_t nslices = !pk ? 0 : pk->nslices_total > 0 ? pk->nslices_total : (V + slice_len - 1) / slice_len;
    PinnedRing ring;
    bool ring_live = false;
    const double t0 = now_ms();
    {
        // even shard boundaries keep every shard 16-byte aligned
        std::vector<int64_t> b(G + 1);
        for (int g = 0; g <= G; ++g) b[g] = std::min<int64_t>(V, ((V * g / G) + 1) & ~int64_t(1));
        b[0] = 0;
        b[G] = V;
        int64_t max_cv = 0;
        size_t max_blocks = 0;
        for (int g = 0; g < G; ++g) {
            DsShard &s = ds->sh[g];
            s.device = devices[g];
            s.v0 = b[g];
            s.cv = b[g + 1] - b[g];
            s.ld = std::max<int64_t>(8, (s.cv + 7) & ~int64_t(7));
            s.sm_count = device_sm_count(s.device);
            // voxel blocks: about an eighth of the shard, at least ~512 MiB of doubles, multiples of 1024 voxels
            int64_t CV = std::max<int64_t>((s.cv + 7) / 8, (int64_t)(512.0 * 1024 * 1024 / (8.0 * n)));
            CV = std::max<int64_t>(1024, ((CV + 1023) / 1024) * 1024);
            if (const char *e = std::getenv("RMG_PERM_CHUNK_VOXELS")) CV = std::max<int64_t>(2, (std::atoll(e) / 2) * 2);
            if (CV > s.cv) CV = std::max<int64_t>(s.cv, 1);
            for (int64_t v = 0; v < s.cv; v += CV) {
                s.blk_v0.push_back(v);
                s.blk_cv.push_back(std::min(CV, s.cv - v));
            }
            if (s.blk_v0.empty()) {            // an empty shard (V < G): one empty block keeps the loops uniform
                s.blk_v0.push_back(0);
                s.blk_cv.push_back(0);
            }
            max_cv = std::max(max_cv, CV);
            max_blocks = std::max(max_blocks, s.blk_v0.size());
            s.pre_done[0].assign(s.blk_v0.size(), 0);
            s.pre_done[1].assign(s.blk_v0.size(), 0);
            s.landed.assign(s.blk_v0.size(), nullptr);
            RMG_CUDA(cudaSetDevice(s.device));
            RMG_CUDA(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
            RMG_CUDA(cudaStreamCreateWithFlags(&s.up, cudaStreamNonBlocking));
            if (s.cv == 0) continue;
            RMG_CUDA(cudaMalloc(&s.dY, (size_t)s.ld * n * sizeof(double)));
            if (pk) {
                RMG_CUDA(cudaMalloc(&s.dU, (size_t)CV * n * es));          // one block of stored samples, rows of CV
                if (pk->dtype == kStoredU16) {
                    const size_t cnt = (size_t)nslices * n * sizeof(double);
                    RMG_CUDA(cudaMalloc(&s.dScale, cnt));
                    RMG_CUDA(cudaMalloc(&s.dOffset, cnt));
                    RMG_CUDA(cudaMemcpyAsync(s.dScale, pk->scale, cnt, cudaMemcpyHostToDevice, s.up));
                    RMG_CUDA(cudaMemcpyAsync(s.dOffset, pk->offset, cnt, cudaMemcpyHostToDevice, s.up));
                }
            }
            if (mask) {
                RMG_CUDA(cudaMalloc(&s.dM, (size_t)s.cv * sizeof(double)));
                RMG_CUDA(cudaMemcpyAsync(s.dM, mask + s.v0, (size_t)s.cv * 8, cudaMemcpyHostToDevice, s.up));
            }
        }
        if (mask && G > 1 && V > 0) {
            RMG_CUDA(cudaSetDevice(ds->sh[0].device));
            RMG_CUDA(cudaMalloc(&ds->dMaskFull, (size_t)V * sizeof(double)));
            RMG_CUDA(cudaMemcpyAsync(ds->dMaskFull, mask, (size_t)V * 8, cudaMemcpyHostToDevice, ds->sh[0].up));
        }
        int rows_per_slot = 1;
        const bool staged = V > 0 && staging_plan(Yv, (double)V * n * (double)es, (size_t)max_cv * es, n, rows_per_slot);
        if (staged) {
            RMG_CUDA(ring.init((size_t)rows_per_slot * (size_t)max_cv * es));
            ring_live = true;
        }
        // blocks outer, devices inner: every link starts at once and every device's first block lands first
        for (size_t blk = 0; blk < max_blocks; ++blk)
            for (int g = 0; g < G; ++g) {
                DsShard &s = ds->sh[g];
                if (blk >= s.blk_v0.size()) continue;
                RMG_CUDA(cudaSetDevice(s.device));
                const int64_t bv0 = s.blk_v0[blk], bcv = s.blk_cv[blk];
                if (bcv > 0) {
                    const int64_t CV = s.blk_cv[0];      // capacity of the stored-sample buffer (rows of CV samples)
                    char *dstdev = pk ? s.dU : reinterpret_cast<char *>(s.dY + bv0);
                    const size_t dpitch = pk ? (size_t)CV * es : (size_t)s.ld * 8;
                    RMG_CUDA(upload_rows(staged ? &ring : nullptr, rows_per_slot, dstdev, dpitch, Yb + (size_t)(s.v0 + bv0) * es,
                                         (size_t)ldY * es, (size_t)bcv * es, n, s.up, s.device));
                    // stored samples become the reader's doubles on the copy stream too: stream order then protects the
                    // one staging block, and the expansion (~1 ms) is noise next to the block's transfer
                    if (pk) {
                        RMG_CUDA(launch_expand_stored(s.dU, pk->dtype, bcv, CV, n, s.dScale, s.dOffset,
                                                      4503599627370496.0 + pk->voxel_min, slice_len, nslices,
                                                      pk->v_offset + s.v0 + bv0, s.dY + bv0, s.ld, s.up));
                        g_launch_count += 1;
                    }
                }
                RMG_CUDA(cudaEventCreateWithFlags(&s.landed[blk], cudaEventDisableTiming));
                RMG_CUDA(cudaEventRecord(s.landed[blk], s.up));
            }
    }
cleanup:
    if (ring_live) ring.release_all();       // waits for the staged copies still in flight
    ds->upload_ms = now_ms() - t0;
    if (rc != RMG_OK) {
        dataset_destroy(ds);
        cudaGetLastError();
        return rc;
    }
    *out = ds;
    return RMG_OK;
}

std::vector<PermShard> dataset_perm_shards(rmg_dataset *ds, std::vector<std::vector<double>> &parts, int nk)
{
    std::vector<PermShard> S(ds->G);
    parts.assign(ds->G, std::vector<double>((size_t)std::max(nk, 1), -INFINITY));
    for (int g = 0; g < ds->G; ++g) {
        DsShard &s = ds->sh[g];
        PermShard &o = S[g];
        o.device = s.device;
        o.sm_count = s.sm_count;
        o.dY = s.dY;
        o.cv = s.cv;
        o.ld = s.ld;
        o.dM = s.dM;
        o.st = s.st;
        o.blk_v0 = s.blk_v0;
        o.blk_cv = s.blk_cv;
        o.landed = s.landed;
        o.pre = s.pre;
        o.pre_done = s.pre_done;
        o.host_part = parts[g].data();
    }
    return S;
}

int dataset_randomize(rmg_dataset *ds, const double *X, int p, const int32_t *idx, int R, const int32_t *cols, int ncols,
                      int two_sided, double *dist, char *err, size_t errlen)
{
    int rc = check_perm_args(p, idx, R, cols, ncols, dist, err, errlen);
    if (rc) return rc;
    if (!X || p < 1 || p > RMG_MAX_P) return fail(err, errlen, RMG_ERR_INVALID, "bad design arguments");
    if (R == 0 || ncols == 0) return RMG_OK;
    std::lock_guard<std::mutex> lk(ds->mu);
    const int nk = R * ncols;
    std::vector<std::vector<double>> parts;
    std::vector<PermShard> S = dataset_perm_shards(ds, parts, nk);
    if ((rc = randomize_shards(S, ds->n, p, X, ds->mask_lo, ds->mask_hi, idx, R, cols, ncols, two_sided, true, err, errlen))) return rc;
    // the one exchange of the path: element-wise max of the per-shard null distributions
    for (int i = 0; i < nk; ++i) {
        double m = -INFINITY;
        for (int g = 0; g < ds->G; ++g)
            if (ds->sh[g].cv > 0) m = std::fmax(m, parts[g][i]);
        dist[i] = m;
    }
    return RMG_OK;
}

// mincLm / mincAnova on the resident shards.  out (host, nullable): V x ncol result; the result also stays resident
// for dataset_fdr.
int dataset_fit(rmg_dataset *ds, const double *X, int p, const int32_t *asgn, double *out, int64_t ldOut, char *err, size_t errlen)
{
    if (!X || p < 1 || p > RMG_MAX_P) return fail(err, errlen, RMG_ERR_INVALID, "p must be in 1..%d (got %d)", RMG_MAX_P, p);
    if (out && ldOut < ds->V) return fail(err, errlen, RMG_ERR_INVALID, "leading dimension smaller than V");
    std::lock_guard<std::mutex> lk(ds->mu);
    int rc = RMG_OK;
    NvtxPhase nv("rmg dataset fit (resident shards)");
    int ncol = 2 * p + 3;
    double ms_max = 0.0;
    std::vector<cudaEvent_t> e0(ds->G, nullptr), e1(ds->G, nullptr);
    {
        for (int g = 0; g < ds->G; ++g) {
            DsShard &s = ds->sh[g];
            if (s.cv == 0) continue;
            RMG_CUDA(cudaSetDevice(s.device));
            std::shared_ptr<CachedDesign> cd = design_cache_get(s.device, X, ds->n, p, asgn, rc, err, errlen);
            if (!cd) goto cleanup;
            if (asgn) ncol = (int)*std::max_element(cd->asgn.begin(), cd->asgn.end());
            if (ncol == 0) continue;
            const size_t need = (size_t)s.ld * ncol * sizeof(double);
            if (need > s.out_cap) {
                RMG_CUDA(cudaStreamSynchronize(s.st));
                if (s.dOut) RMG_CUDA(cudaFree(s.dOut));
                s.dOut = nullptr;
                s.out_cap = 0;
                RMG_CUDA(cudaMalloc(&s.dOut, need));
                s.out_cap = need;
            }
            s.ldo = s.ld;
            for (cudaEvent_t ev : s.landed)
                if (ev) RMG_CUDA(cudaStreamWaitEvent(s.st, ev, 0));
            RMG_CUDA(cudaStreamWaitEvent(s.st, cd->ready, 0));
            RMG_CUDA(cudaEventCreate(&e0[g]));
            RMG_CUDA(cudaEventCreate(&e1[g]));
            LmArgs a{s.dY, s.cv, s.ld, ds->n, p, cd->qt, cd->dd, s.dM, ds->mask_lo, ds->mask_hi, s.dOut, s.ldo};
            a.ncol_out = ncol;
            LmLaunchInfo info;
            RMG_CUDA(cudaEventRecord(e0[g], s.st));
            RMG_CUDA(launch_lm_fit(a, plan_from_env(s.device), s.scr, s.st, &info));
            RMG_CUDA(cudaEventRecord(e1[g], s.st));
            g_launch_count += info.launches;
            g_last_fit_variant = info.used_tma ? 1 : 2;
            if (out)
                RMG_CUDA(cudaMemcpy2DAsync(out + s.v0, (size_t)ldOut * 8, s.dOut, (size_t)s.ldo * 8, (size_t)s.cv * 8, (size_t)ncol,
                                           cudaMemcpyDeviceToHost, s.st));
        }
        for (int g = 0; g < ds->G; ++g) {
            DsShard &s = ds->sh[g];
            if (s.cv == 0 || !e0[g]) continue;
            RMG_CUDA(cudaSetDevice(s.device));
            RMG_CUDA(cudaStreamSynchronize(s.st));
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, e0[g], e1[g]) == cudaSuccess) ms_max = std::max(ms_max, (double)ms);
        }
        g_last_kernel_ms = ms_max;
        ds->out_cols = ncol;
        ds->out_p = p;
        ds->out_kind = asgn ? 2 : 1;
    }
cleanup:
    for (int g = 0; g < ds->G; ++g) {
        if (e0[g]) cudaEventDestroy(e0[g]);
        if (e1[g]) cudaEventDestroy(e1[g]);
    }
    if (rc != RMG_OK) cudaGetLastError();
    return rc;
}

// mincFDR on one column of the resident result of the last fit: the column is gathered on the first device (peer
// copies; V doubles) and the one-GPU FDR runs there.
// method < 0: instead of q-values, mean(p >= lambda[i]) of the column's p-values into `fractions` (the pi0 grid of fast.qvalue)
int dataset_fdr(rmg_dataset *ds, int column, int stat_type, double df1, double df2, double *qvalues, double *thresholds,
                char *err, size_t errlen, int method = RMG_FDR_BH, double pi0 = 1.0, const double *lambda = nullptr, int nlambda = 0,
                double *fractions = nullptr)
{
    std::lock_guard<std::mutex> lk(ds->mu);
    if (ds->out_cols == 0) return fail(err, errlen, RMG_ERR_INVALID, "no fitted result on this dataset yet: fit first");
    if (column < 0 || column >= ds->out_cols) return fail(err, errlen, RMG_ERR_INVALID, "column %d out of range (the result has %d)", column, ds->out_cols);
    if (method >= 0 && !qvalues) return fail(err, errlen, RMG_ERR_INVALID, "qvalues is NULL");
    if (method < 0 && (!lambda || !fractions || nlambda < 1)) return fail(err, errlen, RMG_ERR_INVALID, "lambda / fractions missing");
    if (ds->V == 0) return RMG_OK;
    int rc = RMG_OK;
    DsShard &s0 = ds->sh[0];
    double *dcol = nullptr, *dq = nullptr;
    {
        RMG_CUDA(cudaSetDevice(s0.device));
        RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&dq), (size_t)ds->V * 8, s0.st));
        const double *stat = nullptr, *dmask = nullptr;
        if (ds->G == 1) {
            stat = s0.dOut + (size_t)column * s0.ldo;
            dmask = s0.dM;
        } else {
            RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&dcol), (size_t)ds->V * 8, s0.st));
            RMG_CUDA(cudaStreamSynchronize(s0.st));         // the buffer exists before peers write into it
            for (int g = 0; g < ds->G; ++g) {
                DsShard &s = ds->sh[g];
                if (s.cv == 0) continue;
                RMG_CUDA(cudaSetDevice(s.device));
                RMG_CUDA(cudaMemcpyPeerAsync(dcol + s.v0, s0.device, s.dOut + (size_t)column * s.ldo, s.device, (size_t)s.cv * 8, s.st));
            }
            for (int g = 0; g < ds->G; ++g) {
                if (ds->sh[g].cv == 0) continue;
                RMG_CUDA(cudaSetDevice(ds->sh[g].device));
                RMG_CUDA(cudaStreamSynchronize(ds->sh[g].st));
            }
            RMG_CUDA(cudaSetDevice(s0.device));
            RMG_CUDA(cudaStreamSynchronize(s0.up));         // the whole mask landed at creation time
            stat = dcol;
            dmask = ds->dMaskFull;
        }
        if (method < 0) {
            if ((rc = rmg_fdr_pi0_fractions_device(stat, ds->V, stat_type, df1, df2, dmask, lambda, nlambda, fractions, s0.device, s0.st,
                                                   err, errlen)))
                goto cleanup;
            RMG_CUDA(cudaStreamSynchronize(s0.st));
        } else {
            if ((rc = rmg_fdr_method_device(stat, ds->V, stat_type, df1, df2, dmask, dq, thresholds, method, pi0, s0.device, s0.st, err,
                                            errlen)))
                goto cleanup;
            RMG_CUDA(cudaMemcpyAsync(qvalues, dq, (size_t)ds->V * 8, cudaMemcpyDeviceToHost, s0.st));
            RMG_CUDA(cudaStreamSynchronize(s0.st));
        }
    }
cleanup:
    cudaSetDevice(s0.device);
    if (dcol) cudaFreeAsync(dcol, s0.st);
    if (dq) cudaFreeAsync(dq, s0.st);
    if (rc != RMG_OK) cudaGetLastError();
    return rc;
}

int resolve_gpus(int n_gpus, std::vector<int> &devices, char *err, size_t errlen)
{
    const int have = rmg_device_count();
    if (have <= 0) return fail(err, errlen, RMG_ERR_NO_DEVICE, "no CUDA device available; rminc_b200 has no CPU fallback");
    const int G = n_gpus <= 0 ? have : n_gpus;
    if (G > have) return fail(err, errlen, RMG_ERR_NO_DEVICE, "%d GPUs requested, %d visible", G, have);
    if (G > PinnedRing::kMaxDev) return fail(err, errlen, RMG_ERR_INVALID, "at most %d GPUs per process", PinnedRing::kMaxDev);
    devices.resize(G);
    for (int g = 0; g < G; ++g) devices[g] = g;
    return RMG_OK;
}

}  // namespace rmg

// ---------------------------------------------------------------------------------------------------- C ABI
extern "C" {

int rmg_dataset_create(const double *Y, int64_t V, int64_t ldY, int n, const double *mask, double mask_lo, double mask_hi,
                       int n_gpus, rmg_dataset **ds, char *err, size_t errlen)
{
    if (!ds) return fail(err, errlen, RMG_ERR_INVALID, "ds is NULL");
    *ds = nullptr;
    if (V < 0 || n < 1 || ldY < V || (V > 0 && !Y)) return fail(err, errlen, RMG_ERR_INVALID, "bad dataset arguments");
    std::vector<int> devices;
    int rc = resolve_gpus(n_gpus, devices, err, errlen);
    if (rc) return rc;
    return dataset_create(Y, nullptr, V, ldY, n, mask, mask_lo, mask_hi, devices.data(), (int)devices.size(), ds, err, errlen);
}

int rmg_dataset_create_stored(const void *U, int dtype, int64_t V, int64_t ldU, int n, const double *scale, const double *offset,
                              double voxel_min, int64_t slice_len, const double *mask, double mask_lo, double mask_hi, int n_gpus,
                              rmg_dataset **ds, char *err, size_t errlen)
{
    if (!ds) return fail(err, errlen, RMG_ERR_INVALID, "ds is NULL");
    *ds = nullptr;
    if (V < 0 || n < 1 || ldU < V || (V > 0 && !U)) return fail(err, errlen, RMG_ERR_INVALID, "bad dataset arguments");
    int rc = check_stored_args(dtype, scale, offset, voxel_min, slice_len, 1, err, errlen);
    if (rc) return rc;
    std::vector<int> devices;
    if ((rc = resolve_gpus(n_gpus, devices, err, errlen))) return rc;
    const StoredSpec pk{dtype, scale, offset, voxel_min, dtype == RMG_STORED_U16 ? slice_len : std::max<int64_t>(V, 1)};
    return dataset_create(U, &pk, V, ldU, n, mask, mask_lo, mask_hi, devices.data(), (int)devices.size(), ds, err, errlen);
}

void rmg_dataset_free(rmg_dataset *ds) { dataset_destroy(ds); }

int64_t rmg_dataset_voxels(const rmg_dataset *ds) { return ds ? ds->V : 0; }
int rmg_dataset_subjects(const rmg_dataset *ds) { return ds ? ds->n : 0; }
int rmg_dataset_gpus(const rmg_dataset *ds) { return ds ? ds->G : 0; }

int rmg_dataset_sync(rmg_dataset *ds, char *err, size_t errlen)
{
    if (!ds) return fail(err, errlen, RMG_ERR_INVALID, "ds is NULL");
    int rc = RMG_OK;
    for (DsShard &s : ds->sh) {
        RMG_CUDA(cudaSetDevice(s.device));
        RMG_CUDA(cudaStreamSynchronize(s.up));
        RMG_CUDA(cudaStreamSynchronize(s.st));
    }
cleanup:
    return rc;
}

int rmg_dataset_lm_fit(rmg_dataset *ds, const double *X, int p, double *out, int64_t ldOut, char *err, size_t errlen)
{
    if (!ds) return fail(err, errlen, RMG_ERR_INVALID, "ds is NULL");
    return dataset_fit(ds, X, p, nullptr, out, ldOut, err, errlen);
}

int rmg_dataset_anova(rmg_dataset *ds, const double *X, int p, const int32_t *asgn, double *out, int64_t ldOut, char *err,
                      size_t errlen)
{
    if (!ds) return fail(err, errlen, RMG_ERR_INVALID, "ds is NULL");
    if (!asgn) return fail(err, errlen, RMG_ERR_INVALID, "asgn is NULL");
    return dataset_fit(ds, X, p, asgn, out, ldOut, err, errlen);
}

int rmg_dataset_randomize(rmg_dataset *ds, const double *X, int p, const int32_t *idx, int R, const int32_t *cols, int ncols,
                          int two_sided, double *dist, char *err, size_t errlen)
{
    if (!ds) return fail(err, errlen, RMG_ERR_INVALID, "ds is NULL");
    if (ds->n <= p) {
        // the factorisation below reports rank problems itself; n <= p is the reference's own failure mode (rdf <= 0)
    }
    return dataset_randomize(ds, X, p, idx, R, cols, ncols, two_sided, dist, err, errlen);
}

int rmg_dataset_fdr(rmg_dataset *ds, int column, int stat_type, double df1, double df2, double *qvalues, double *thresholds,
                    char *err, size_t errlen)
{
    if (!ds) return fail(err, errlen, RMG_ERR_INVALID, "ds is NULL");
    return dataset_fdr(ds, column, stat_type, df1, df2, qvalues, thresholds, err, errlen);
}

int rmg_dataset_fdr_method(rmg_dataset *ds, int column, int stat_type, double df1, double df2, double *qvalues, double *thresholds,
                           int method, double pi0, char *err, size_t errlen)
{
    if (!ds) return fail(err, errlen, RMG_ERR_INVALID, "ds is NULL");
    if (method < RMG_FDR_BH || method > RMG_FDR_QVALUE) return fail(err, errlen, RMG_ERR_INVALID, "unknown FDR method %d", method);
    return dataset_fdr(ds, column, stat_type, df1, df2, qvalues, thresholds, err, errlen, method, pi0);
}

int rmg_dataset_fdr_pi0_fractions(rmg_dataset *ds, int column, int stat_type, double df1, double df2, const double *lambda,
                                  int nlambda, double *fractions, char *err, size_t errlen)
{
    if (!ds) return fail(err, errlen, RMG_ERR_INVALID, "ds is NULL");
    return dataset_fdr(ds, column, stat_type, df1, df2, nullptr, nullptr, err, errlen, -1, 1.0, lambda, nlambda, fractions);
}

void rmg_release_host_buffers(void) { pinned_trim(); }

// Page-locked host memory with its pages interleaved over the NUMA nodes of the process (see pinned_acquire): the
// staging area a host layer should assemble Y in, so that every GPU's DMA engine reads it at the link rate.
void *rmg_host_alloc(size_t bytes)
{
    if (rmg_device_count() <= 0) return nullptr;
    return pinned_acquire(bytes);
}

void rmg_host_free(void *ptr)
{
    if (!ptr) return;
    pinned_release(ptr);
    pinned_trim();
}

// Concurrent host -> device copy rate of this box: bytes_per_gpu bytes of `host` (pinned; device g reads the g-th
// slice) to each of n_gpus devices at once, `repeats` times.  gbs[g] = rate seen by device g (CUDA events on its
// stream); *aggregate_gbs = all bytes / wall time of the slowest.  The ceiling the end-to-end numbers are held against.
int rmg_h2d_probe(const void *host, size_t bytes_per_gpu, int n_gpus, int repeats, double *gbs, double *aggregate_gbs,
                  char *err, size_t errlen)
{
    std::vector<int> devices;
    int rc = resolve_gpus(n_gpus, devices, err, errlen);
    if (rc) return rc;
    if (!host || bytes_per_gpu == 0 || repeats < 1) return fail(err, errlen, RMG_ERR_INVALID, "bad probe arguments");
    const int G = (int)devices.size();
    std::vector<void *> dst(G, nullptr);
    std::vector<cudaStream_t> st(G, nullptr);
    std::vector<cudaEvent_t> e0(G, nullptr), e1(G, nullptr);
    {
        for (int g = 0; g < G; ++g) {
            RMG_CUDA(cudaSetDevice(devices[g]));
            RMG_CUDA(cudaMalloc(&dst[g], bytes_per_gpu));
            RMG_CUDA(cudaStreamCreateWithFlags(&st[g], cudaStreamNonBlocking));
            RMG_CUDA(cudaEventCreate(&e0[g]));
            RMG_CUDA(cudaEventCreate(&e1[g]));
            RMG_CUDA(cudaMemcpyAsync(dst[g], static_cast<const char *>(host) + (size_t)g * bytes_per_gpu, bytes_per_gpu,
                                     cudaMemcpyHostToDevice, st[g]));      // warm-up
        }
        for (int g = 0; g < G; ++g) {
            RMG_CUDA(cudaSetDevice(devices[g]));
            RMG_CUDA(cudaStreamSynchronize(st[g]));
        }
        const auto t0 = std::chrono::steady_clock::now();
        for (int g = 0; g < G; ++g) {
            RMG_CUDA(cudaSetDevice(devices[g]));
            RMG_CUDA(cudaEventRecord(e0[g], st[g]));
            for (int r = 0; r < repeats; ++r)
                RMG_CUDA(cudaMemcpyAsync(dst[g], static_cast<const char *>(host) + (size_t)g * bytes_per_gpu, bytes_per_gpu,
                                         cudaMemcpyHostToDevice, st[g]));
            RMG_CUDA(cudaEventRecord(e1[g], st[g]));
        }
        for (int g = 0; g < G; ++g) {
            RMG_CUDA(cudaSetDevice(devices[g]));
            RMG_CUDA(cudaStreamSynchronize(st[g]));
        }
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (aggregate_gbs) *aggregate_gbs = (double)bytes_per_gpu * repeats * G / sec / 1e9;
        for (int g = 0; g < G; ++g) {
            float ms = 0.f;
            RMG_CUDA(cudaEventElapsedTime(&ms, e0[g], e1[g]));
            if (gbs) gbs[g] = (double)bytes_per_gpu * repeats / (ms * 1e-3) / 1e9;
        }
    }
cleanup:
    for (int g = 0; g < G; ++g) {
        if (cudaSetDevice(devices[g]) != cudaSuccess) continue;
        if (st[g]) cudaStreamSynchronize(st[g]);
        if (dst[g]) cudaFree(dst[g]);
        if (e0[g]) cudaEventDestroy(e0[g]);
        if (e1[g]) cudaEventDestroy(e1[g]);
        if (st[g]) cudaStreamDestroy(st[g]);
    }
    if (rc != RMG_OK) cudaGetLastError();
    return rc;
}

// ---- mincRandomize on a model with a voxel-wise covariate (y ~ static terms + a column of files).
// The reference's loop re-evaluates the whole mincLm call with the PREDICTOR rows of the data frame re-sampled -- the
// column of covariate files moves with the static columns (R/minc_voxel_statistics.R:1468-1477) -- so permutation r fits
// y_v on P_r [Xs | z_v] at every voxel (src/modelling_functions.c:1118-1144 re-factorises that matrix per voxel).  Here:
// the static part Xs[idx_r, ] is factored once per permutation on the host; on the device the covariate kernel pairs
// subject j of Y with plane idx_r[j] of Z (CovArgs::zrow), result columns reduce to their maxima.  One streaming pass
// over Y and Z per permutation (HBM-bound, 16 n bytes per voxel), voxel shards over the devices, host max.
static int randomize_cov_transient(const double *Y, const double *Z, int64_t V, int64_t ldY, int n, const double *Xs_in, int ps_in,
                                   const double *mask, double mask_lo, double mask_hi, const int32_t *idx, int R,
                                   const int32_t *cols, int ncols, int two_sided, double *dist, const std::vector<int> &devices,
                                   char *err, size_t errlen)
{
    if (V > 0 && (!Y || !Z)) return fail(err, errlen, RMG_ERR_INVALID, "Y or Z is NULL");
    if (n < 1) return fail(err, errlen, RMG_ERR_INVALID, "n must be >= 1 (got %d)", n);
    std::vector<double> ones;
    int ps = ps_in;
    const double *Xs = Xs_in;
    if (!Xs || ps <= 0) {                       // the static part defaults to the intercept (src/vertex_modelling.c:36-39)
        ones.assign((size_t)n, 1.0);
        Xs = ones.data();
        ps = 1;
    }
    const int p = ps + 1;
    int rc = check_fit_args(Y, V, ldY, n, Xs, ps, Y, ldY, err, errlen);
    if (rc) return rc;
    if (ps > 16) return fail(err, errlen, RMG_ERR_INVALID, "voxel-wise covariate designs support at most 16 static columns (got %d)", ps);
    if ((rc = check_perm_args(p, idx, R, cols, ncols, dist, err, errlen))) return rc;
    if (R == 0 || ncols == 0) return RMG_OK;
    PermSet pd;
    if ((rc = factor_permutations(Xs, n, ps, idx, R, pd, err, errlen))) return rc;
    for (int r = 0; r < R; ++r)
        if (pd.design(r).k < ps)
            return fail(err, errlen, RMG_ERR_INVALID,
                        "permutation %d: rank-deficient static design (rank %d < %d columns) with a voxel-wise covariate is not supported",
                        r + 1, pd.design(r).k, ps);
    const int G = (int)devices.size(), nk = R * ncols, ncol_total = 2 * p + 3, KP = lm_padded_p(ps);
    rmg_dataset *dy = nullptr, *dz = nullptr;
    if ((rc = dataset_create(Y, nullptr, V, ldY, n, mask, mask_lo, mask_hi, devices.data(), G, &dy, err, errlen))) return rc;
    if ((rc = dataset_create(Z, nullptr, V, ldY, n, nullptr, 0.0, 0.0, devices.data(), G, &dz, err, errlen))) {
        dataset_destroy(dy);
        return rc;
    }
    // host copies of every permuted design (shared by the shards)
    const size_t qstride = make_qt(pd.base, KP).size();
    std::vector<DevDesign> hdd(R);
    std::vector<double> hqt((size_t)R * qstride);
    parallel_for(R, [&](int r0, int r1) {
        for (int r = r0; r < r1; ++r) {
            const Design Dr = pd.materialise(r);
            fill_dev_design(Dr, hdd[r]);
            const std::vector<double> q = make_qt(Dr, KP);
            std::copy(q.begin(), q.end(), hqt.begin() + (size_t)r * qstride);
        }
    });
    ColList cl;
    std::memset(&cl, 0, sizeof cl);
    for (int c = 0; c < ncols; ++c) cl.c[c] = cols[c];
    std::vector<std::vector<double>> parts(G, std::vector<double>((size_t)nk, -INFINITY));
    std::vector<int> rcs(G, RMG_OK), status(G, 0);
    std::vector<std::string> msgs(G, std::string(256, '\0'));
    std::vector<int64_t> launches(G, 0);
    std::vector<std::thread> th;
    for (int g = 0; g < G; ++g) {
        if (dy->sh[g].cv <= 0) continue;
        th.emplace_back([&, g]() {
            DsShard &sy = dy->sh[g], &sz = dz->sh[g];
            char *e_ = &msgs[g][0];
            const size_t el_ = msgs[g].size();
            auto body = [&](char *err, size_t errlen) -> int {
                int rc = RMG_OK;
                double *scratch = nullptr, *qt = nullptr, *dD = nullptr;
                DevDesign *dd = nullptr;
                int32_t *didx = nullptr;
                int *dstatus = nullptr;
                unsigned long long *keys = nullptr;
                cudaStream_t st = sy.st;
                LmPlan plan;
                plan.sm_count = sy.sm_count;
                const int64_t ldo = (sy.cv + 1) & ~int64_t(1);
                {
                    RMG_CUDA(cudaSetDevice(sy.device));
                    for (cudaEvent_t ev : sy.landed)
                        if (ev) RMG_CUDA(cudaStreamWaitEvent(st, ev, 0));
                    for (cudaEvent_t ev : sz.landed)
                        if (ev) RMG_CUDA(cudaStreamWaitEvent(st, ev, 0));
                    RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&scratch), (size_t)ldo * ncol_total * sizeof(double), st));
                    RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&dd), sizeof(DevDesign) * (size_t)R, st));
                    RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&qt), sizeof(double) * hqt.size(), st));
                    RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&didx), sizeof(int32_t) * (size_t)R * n, st));
                    RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&dstatus), sizeof(int), st));
                    RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&keys), sizeof(unsigned long long) * (size_t)nk, st));
                    RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&dD), sizeof(double) * (size_t)nk, st));
                    RMG_CUDA(cudaMemsetAsync(dstatus, 0, sizeof(int), st));
                    RMG_CUDA(cudaMemcpyAsync(dd, hdd.data(), sizeof(DevDesign) * (size_t)R, cudaMemcpyHostToDevice, st));
                    RMG_CUDA(cudaMemcpyAsync(qt, hqt.data(), sizeof(double) * hqt.size(), cudaMemcpyHostToDevice, st));
                    RMG_CUDA(cudaMemcpyAsync(didx, idx, sizeof(int32_t) * (size_t)R * n, cudaMemcpyHostToDevice, st));
                    keys_init_kernel<<<(nk + 255) / 256, 256, 0, st>>>(keys, nk);
                    ++launches[g];
                    const int cm_grid = (int)std::min<int64_t>((sy.cv + 255) / 256, (int64_t)plan.sm_count * 8);
                    for (int r = 0; r < R; ++r) {
                        CovArgs ca{sy.dY, sz.dY, sy.cv, sy.ld, n, ps, qt + (size_t)r * qstride, dd + r, sy.dM, mask_lo, mask_hi,
                                   scratch, ldo, dstatus, didx + (size_t)r * n};
                        LmLaunchInfo info;
                        RMG_CUDA(launch_lm_cov(ca, plan, st, &info));
                        launches[g] += info.launches;
                        colmax_kernel<<<cm_grid, 256, 0, st>>>(scratch, sy.cv, ldo, cl, ncols, two_sided, keys, R, r);
                        ++launches[g];
                    }
                    RMG_CUDA(cudaGetLastError());
                    keys_decode_kernel<<<(nk + 255) / 256, 256, 0, st>>>(keys, dD, nk);
                    ++launches[g];
                    RMG_CUDA(cudaMemcpyAsync(parts[g].data(), dD, sizeof(double) * (size_t)nk, cudaMemcpyDeviceToHost, st));
                    RMG_CUDA(cudaMemcpyAsync(&status[g], dstatus, sizeof(int), cudaMemcpyDeviceToHost, st));
                    RMG_CUDA(cudaStreamSynchronize(st));
                }
            cleanup:
                void *all[] = {scratch, dd, qt, didx, dstatus, keys, dD};
                for (void *x : all)
                    if (x) cudaFreeAsync(x, st);
                if (rc != RMG_OK) cudaGetLastError();
                return rc;
            };
            rcs[g] = body(e_, el_);
        });
    }
    for (auto &t : th) t.join();
    for (int g = 0; g < G; ++g) g_launch_count += launches[g];
    dataset_destroy(dy);
    dataset_destroy(dz);
    for (int g = 0; g < G; ++g) {
        if (rcs[g] != RMG_OK) return fail(err, errlen, rcs[g], "gpu %d: %s", g, msgs[g].c_str());
        // an all-zero covariate at a fitted voxel: the reference's dpotri error aborts the call (src/modelling_functions.c:551-557)
        if (status[g] > 0) return fail(err, errlen, RMG_ERR_SINGULAR, "%s", singular_message(status[g]).c_str());
    }
    for (int i = 0; i < nk; ++i) {
        double m = -INFINITY;
        for (int g = 0; g < G; ++g)
            if (!parts[g].empty()) m = std::fmax(m, parts[g][i]);
        dist[i] = m;
    }
    return RMG_OK;
}

int rmg_randomize_cov(const double *Y, const double *Z, int64_t V, int64_t ldY, int n, const double *Xs, int ps, const double *mask,
                      double mask_lo, double mask_hi, const int32_t *idx, int R, const int32_t *cols, int ncols, int two_sided,
                      double *dist, int device, char *err, size_t errlen)
{
    int rc = select_device(device, err, errlen);
    if (rc) return rc;
    return randomize_cov_transient(Y, Z, V, ldY, n, Xs, ps, mask, mask_lo, mask_hi, idx, R, cols, ncols, two_sided, dist,
                                   std::vector<int>{device}, err, errlen);
}

int rmg_randomize_cov_multi(const double *Y, const double *Z, int64_t V, int64_t ldY, int n, const double *Xs, int ps,
                            const double *mask, double mask_lo, double mask_hi, const int32_t *idx, int R, const int32_t *cols,
                            int ncols, int two_sided, double *dist, int n_gpus, char *err, size_t errlen)
{
    std::vector<int> devices;
    int rc = resolve_gpus(n_gpus, devices, err, errlen);
    if (rc) return rc;
    return randomize_cov_transient(Y, Z, V, ldY, n, Xs, ps, mask, mask_lo, mask_hi, idx, R, cols, ncols, two_sided, dist, devices,
                                   err, errlen);
}

// ---- mincRandomize on host data = a transient dataset whose upload is consumed block by block as it lands
static int randomize_transient(const void *Yv, const StoredSpec *pk, int64_t V, int64_t ldY, int n, const double *X, int p,
                               const double *mask, double mask_lo, double mask_hi, const int32_t *idx, int R,
                               const int32_t *cols, int ncols, int two_sided, double *dist, const std::vector<int> &devices,
                               char *err, size_t errlen)
{
    int rc = check_fit_args(Yv, V, ldY, n, X, p, Yv, ldY, err, errlen);
    if (rc) return rc;
    if ((rc = check_perm_args(p, idx, R, cols, ncols, dist, err, errlen))) return rc;
    if (R == 0 || ncols == 0) return RMG_OK;
    {   // fail on a singular permuted design before 28 GB cross the link, as the reference fails before any voxel
        PermSet probe;
        if ((rc = factor_permutations(X, n, p, idx, std::min(R, 1), probe, err, errlen))) return rc;
    }
    rmg_dataset *ds = nullptr;
    if ((rc = dataset_create(Yv, pk, V, ldY, n, mask, mask_lo, mask_hi, devices.data(), (int)devices.size(), &ds, err, errlen))) return rc;
    rc = dataset_randomize(ds, X, p, idx, R, cols, ncols, two_sided, dist, err, errlen);
    dataset_destroy(ds);
    return rc;
}

int rmg_randomize(const double *Y, int64_t V, int64_t ldY, int n, const double *X, int p, const double *mask,
                  double mask_lo, double mask_hi, const int32_t *idx, int R, const int32_t *cols, int ncols,
                  int two_sided, double *dist, int device, char *err, size_t errlen)
{
    int rc = select_device(device, err, errlen);
    if (rc) return rc;
    return randomize_transient(Y, nullptr, V, ldY, n, X, p, mask, mask_lo, mask_hi, idx, R, cols, ncols, two_sided, dist,
                               std::vector<int>{device}, err, errlen);
}

int rmg_randomize_stored(const void *U, int dtype, int64_t V, int64_t ldU, int n, const double *scale, const double *offset,
                         double voxel_min, int64_t slice_len, const double *X, int p, const double *mask, double mask_lo,
                         double mask_hi, const int32_t *idx, int R, const int32_t *cols, int ncols, int two_sided, double *dist,
                         int device, char *err, size_t errlen)
{
    int rc = check_stored_args(dtype, scale, offset, voxel_min, slice_len, p, err, errlen);
    if (rc) return rc;
    if ((rc = select_device(device, err, errlen))) return rc;
    const StoredSpec pk{dtype, scale, offset, voxel_min, dtype == RMG_STORED_U16 ? slice_len : std::max<int64_t>(V, 1)};
    return randomize_transient(U, &pk, V, ldU, n, X, p, mask, mask_lo, mask_hi, idx, R, cols, ncols, two_sided, dist,
                               std::vector<int>{device}, err, errlen);
}

int rmg_randomize_multi(const double *Y, int64_t V, int64_t ldY, int n, const double *X, int p, const double *mask,
                        double mask_lo, double mask_hi, const int32_t *idx, int R, const int32_t *cols, int ncols,
                        int two_sided, double *dist, int n_gpus, char *err, size_t errlen)
{
    std::vector<int> devices;
    int rc = resolve_gpus(n_gpus, devices, err, errlen);
    if (rc) return rc;
    return randomize_transient(Y, nullptr, V, ldY, n, X, p, mask, mask_lo, mask_hi, idx, R, cols, ncols, two_sided, dist, devices,
                               err, errlen);
}

int rmg_randomize_stored_multi(const void *U, int dtype, int64_t V, int64_t ldU, int n, const double *scale,
                               const double *offset, double voxel_min, int64_t slice_len, const double *X, int p,
                               const double *mask, double mask_lo, double mask_hi, const int32_t *idx, int R,
                               const int32_t *cols, int ncols, int two_sided, double *dist, int n_gpus, char *err,
                               size_t errlen)
{
    int rc = check_stored_args(dtype, scale, offset, voxel_min, slice_len, p, err, errlen);
    if (rc) return rc;
    std::vector<int> devices;
    if ((rc = resolve_gpus(n_gpus, devices, err, errlen))) return rc;
    const StoredSpec pk{dtype, scale, offset, voxel_min, dtype == RMG_STORED_U16 ? slice_len : std::max<int64_t>(V, 1)};
    return randomize_transient(U, &pk, V, ldU, n, X, p, mask, mask_lo, mask_hi, idx, R, cols, ncols, two_sided, dist, devices,
                               err, errlen);
}

// Device-resident variant: a borrowed shard on the caller's stream; returns without synchronising.
int rmg_randomize_device(const double *dY, int64_t V, int64_t ldY, int n, const double *X, int p,
                         const double *dmask, double mask_lo, double mask_hi, const int32_t *idx, int R,
                         const int32_t *cols, int ncols, int two_sided, double *ddist, int device, void *stream,
                         char *err, size_t errlen)
{
    int rc = check_fit_args(dY, V, ldY, n, X, p, dY, ldY, err, errlen);
    if (rc) return rc;
    if ((rc = check_perm_args(p, idx, R, cols, ncols, ddist, err, errlen))) return rc;
    if ((rc = select_device(device, err, errlen))) return rc;
    if (R == 0 || ncols == 0) return RMG_OK;
    std::vector<PermShard> S(1);
    PermShard &s = S[0];
    s.device = device;
    s.sm_count = device_sm_count(device);
    s.dY = dY;
    s.cv = V;
    s.ld = ldY;
    s.dM = dmask;
    s.st = static_cast<cudaStream_t>(stream);
    s.blk_v0 = {0};
    s.blk_cv = {V};
    s.ddist = ddist;
    if (V == 0) {
        // max over no voxels: the reference's max() of an empty column is -Inf
        std::vector<double> ninf((size_t)R * ncols, -INFINITY);
        cudaError_t e = cudaMemcpyAsync(ddist, ninf.data(), ninf.size() * 8, cudaMemcpyHostToDevice, s.st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s.st);
        return e == cudaSuccess ? RMG_OK : fail(err, errlen, RMG_ERR_CUDA, "%s", cudaGetErrorString(e));
    }
    return randomize_shards(S, n, p, X, mask_lo, mask_hi, idx, R, cols, ncols, two_sided, false, err, errlen);
}

}  // extern "C"
