// capi.cu -- C ABI (include/rminc_b200.h) of the fit path: argument checking, design
// factorisation, host<->device staging pipeline, multi-GPU voxel sharding.
//
// Reference interfaces replaced: minc2_model(method="lm") src/modelling_functions.c:743-745,
// vertex_lm_loop src/vertex_modelling.c:8, and the parallel= sharding of
// R/minc_voxel_statistics.R:848-906.  No CPU fallback exists: every compute entry point
// fails with RMG_ERR_NO_DEVICE when no CUDA device is usable.
#include <algorithm>
#include <cmath>
#include <memory>
#include <mutex>
#include <thread>

#if defined(__linux__)
#include <sys/syscall.h>
#include <unistd.h>
#endif

#include "capi_common.hpp"

namespace rmg {

std::atomic<int64_t> g_launch_count{0};
thread_local double g_last_kernel_ms = 0.0;
thread_local int g_last_fit_variant = 0;

int select_device(int device, char *err, size_t errlen)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        cudaGetLastError();
        return fail(err, errlen, RMG_ERR_NO_DEVICE,
                    "no CUDA device available (%s); rminc_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= count)
        return fail(err, errlen, RMG_ERR_NO_DEVICE, "device %d out of range (have %d)", device, count);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(err, errlen, RMG_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    // Keep stream-ordered allocations in the pool across synchronisations: with the default
    // release threshold (0) every cudaMallocAsync after a sync goes back to the OS (~ms).
    static std::atomic<uint64_t> pool_ready{0};
    if (device < 64 && !(pool_ready.load() & (1ull << device))) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        // experiment knob: L2 fetch granularity in bytes (32 / 64 / 128) for the scattered row segments of masked fits
        if (const char *g = std::getenv("RMG_L2_FETCH")) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)std::atoi(g));
        cudaGetLastError();
        pool_ready.fetch_or(1ull << device);
    }
    return RMG_OK;
}

int host_threads()
{
    static int cached = 0;
    if (cached) return cached;
    int nt = 0;
    if (const char *e = std::getenv("RMG_HOST_THREADS")) nt = std::atoi(e);
    if (nt <= 0) {
        // hardware_concurrency() honours the cpuset of the container.  Under torchrun the ranks of the box share the
        // cores; a single process (R, or rank 0 driving every GPU) owns them all.  (Dividing by the number of VISIBLE
        // GPUs, as round 1 did, left a one-GPU call on an 8-GPU box with an eighth of the cores.)
        const unsigned hw = std::thread::hardware_concurrency();
        int ranks = 1;
        if (const char *e = std::getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, std::atoi(e));
        nt = (int)(hw ? hw : 1) / ranks;
    }
    cached = std::max(1, std::min(nt, 64));
    return cached;
}

// ---- pinned host pool
namespace {
struct PinnedBuf {
    void *ptr;
    size_t bytes;
    bool in_use;
};
std::mutex g_pin_mu;
std::vector<PinnedBuf> g_pin;

// Interleave the pages of the next allocations over every NUMA node this process may allocate on (raw syscalls:
// libnuma is not a dependency).  Best effort; restores the default policy on destruction.
struct ScopedInterleave {
    bool active = false;
    ScopedInterleave()
    {
#if defined(__linux__) && defined(SYS_set_mempolicy) && defined(SYS_get_mempolicy)
        if (std::getenv("RMG_NO_NUMA_INTERLEAVE")) return;
        unsigned long mask[16] = {0};
        // MPOL_F_MEMS_ALLOWED = 4: the nodes the cpuset lets us allocate on
        if (syscall(SYS_get_mempolicy, nullptr, mask, sizeof(mask) * 8, nullptr, 4) != 0) return;
        int nodes = 0;
        for (unsigned long w : mask) nodes += __builtin_popcountl(w);
        if (nodes < 2) return;
        active = syscall(SYS_set_mempolicy, 3 /* MPOL_INTERLEAVE */, mask, sizeof(mask) * 8) == 0;
#endif
    }
    ~ScopedInterleave()
    {
#if defined(__linux__) && defined(SYS_set_mempolicy)
        if (active) syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, nullptr, 0);
#endif
    }
};
}  // namespace

void *pinned_acquire(size_t bytes)
{
    if (bytes == 0) bytes = 16;
    {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        PinnedBuf *best = nullptr;
        for (auto &b : g_pin)
            if (!b.in_use && b.bytes >= bytes && (!best || b.bytes < best->bytes)) best = &b;
        // a much larger idle buffer is not worth holding hostage for a small request
        if (best && best->bytes <= 4 * bytes + (1u << 20)) {
            best->in_use = true;
            return best->ptr;
        }
    }
    void *p = nullptr;
    {
        ScopedInterleave numa;
        if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
    }
    std::lock_guard<std::mutex> lk(g_pin_mu);
    g_pin.push_back({p, bytes, true});
    return p;
}

void pinned_release(void *ptr)
{
    std::lock_guard<std::mutex> lk(g_pin_mu);
    for (auto &b : g_pin)
        if (b.ptr == ptr) b.in_use = false;
}

void pinned_trim()
{
    std::vector<void *> dead;
    {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        for (size_t i = 0; i < g_pin.size();)
            if (!g_pin[i].in_use) {
                dead.push_back(g_pin[i].ptr);
                g_pin.erase(g_pin.begin() + i);
            } else {
                ++i;
            }
    }
    for (void *p : dead) cudaFreeHost(p);
    cudaGetLastError();
}

int device_sm_count(int device)
{
    int sms = 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
        cudaGetLastError();
        sms = 148;
    }
    return sms;
}

int check_fit_args(const void *Y, int64_t V, int64_t ldY, int n, const void *X, int p, const void *out,
                   int64_t ldOut, char *err, size_t errlen)
{
    if (!X) return fail(err, errlen, RMG_ERR_INVALID, "X is NULL");
    if (V < 0) return fail(err, errlen, RMG_ERR_INVALID, "V < 0");
    if (V > 0 && (!Y || !out)) return fail(err, errlen, RMG_ERR_INVALID, "Y or out is NULL");
    if (n < 1) return fail(err, errlen, RMG_ERR_INVALID, "n must be >= 1 (got %d)", n);
    if (p < 1 || p > RMG_MAX_P) return fail(err, errlen, RMG_ERR_INVALID, "p must be in 1..%d (got %d)", RMG_MAX_P, p);
    if (ldY < V || ldOut < V) return fail(err, errlen, RMG_ERR_INVALID, "leading dimension smaller than V");
    return RMG_OK;
}

void fill_dev_design(const Design &D, DevDesign &dd)
{
    std::memset(&dd, 0, sizeof dd);
    dd.n = D.n;
    dd.p = D.p;
    dd.k = D.k;
    dd.shift = D.has_intercept ? 1 : 0;
    dd.gap = D.gap;
    dd.inv_n = 1.0 / D.n;
    dd.rdf = (double)(D.n - D.p);
    dd.pm1 = (double)(D.p - 1);
    dd.loglik_c = std::log(2 * M_PI) + 1 - std::log((double)D.n);
    dd.mhalf_n = -.5 * D.n;
    dd.loglik_add = D.loglik_add;
    if (!D.rowscale.empty()) dd.inv_n = 1.0 / D.sumw;   // weighted mean of the fitted values
    dd.mode = 0;
    dd.nterms = 0;
    dd.ncol_out = 2 * D.p + 3;
    for (int i = 0; i < D.p; ++i) {
        dd.q1[i] = D.q1[i];
        dd.ct[i] = D.ct[i];
        for (int j = 0; j < D.p; ++j) dd.Rinv[i * kMaxPDev + j] = D.Rinv[i + (size_t)j * D.p];
    }
}

// voxel_anova (src/minc_anova.c:4-80): effects[i] (pivoted position i) is booked under asgn[i].
void set_anova_mode(const Design &D, const int32_t *asgn, DevDesign &dd)
{
    int nterms = 0;
    for (int i = 0; i < D.p; ++i) nterms = std::max(nterms, (int)asgn[i]);
    dd.mode = 1;
    dd.nterms = nterms;
    dd.ncol_out = nterms;
    std::vector<int> df(nterms + 1, 0);
    for (int i = 0; i < D.p; ++i) {
        dd.term[i] = asgn[i];
        df[asgn[i]]++;
    }
    for (int t = 0; t <= nterms; ++t) dd.inv_df[t] = df[t] ? 1.0 / df[t] : 0.0;   // 0/0 -> NaN as the reference's ss/df
}

std::vector<double> make_qt(const Design &D, int KP)
{
    // rows padded with zeros to a multiple of 16 subjects: the TMA kernel bulk-copies whole stages
    const size_t rows = ((size_t)D.n + 15) / 16 * 16;
    std::vector<double> qt(rows * KP, 0.0);
    for (int c = 0; c < D.k; ++c)
        for (int i = 0; i < D.n; ++i) qt[(size_t)i * KP + c] = D.Q[i + (size_t)c * D.n];
    return qt;
}

int factor_or_fail(const double *X, int n, int p, Design &D, char *err, size_t errlen, const double *w)
{
    for (size_t i = 0; i < (size_t)n * p; ++i)
        if (!std::isfinite(X[i])) return fail(err, errlen, RMG_ERR_INVALID, "model matrix contains non-finite values");
    if (w)
        for (int i = 0; i < n; ++i)
            if (!(w[i] > 0.0) || !std::isfinite(w[i]))
                return fail(err, errlen, RMG_ERR_INVALID, "weights must be finite and > 0 (weights[%d] = %g)", i, w[i]);
    const int info = factor_design(X, n, p, D, w);
    if (info > 0) return fail(err, errlen, RMG_ERR_SINGULAR, "%s", singular_message(info).c_str());
    return RMG_OK;
}

namespace {
std::mutex g_cache_mu;
std::vector<std::shared_ptr<CachedDesign>> g_cache;
uint64_t g_cache_tick = 0;
constexpr size_t kCacheEntries = 16;

}  // namespace

// Looks up (or builds) the cached design for X on the current device.  rc != RMG_OK on error.
std::shared_ptr<CachedDesign> design_cache_get(int device, const double *X, int n, int p, const int32_t *asgn,
                                               int &rc, char *err, size_t errlen, const double *w)
{
    const size_t cnt = (size_t)n * p;
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        for (auto &e : g_cache)
            if (e->device == device && e->D.n == n && e->D.p == p && std::memcmp(e->X.data(), X, cnt * 8) == 0 &&
                e->asgn.size() == (asgn ? (size_t)p : 0) &&
                (!asgn || std::memcmp(e->asgn.data(), asgn, (size_t)p * sizeof(int32_t)) == 0) &&
                e->w.size() == (w ? (size_t)n : 0) && (!w || std::memcmp(e->w.data(), w, (size_t)n * 8) == 0)) {
                e->last_use = ++g_cache_tick;
                rc = RMG_OK;
                return e;
            }
    }
    auto e = std::make_shared<CachedDesign>();
    if ((rc = factor_or_fail(X, n, p, e->D, err, errlen, w))) return nullptr;
    if (w) e->w.assign(w, w + n);
    e->device = device;
    e->X.assign(X, X + cnt);
    DevDesign h;
    fill_dev_design(e->D, h);
    if (asgn) {
        for (int i = 0; i < p; ++i)
            if (asgn[i] < 0 || asgn[i] > p) {
                rc = fail(err, errlen, RMG_ERR_INVALID, "asgn[%d] = %d out of range", i, asgn[i]);
                return nullptr;
            }
        if (e->D.k < p) {
            rc = fail(err, errlen, RMG_ERR_INVALID,
                      "rank-deficient design (rank %d < %d columns): the anova path needs a full-rank model matrix", e->D.k, p);
            return nullptr;
        }
        e->asgn.assign(asgn, asgn + p);
        set_anova_mode(e->D, asgn, h);
    }
    const std::vector<double> qt_h = make_qt(e->D, lm_padded_p(p));
    cudaStream_t up = nullptr;
    cudaError_t ce = cudaMalloc(reinterpret_cast<void **>(&e->dd), sizeof(DevDesign));
    if (ce == cudaSuccess) ce = cudaMalloc(reinterpret_cast<void **>(&e->qt), qt_h.size() * sizeof(double));
    if (ce == cudaSuccess && w) ce = cudaMalloc(reinterpret_cast<void **>(&e->rowscale), (size_t)n * sizeof(double));
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&e->ready, cudaEventDisableTiming);
    // upload on a private stream so a miss never serialises with the caller's queued work
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(e->dd, &h, sizeof h, cudaMemcpyHostToDevice, up);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(e->qt, qt_h.data(), qt_h.size() * sizeof(double), cudaMemcpyHostToDevice, up);
    if (ce == cudaSuccess && w)
        ce = cudaMemcpyAsync(e->rowscale, e->D.rowscale.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice, up);
    if (ce == cudaSuccess) ce = cudaEventRecord(e->ready, up);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(up);   // pageable sources die with this scope
    if (up) cudaStreamDestroy(up);
    if (ce != cudaSuccess) {
        rc = fail(err, errlen, RMG_ERR_CUDA, "uploading the design failed: %s", cudaGetErrorString(ce));
        cudaGetLastError();
        return nullptr;
    }
    std::lock_guard<std::mutex> lk(g_cache_mu);
    e->last_use = ++g_cache_tick;
    if (g_cache.size() >= kCacheEntries) {
        auto victim = std::min_element(g_cache.begin(), g_cache.end(),
                                       [](const auto &a, const auto &b) { return a->last_use < b->last_use; });
        g_cache.erase(victim);   // freed when the last user drops its reference
    }
    g_cache.push_back(e);
    rc = RMG_OK;
    return e;
}

LmPlan plan_from_env(int device)
{
    LmPlan plan;
    plan.sm_count = device_sm_count(device);
    if (const char *v = std::getenv("RMG_FIT_VARIANT")) {
        if (!std::strcmp(v, "tma")) plan.variant = LmVariant::Tma;
        else if (!std::strcmp(v, "ldg")) plan.variant = LmVariant::Ldg;
    }
    if (const char *s = std::getenv("RMG_FIT_SPLIT")) plan.split = std::atoi(s);
    if (const char *s = std::getenv("RMG_TMA_CFG")) plan.tma_cfg = std::atoi(s);
    if (const char *s = std::getenv("RMG_TMA_QSTREAM")) plan.q_stream = std::atoi(s);
    if (const char *s = std::getenv("RMG_NO_COMPACT")) plan.no_compact = std::atoi(s);
    if (const char *s = std::getenv("RMG_STORED_EXACT")) plan.stored_exact = std::atoi(s);
    return plan;
}

}  // namespace rmg

using namespace rmg;

extern "C" {

int rmg_version(void) { return 100; }
int rmg_max_p(void) { return RMG_MAX_P; }

int rmg_device_count(void)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return count;
}

double rmg_last_kernel_ms(void) { return g_last_kernel_ms; }
int rmg_last_fit_variant(void) { return g_last_fit_variant; }
int64_t rmg_launch_count(void) { return g_launch_count.load(); }

int rmg_design_factor(const double *X, int n, int p, int32_t *pivot, int *rank, double *Q, double *Rinv,
                      double *cdiag, int *dpotri_info, char *err, size_t errlen)
{
    if (!X || n < 1 || p < 1 || p > RMG_MAX_P) return fail(err, errlen, RMG_ERR_INVALID, "bad design arguments");
    Design D;
    const int info = factor_design(X, n, p, D);
    if (pivot) std::copy(D.pivot.begin(), D.pivot.end(), pivot);
    if (rank) *rank = D.k;
    if (Q) std::copy(D.Q.begin(), D.Q.end(), Q);
    if (Rinv) std::copy(D.Rinv.begin(), D.Rinv.end(), Rinv);
    if (cdiag) std::copy(D.cdiag.begin(), D.cdiag.end(), cdiag);
    if (dpotri_info) *dpotri_info = info;
    if (info > 0) return fail(err, errlen, RMG_ERR_SINGULAR, "%s", singular_message(info).c_str());
    return RMG_OK;
}

static int fit_device_impl(const double *dY, int64_t V, int64_t ldY, int n, const double *X, int p, const int32_t *asgn,
                           const double *dmask, double mask_lo, double mask_hi, double *dout, int64_t ldOut,
                           int device, void *stream, char *err, size_t errlen)
{
    int rc = check_fit_args(dY, V, ldY, n, X, p, dout, ldOut, err, errlen);
    if (rc) return rc;
    if ((rc = select_device(device, err, errlen))) return rc;
    std::shared_ptr<CachedDesign> cd = design_cache_get(device, X, n, p, asgn, rc, err, errlen);
    if (!cd) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    LmScratch scr;
    LmLaunchInfo info;
    {
        RMG_CUDA(cudaStreamWaitEvent(st, cd->ready, 0));
        LmArgs a{dY, V, ldY, n, p, cd->qt, cd->dd, dmask, mask_lo, mask_hi, dout, ldOut};
        a.ncol_out = asgn ? (int)*std::max_element(cd->asgn.begin(), cd->asgn.end()) : 2 * p + 3;
        if (a.ncol_out == 0) goto cleanup;   // intercept-only anova: no terms, nothing to write
        RMG_CUDA(launch_lm_fit(a, plan_from_env(device), scr, st, &info));
        g_launch_count += info.launches;
        g_last_fit_variant = info.used_tma ? 1 : 2;
    }
cleanup:
    scr.release(st);
    return rc;
}

int rmg_lm_fit_device(const double *dY, int64_t V, int64_t ldY, int n, const double *X, int p,
                      const double *dmask, double mask_lo, double mask_hi, double *dout, int64_t ldOut,
                      int device, void *stream, char *err, size_t errlen)
{
    return fit_device_impl(dY, V, ldY, n, X, p, nullptr, dmask, mask_lo, mask_hi, dout, ldOut, device, stream, err, errlen);
}

int rmg_anova_fit_device(const double *dY, int64_t V, int64_t ldY, int n, const double *X, int p, const int32_t *asgn,
                         const double *dmask, double mask_lo, double mask_hi, double *dout, int64_t ldOut,
                         int device, void *stream, char *err, size_t errlen)
{
    if (!asgn) return fail(err, errlen, RMG_ERR_INVALID, "asgn is NULL");
    return fit_device_impl(dY, V, ldY, n, X, p, asgn, dmask, mask_lo, mask_hi, dout, ldOut, device, stream, err, errlen);
}

static int fit_host_impl(const double *Y, int64_t V, int64_t ldY, int n, const double *X, int p, const int32_t *asgn,
                         const double *mask, double mask_lo, double mask_hi, double *out, int64_t ldOut, int device,
                         char *err, size_t errlen, const double *w = nullptr, const double *Z = nullptr,
                         const StoredSpec *pk = nullptr)
{
    const size_t es = !pk ? sizeof(double) : pk->dtype == kStoredU16 ? 2 : 4;   // bytes per sample of Y
    const char *Yb = reinterpret_cast<const char *>(Y);
    NvtxPhase nv("rmg fit (host buffers): chunked H2D -> kernel -> D2H");
    // Z != nullptr: voxel-wise covariate design [X | z_v]; X is the static part (lm_cov.cu)
    int rc = check_fit_args(Y, V, ldY, n, X, p, out, ldOut, err, errlen);
    if (rc) return rc;
    if (Z && p + 1 > 17) return fail(err, errlen, RMG_ERR_INVALID, "voxel-wise covariate designs support at most 16 static columns (got %d)", p);
    {   // the reference fails on a singular design before it touches any voxel; so do we, GPU or not
        Design probe;
        if (rmg_device_count() <= 0 && (rc = factor_or_fail(X, n, p, probe, err, errlen, w))) return rc;
    }
    if ((rc = select_device(device, err, errlen))) return rc;
    std::shared_ptr<CachedDesign> cd = design_cache_get(device, X, n, p, asgn, rc, err, errlen, w);
    if (!cd) return rc;
    if (Z && cd->D.k < p)
        return fail(err, errlen, RMG_ERR_INVALID,
                    "rank-deficient static design (rank %d < %d columns) with a voxel-wise covariate is not supported", cd->D.k, p);
    g_last_kernel_ms = 0.0;
    const int ncol = Z ? 2 * (p + 1) + 3 : asgn ? (int)*std::max_element(cd->asgn.begin(), cd->asgn.end()) : 2 * p + 3;
    if (V == 0 || ncol == 0) return RMG_OK;

    // Voxel chunks of ~1 GiB of Y (measured on B200/PCIe5: 256 MiB chunks reach 61 % of the pinned
    // copy rate, 1 GiB chunks 98 %), multiples of 1024 voxels so TMA tiles stay aligned.
    int64_t CV = (int64_t)(1024.0 * 1024 * 1024 / ((double)es * n));
    CV = std::max<int64_t>(1024, (CV / 1024) * 1024);
    if (const char *s = std::getenv("RMG_CHUNK_VOXELS")) CV = std::max<int64_t>(2, (std::atoll(s) / 2) * 2);
    if (CV > V) CV = V;
    const int64_t nchunks = (V + CV - 1) / CV;
    const int NB = (int)std::min<int64_t>(3, nchunks);
    const int64_t ldc = (CV + 7) & ~int64_t(7);  // rows stay 16-byte aligned for every sample type
    const int64_t slice_len = pk ? std::max<int64_t>(1, pk->slice_len) : 1;
    const int64_t nslices = !pk ? 0 : pk->nslices_total > 0 ? pk->nslices_total : (V + slice_len - 1) / slice_len;
    double *dScale = nullptr, *dOffset = nullptr;

    cudaStream_t st[3] = {nullptr, nullptr, nullptr};
    double *dY[3] = {nullptr, nullptr, nullptr}, *dO[3] = {nullptr, nullptr, nullptr}, *dM[3] = {nullptr, nullptr, nullptr};
    double *dZ[3] = {nullptr, nullptr, nullptr};
    int *dstatus = nullptr;
    std::vector<cudaEvent_t> ev0(nchunks, nullptr), ev1(nchunks, nullptr);
    LmScratch scr[3];
    const LmPlan plan = plan_from_env(device);
    int64_t launches = 0;
    PinnedRing ring;
    // (small transfers are not worth four cudaHostAlloc calls)
    // RMG_STAGING_ROWS=k forces the staging ring with k subjects per slot at any size (tests)
    const int forced_rows = std::getenv("RMG_STAGING_ROWS") ? std::atoi(std::getenv("RMG_STAGING_ROWS")) : 0;
    const bool staged = !is_pinned_host(Y) && !std::getenv("RMG_NO_STAGING") &&
                        (forced_rows > 0 || (double)V * n * (double)es >= 256.0 * 1024 * 1024);
    // rows (subjects) per staging slot: ~128 MiB
    const int rows_per_slot = forced_rows > 0 ? std::min(forced_rows, n)
        : (int)std::max<int64_t>(1, std::min<int64_t>(n, (128ll << 20) / ((int64_t)es * std::max<int64_t>(CV, 1))));
    {
        if (staged) RMG_CUDA(ring.init((size_t)rows_per_slot * (size_t)std::min(CV, V) * es));
        if (pk && pk->dtype == kStoredU16) {
            const size_t cnt = (size_t)nslices * n * sizeof(double);
            RMG_CUDA(cudaMalloc(&dScale, cnt));
            RMG_CUDA(cudaMalloc(&dOffset, cnt));
            RMG_CUDA(cudaMemcpy(dScale, pk->scale, cnt, cudaMemcpyHostToDevice));
            RMG_CUDA(cudaMemcpy(dOffset, pk->offset, cnt, cudaMemcpyHostToDevice));
        }
        for (int b = 0; b < NB; ++b) {
            RMG_CUDA(cudaStreamCreateWithFlags(&st[b], cudaStreamNonBlocking));
            // stream-ordered allocations: the pool keeps the chunk buffers across calls (release threshold set in
            // select_device), so a repeated call pays no cudaMalloc / cudaFree of gigabyte buffers
            RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&dY[b]), (size_t)ldc * n * es, st[b]));
            RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&dO[b]), (size_t)ldc * ncol * sizeof(double), st[b]));
            if (mask) RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&dM[b]), (size_t)ldc * sizeof(double), st[b]));
            if (Z) RMG_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&dZ[b]), (size_t)ldc * n * sizeof(double), st[b]));
        }
        if (Z) {
            RMG_CUDA(cudaMalloc(&dstatus, sizeof(int)));
            RMG_CUDA(cudaMemset(dstatus, 0, sizeof(int)));
        }
        for (int b = 0; b < NB; ++b) RMG_CUDA(cudaStreamWaitEvent(st[b], cd->ready, 0));
        for (int64_t c = 0; c < nchunks; ++c) {
            const int b = (int)(c % NB);
            const int64_t v0 = c * CV, cv = std::min(CV, V - v0);
            RMG_CUDA(cudaEventCreate(&ev0[c]));
            RMG_CUDA(cudaEventCreate(&ev1[c]));
            for (int operand = 0; operand < (Z ? 2 : 1); ++operand) {
                const char *src = operand ? reinterpret_cast<const char *>(Z) : Yb;
                char *dstdev = reinterpret_cast<char *>(operand ? dZ[b] : dY[b]);
                if (!staged) {
                    RMG_CUDA(cudaMemcpy2DAsync(dstdev, (size_t)ldc * es, src + (size_t)v0 * es, (size_t)ldY * es, (size_t)cv * es,
                                               (size_t)n, cudaMemcpyHostToDevice, st[b]));
                    continue;
                }
                for (int r0 = 0; r0 < n; r0 += rows_per_slot) {
                    const int rows = std::min(rows_per_slot, n - r0);
                    int slot = 0;
                    RMG_CUDA(ring.acquire(slot));
                    char *dst = ring.buf[slot];
                    // all host threads copy the strided rows of this group into the pinned slot
                    const int64_t pieces = (int64_t)rows * 8;   // 8 column pieces per row keep every thread busy
                    parallel_for((int)pieces, [&](int i0, int i1) {
                        for (int i = i0; i < i1; ++i) {
                            const int r = i / 8, part = i % 8;
                            const int64_t a0 = cv * part / 8, a1 = cv * (part + 1) / 8;
                            staging_copy(dst + ((size_t)r * cv + a0) * es, src + ((size_t)v0 + (size_t)(r0 + r) * ldY + a0) * es,
                                         (size_t)(a1 - a0) * es);
                        }
                    });
                    RMG_CUDA(cudaMemcpy2DAsync(dstdev + (size_t)r0 * ldc * es, (size_t)ldc * es, dst, (size_t)cv * es, (size_t)cv * es,
                                               (size_t)rows, cudaMemcpyHostToDevice, st[b]));
                    RMG_CUDA(ring.mark(slot, device, st[b]));
                }
            }
            if (mask) RMG_CUDA(cudaMemcpyAsync(dM[b], mask + v0, (size_t)cv * 8, cudaMemcpyHostToDevice, st[b]));
            LmArgs a{dY[b], cv, ldc, n, p, cd->qt, cd->dd, mask ? dM[b] : nullptr, mask_lo, mask_hi, dO[b], ldc};
            a.ncol_out = ncol;
            a.rowscale = cd->rowscale;
            LmLaunchInfo info;
            RMG_CUDA(cudaEventRecord(ev0[c], st[b]));
            if (Z) {
                CovArgs ca{dY[b], dZ[b], cv, ldc, n, p, cd->qt, cd->dd, mask ? dM[b] : nullptr, mask_lo, mask_hi, dO[b], ldc, dstatus};
                RMG_CUDA(launch_lm_cov(ca, plan, st[b], &info));
            } else if (pk) {
                PackedArgs pa{dY[b], pk->dtype, cv, ldc, n, p, cd->qt, cd->dd, mask ? dM[b] : nullptr, mask_lo, mask_hi, dO[b], ldc,
                              dScale, dOffset, 4503599627370496.0 + pk->voxel_min, slice_len, nslices, pk->v_offset + v0};
                RMG_CUDA(launch_lm_packed(pa, plan, scr[b], st[b], &info));
            } else {
                RMG_CUDA(launch_lm_fit(a, plan, scr[b], st[b], &info));
            }
            RMG_CUDA(cudaEventRecord(ev1[c], st[b]));
            launches += info.launches;
            g_last_fit_variant = Z ? 3 : pk ? (info.used_tma ? 5 : 4) : info.used_tma ? 1 : 2;
            RMG_CUDA(cudaMemcpy2DAsync(out + v0, (size_t)ldOut * 8, dO[b], (size_t)ldc * 8, (size_t)cv * 8, (size_t)ncol,
                                       cudaMemcpyDeviceToHost, st[b]));
        }
        for (int b = 0; b < NB; ++b) RMG_CUDA(cudaStreamSynchronize(st[b]));
        double ms = 0.0;
        for (int64_t c = 0; c < nchunks; ++c) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, ev0[c], ev1[c]) == cudaSuccess) ms += t;
        }
        g_last_kernel_ms = ms;
        g_launch_count += launches;
        if (Z) {
            int status = 0;
            RMG_CUDA(cudaMemcpy(&status, dstatus, sizeof(int), cudaMemcpyDeviceToHost));
            // an all-zero covariate at a fitted voxel: exact zero on R's diagonal, the reference's dpotri error
            if (status > 0) rc = fail(err, errlen, RMG_ERR_SINGULAR, "%s", singular_message(status).c_str());
        }
    }
cleanup:
    for (int b = 0; b < 3; ++b) {
        if (st[b]) cudaStreamSynchronize(st[b]);
    }
    for (int b = 0; b < 3; ++b) {
        scr[b].release(st[b]);
        if (st[b]) cudaStreamSynchronize(st[b]);
        if (dY[b]) cudaFreeAsync(dY[b], st[b]);
        if (dO[b]) cudaFreeAsync(dO[b], st[b]);
        if (dM[b]) cudaFreeAsync(dM[b], st[b]);
        if (dZ[b]) cudaFreeAsync(dZ[b], st[b]);
        if (st[b]) cudaStreamSynchronize(st[b]);
        if (st[b]) cudaStreamDestroy(st[b]);
    }
    if (dstatus) cudaFree(dstatus);
    if (dScale) cudaFree(dScale);
    if (dOffset) cudaFree(dOffset);
    for (auto e : ev0) if (e) cudaEventDestroy(e);
    for (auto e : ev1) if (e) cudaEventDestroy(e);
    ring.release_all();
    if (rc != RMG_OK) cudaGetLastError();
    return rc;
}

int rmg_lm_fit(const double *Y, int64_t V, int64_t ldY, int n, const double *X, int p, const double *mask,
               double mask_lo, double mask_hi, double *out, int64_t ldOut, int device, char *err, size_t errlen)
{
    return fit_host_impl(Y, V, ldY, n, X, p, nullptr, mask, mask_lo, mask_hi, out, ldOut, device, err, errlen);
}

int rmg_anova_fit(const double *Y, int64_t V, int64_t ldY, int n, const double *X, int p, const int32_t *asgn,
                  const double *mask, double mask_lo, double mask_hi, double *out, int64_t ldOut, int device,
                  char *err, size_t errlen)
{
    if (!asgn) return fail(err, errlen, RMG_ERR_INVALID, "asgn is NULL");
    return fit_host_impl(Y, V, ldY, n, X, p, asgn, mask, mask_lo, mask_hi, out, ldOut, device, err, errlen);
}

int rmg_vertex_wlm(const double *Y, int64_t V, int64_t ldY, int n, const double *X, int p, const double *weights,
                   double *out, int64_t ldOut, int device, char *err, size_t errlen)
{
    if (!weights) return fail(err, errlen, RMG_ERR_INVALID, "weights is NULL");
    return fit_host_impl(Y, V, ldY, n, X, p, nullptr, nullptr, 0.0, 0.0, out, ldOut, device, err, errlen, weights);
}

int rmg_vertex_lm(const double *Y, int64_t V, int64_t ldY, int n, const double *X, int p, double *out,
                  int64_t ldOut, int device, char *err, size_t errlen)
{
    // vertex_lm_loop has no mask and no I/O (src/vertex_modelling.c:104-158)
    return rmg_lm_fit(Y, V, ldY, n, X, p, nullptr, 0.0, 0.0, out, ldOut, device, err, errlen);
}

int rmg_lm_fit_stored(const void *U, int dtype, int64_t V, int64_t ldU, int n, const double *scale, const double *offset,
                      double voxel_min, int64_t slice_len, const double *X, int p, const double *mask, double mask_lo,
                      double mask_hi, double *out, int64_t ldOut, int device, char *err, size_t errlen)
{
    int rc = check_stored_args(dtype, scale, offset, voxel_min, slice_len, p, err, errlen);
    if (rc) return rc;
    const StoredSpec pk{dtype, scale, offset, voxel_min, slice_len};
    return fit_host_impl(static_cast<const double *>(U), V, ldU, n, X, p, nullptr, mask, mask_lo, mask_hi, out, ldOut, device,
                         err, errlen, nullptr, nullptr, &pk);
}

int rmg_lm_fit_stored_device(const void *dU, int dtype, int64_t V, int64_t ldU, int n, const double *dscale,
                             const double *doffset, double voxel_min, int64_t slice_len, int64_t nslices,
                             int64_t v_offset, const double *X, int p,
                             const double *dmask, double mask_lo, double mask_hi, double *dout, int64_t ldOut, int device,
                             void *stream, char *err, size_t errlen)
{
    int rc = check_fit_args(dU, V, ldU, n, X, p, dout, ldOut, err, errlen);
    if (rc) return rc;
    if ((rc = check_stored_args(dtype, dscale, doffset, voxel_min, slice_len, p, err, errlen))) return rc;
    if ((rc = select_device(device, err, errlen))) return rc;
    std::shared_ptr<CachedDesign> cd = design_cache_get(device, X, n, p, nullptr, rc, err, errlen);
    if (!cd) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    LmLaunchInfo info;
    LmScratch scr;
    {
        RMG_CUDA(cudaStreamWaitEvent(st, cd->ready, 0));
        const int64_t sl = dtype == RMG_STORED_U16 ? slice_len : std::max<int64_t>(V, 1);
        if (v_offset < 0) { rc = fail(err, errlen, RMG_ERR_INVALID, "v_offset < 0"); goto cleanup; }
        PackedArgs pa{dU, dtype, V, ldU, n, p, cd->qt, cd->dd, dmask, mask_lo, mask_hi, dout, ldOut,
                      dscale, doffset, 4503599627370496.0 + voxel_min, sl, nslices > 0 ? nslices : (v_offset + V + sl - 1) / sl,
                      v_offset};
        RMG_CUDA(launch_lm_packed(pa, plan_from_env(device), scr, st, &info));
        g_launch_count += info.launches;
        g_last_fit_variant = info.used_tma ? 5 : 4;
    }
cleanup:
    scr.release(st);
    return rc;
}

// Voxel-wise covariate: the static part defaults to the intercept column (src/vertex_modelling.c:36-39)
static const double *static_part_or_ones(const double *Xs, int &ps, int n, std::vector<double> &ones)
{
    if (Xs && ps > 0) return Xs;
    ones.assign((size_t)std::max(n, 1), 1.0);
    ps = 1;
    return ones.data();
}

int rmg_lm_fit_cov(const double *Y, const double *Z, int64_t V, int64_t ldY, int n, const double *Xs, int ps,
                   const double *mask, double mask_lo, double mask_hi, double *out, int64_t ldOut, int device,
                   char *err, size_t errlen)
{
    if (V > 0 && !Z) return fail(err, errlen, RMG_ERR_INVALID, "Z is NULL");
    if (n < 1) return fail(err, errlen, RMG_ERR_INVALID, "n must be >= 1 (got %d)", n);
    std::vector<double> ones;
    const double *X = static_part_or_ones(Xs, ps, n, ones);
    static const double dummy = 0.0;
    return fit_host_impl(Y, V, ldY, n, X, ps, nullptr, mask, mask_lo, mask_hi, out, ldOut, device, err, errlen, nullptr,
                         Z ? Z : &dummy);
}

int rmg_vertex_lm_cov(const double *Y, const double *Z, int64_t V, int64_t ldY, int n, const double *Xs, int ps,
                      double *out, int64_t ldOut, int device, char *err, size_t errlen)
{
    return rmg_lm_fit_cov(Y, Z, V, ldY, n, Xs, ps, nullptr, 0.0, 0.0, out, ldOut, device, err, errlen);
}

int rmg_lm_fit_cov_device(const double *dY, const double *dZ, int64_t V, int64_t ldY, int n, const double *Xs, int ps,
                          const double *dmask, double mask_lo, double mask_hi, double *dout, int64_t ldOut,
                          int32_t *dstatus, int device, void *stream, char *err, size_t errlen)
{
    if (V > 0 && !dZ) return fail(err, errlen, RMG_ERR_INVALID, "Z is NULL");
    if (n < 1) return fail(err, errlen, RMG_ERR_INVALID, "n must be >= 1 (got %d)", n);
    std::vector<double> ones;
    const double *X = static_part_or_ones(Xs, ps, n, ones);
    int rc = check_fit_args(dY, V, ldY, n, X, ps, dout, ldOut, err, errlen);
    if (rc) return rc;
    if (ps > 16) return fail(err, errlen, RMG_ERR_INVALID, "voxel-wise covariate designs support at most 16 static columns (got %d)", ps);
    if ((rc = select_device(device, err, errlen))) return rc;
    std::shared_ptr<CachedDesign> cd = design_cache_get(device, X, n, ps, nullptr, rc, err, errlen);
    if (!cd) return rc;
    if (cd->D.k < ps)
        return fail(err, errlen, RMG_ERR_INVALID,
                    "rank-deficient static design (rank %d < %d columns) with a voxel-wise covariate is not supported", cd->D.k, ps);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    LmLaunchInfo info;
    {
        RMG_CUDA(cudaStreamWaitEvent(st, cd->ready, 0));
        CovArgs ca{dY, dZ, V, ldY, n, ps, cd->qt, cd->dd, dmask, mask_lo, mask_hi, dout, ldOut, dstatus};
        RMG_CUDA(launch_lm_cov(ca, plan_from_env(device), st, &info));
        g_launch_count += info.launches;
        g_last_fit_variant = 3;
    }
cleanup:
    return rc;
}

}  // extern "C"

// Contiguous voxel shards over devices 0..G-1, one host thread per GPU (R is one process and must never fork a CUDA
// context).  fn(g, v0, cv, err, errlen) fits voxels [v0, v0 + cv) on device g.
template <class Fn>
static int run_sharded(int64_t V, int n_gpus, char *err, size_t errlen, Fn fn)
{
    const int have = rmg_device_count();
    if (have <= 0) return fail(err, errlen, RMG_ERR_NO_DEVICE, "no CUDA device available; rminc_b200 has no CPU fallback");
    const int G = n_gpus <= 0 ? have : n_gpus;
    if (G > have) return fail(err, errlen, RMG_ERR_NO_DEVICE, "%d GPUs requested, %d visible", G, have);
    // even boundaries keep every shard 16-byte aligned
    std::vector<int64_t> b(G + 1);
    for (int g = 0; g <= G; ++g) b[g] = std::min<int64_t>(V, ((V * g / G) + 1) & ~int64_t(1));
    b[0] = 0;
    b[G] = V;
    std::vector<int> rcs(G, RMG_OK);
    std::vector<std::string> msgs(G, std::string(256, '\0'));
    std::vector<double> kms(G, 0.0);
    std::vector<std::thread> th;
    for (int g = 0; g < G; ++g) {
        th.emplace_back([&, g]() {
            const int64_t v0 = b[g], cv = b[g + 1] - b[g];
            if (cv <= 0) return;
            rcs[g] = fn(g, v0, cv, &msgs[g][0], msgs[g].size());
            kms[g] = rmg_last_kernel_ms();
        });
    }
    for (auto &t : th) t.join();
    g_last_kernel_ms = *std::max_element(kms.begin(), kms.end());
    for (int g = 0; g < G; ++g)
        if (rcs[g] != RMG_OK) return fail(err, errlen, rcs[g], "gpu %d: %s", g, msgs[g].c_str());
    return RMG_OK;
}

extern "C" {

int rmg_lm_fit_multi(const double *Y, int64_t V, int64_t ldY, int n, const double *X, int p, const double *mask,
                     double mask_lo, double mask_hi, double *out, int64_t ldOut, int n_gpus, char *err,
                     size_t errlen)
{
    int rc = check_fit_args(Y, V, ldY, n, X, p, out, ldOut, err, errlen);
    if (rc) return rc;
    return run_sharded(V, n_gpus, err, errlen, [&](int g, int64_t v0, int64_t cv, char *e, size_t el) {
        return rmg_lm_fit(Y + v0, cv, ldY, n, X, p, mask ? mask + v0 : nullptr, mask_lo, mask_hi, out + v0, ldOut, g, e, el);
    });
}

int rmg_lm_fit_cov_multi(const double *Y, const double *Z, int64_t V, int64_t ldY, int n, const double *Xs, int ps,
                         const double *mask, double mask_lo, double mask_hi, double *out, int64_t ldOut, int n_gpus,
                         char *err, size_t errlen)
{
    if (V > 0 && (!Y || !Z || !out)) return fail(err, errlen, RMG_ERR_INVALID, "Y, Z or out is NULL");
    if (ldY < V || ldOut < V) return fail(err, errlen, RMG_ERR_INVALID, "leading dimension smaller than V");
    return run_sharded(V, n_gpus, err, errlen, [&](int g, int64_t v0, int64_t cv, char *e, size_t el) {
        return rmg_lm_fit_cov(Y + v0, Z + v0, cv, ldY, n, Xs, ps, mask ? mask + v0 : nullptr, mask_lo, mask_hi, out + v0,
                              ldOut, g, e, el);
    });
}

int rmg_lm_fit_stored_multi(const void *U, int dtype, int64_t V, int64_t ldU, int n, const double *scale,
                            const double *offset, double voxel_min, int64_t slice_len, const double *X, int p,
                            const double *mask, double mask_lo, double mask_hi, double *out, int64_t ldOut, int n_gpus,
                            char *err, size_t errlen)
{
    int rc = check_stored_args(dtype, scale, offset, voxel_min, slice_len, p, err, errlen);
    if (rc) return rc;
    if ((rc = check_fit_args(U, V, ldU, n, X, p, out, ldOut, err, errlen))) return rc;
    const size_t es = dtype == RMG_STORED_U16 ? 2 : 4;
    const int64_t sl = dtype == RMG_STORED_U16 ? slice_len : std::max<int64_t>(V, 1);
    const int64_t nslices_total = (V + sl - 1) / sl;
    return run_sharded(V, n_gpus, err, errlen, [&](int g, int64_t v0, int64_t cv, char *e, size_t el) {
        StoredSpec pk{dtype, scale, offset, voxel_min, sl};
        pk.v_offset = v0;                    // slices keep their global index
        pk.nslices_total = nslices_total;
        return fit_host_impl(reinterpret_cast<const double *>(static_cast<const char *>(U) + (size_t)v0 * es), cv, ldU, n, X, p,
                             nullptr, mask ? mask + v0 : nullptr, mask_lo, mask_hi, out + v0, ldOut, g, e, el, nullptr, nullptr, &pk);
    });
}

}  // extern "C"
