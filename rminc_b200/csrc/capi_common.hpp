// capi_common.hpp -- helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rminc_b200.h"
#include "design.hpp"
#include "lm_common.cuh"
#include "lm_launch.hpp"

namespace rmg {

extern std::atomic<int64_t> g_launch_count;
extern thread_local double g_last_kernel_ms;
extern thread_local int g_last_fit_variant;

// NVTX ranges for a profiler timeline: one range for the call, one nested range per phase.  Header-only
// NVTX 3: without an attached tool every call is a no-op through a null function pointer.
struct NvtxPhase {
    int depth = 0;
    explicit NvtxPhase(const char *call)
    {
        nvtxRangePushA(call);
        depth = 1;
    }
    void next(const char *phase)
    {
        if (depth == 2) nvtxRangePop();
        nvtxRangePushA(phase);
        depth = 2;
    }
    ~NvtxPhase()
    {
        while (depth-- > 0) nvtxRangePop();
    }
    NvtxPhase(const NvtxPhase &) = delete;
    NvtxPhase &operator=(const NvtxPhase &) = delete;
};

inline int fail(char *err, size_t errlen, int code, const char *fmt, ...)
{
    if (err && errlen) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(err, errlen, fmt, ap);
        va_end(ap);
    }
    return code;
}

#define RMG_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            rc = fail(err, errlen, RMG_ERR_CUDA, "%s failed: %s (%s:%d)", #call,                \
                      cudaGetErrorString(e__), __FILE__, __LINE__);                             \
            goto cleanup;                                                                       \
        }                                                                                       \
    } while (0)

// Validates the device ordinal and makes it current.  0 on success.
int select_device(int device, char *err, size_t errlen);
int device_sm_count(int device);

// Common argument validation of the fit entry points.
int check_fit_args(const void *Y, int64_t V, int64_t ldY, int n, const void *X, int p, const void *out,
                   int64_t ldOut, char *err, size_t errlen);

// Host images of the per-design device constants.
void fill_dev_design(const Design &D, DevDesign &dd);
// Row-major [n][KP] copy of Q (columns >= rank zero).
std::vector<double> make_qt(const Design &D, int KP);

// Process-wide pool of pinned (page-locked, portable) host buffers.  cudaHostAlloc costs ~0.3 ms per MiB, so the staging
// ring of a pageable upload (4 x 128 MiB) and the packed permutation operand are leased from here and kept across
// calls; rmg_release_host_buffers() returns every idle buffer to the OS.  Buffers are allocated with the pages
// interleaved over the NUMA nodes the process may use, so that the DMA engines of GPUs on either socket read them at
// the same rate.
void *pinned_acquire(size_t bytes);       // nullptr when the allocation fails
void pinned_release(void *ptr);
void pinned_trim();

// Host worker threads available to THIS caller.  One process per GPU (torchrun): the cores are shared with the other
// ranks of the box (LOCAL_WORLD_SIZE); a single process driving several GPUs owns them all.  RMG_HOST_THREADS overrides.
int host_threads();
// > 1 while this thread is one of several concurrent per-GPU workers of a call: parallel_for takes its share of the cores
extern thread_local int tl_host_share;

// Runs fn(begin, end) over [0, count) on up to host_threads() threads (host-side preparation of the
// permutation test: factorisations and operand packing; staging copies of pageable uploads).

template <class Fn>
inline void parallel_for(int count, Fn fn)
{
    int nt = std::max(1, host_threads() / std::max(1, tl_host_share));
    if (nt > count / 8) nt = count / 8;
    if (nt <= 1) {
        fn(0, count);
        return;
    }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) {
        const int a = (int)((int64_t)count * t / nt), b = (int)((int64_t)count * (t + 1) / nt);
        th.emplace_back([=, &fn]() { fn(a, b); });
    }
    for (auto &t : th) t.join();
}

// Copy into a pinned staging slot with non-temporal stores: the slot is written once and read by the DMA engine, never by
// this core, so write-allocating it (memcpy's regular stores below its non-temporal threshold) costs a third more memory
// traffic and evicts the source lines still to be read.  RMG_STAGING_MEMCPY=1 selects plain memcpy (the twin).
inline void staging_copy(void *dst, const void *src, size_t bytes)
{
#if defined(__SSE2__)
    static const bool plain = [] { const char *e = std::getenv("RMG_STAGING_MEMCPY"); return e && std::atoi(e) != 0; }();
    if (plain || bytes < 4096) { std::memcpy(dst, src, bytes); return; }
    char *d = static_cast<char *>(dst);
    const char *s_ = static_cast<const char *>(src);
    const size_t head = (16 - (reinterpret_cast<uintptr_t>(d) & 15)) & 15;
    std::memcpy(d, s_, head);
    d += head; s_ += head; bytes -= head;
    size_t blocks = bytes / 64;
    while (blocks--) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s_));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s_ + 16));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s_ + 32));
        const __m128i e = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s_ + 48));
        _mm_stream_si128(reinterpret_cast<__m128i *>(d), a);
        _mm_stream_si128(reinterpret_cast<__m128i *>(d + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i *>(d + 32), c);
        _mm_stream_si128(reinterpret_cast<__m128i *>(d + 48), e);
        s_ += 64; d += 64;
    }
    _mm_sfence();
    std::memcpy(d, s_, bytes % 64);
#else
    std::memcpy(dst, src, bytes);
#endif
}

// Pageable host memory (R's heap, plain numpy) reaches only ~9.5 GB/s through cudaMemcpy2DAsync
// (the driver stages it single-threaded); pinned memory reaches the PCIe rate (55 GB/s).  For
// pageable Y the rows of a chunk are therefore copied by all host threads into a small ring of
// pinned staging buffers, and the DMA engines drain the ring while the next buffer is being filled.
// One ring can feed several devices: a slot remembers the device whose stream last read it.
struct PinnedRing {
    static constexpr int kSlots = 4;
    static constexpr int kMaxDev = 16;
    char *buf[kSlots] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t done[kSlots][kMaxDev] = {};
    int busy_dev[kSlots] = {-1, -1, -1, -1};
    size_t slot_bytes = 0;
    int next = 0;
    cudaError_t init(size_t bytes)
    {
        slot_bytes = bytes;
        for (int s = 0; s < kSlots; ++s) {
            buf[s] = static_cast<char *>(pinned_acquire(bytes));
            if (!buf[s]) return cudaErrorMemoryAllocation;
        }
        return cudaSuccess;
    }
    // next free slot (waits for the DMA that last read it)
    cudaError_t acquire(int &slot)
    {
        slot = next;
        next = (next + 1) % kSlots;
        if (busy_dev[slot] >= 0) {
            cudaError_t e = cudaEventSynchronize(done[slot][busy_dev[slot]]);
            if (e != cudaSuccess) return e;
            busy_dev[slot] = -1;
        }
        return cudaSuccess;
    }
    // the copy out of `slot` has been queued on `st`, a stream of `device` (which is current)
    cudaError_t mark(int slot, int device, cudaStream_t st)
    {
        if (device < 0 || device >= kMaxDev) return cudaErrorInvalidDevice;
        if (!done[slot][device]) {
            cudaError_t e = cudaEventCreateWithFlags(&done[slot][device], cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
        }
        cudaError_t e = cudaEventRecord(done[slot][device], st);
        if (e == cudaSuccess) busy_dev[slot] = device;
        return e;
    }
    // waits for every queued copy, then hands the slots back to the pool
    void release_all()
    {
        for (int s = 0; s < kSlots; ++s) {
            if (busy_dev[s] >= 0) cudaEventSynchronize(done[s][busy_dev[s]]);
            busy_dev[s] = -1;
            for (int d = 0; d < kMaxDev; ++d)
                if (done[s][d]) {
                    cudaEventDestroy(done[s][d]);
                    done[s][d] = nullptr;
                }
            if (buf[s]) pinned_release(buf[s]);
            buf[s] = nullptr;
        }
    }
};

inline bool is_pinned_host(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

// Should a host matrix of total_bytes travel through the ring, and with how many rows (subjects) per ~128 MiB slot?
// RMG_STAGING_ROWS=k forces the ring with k rows per slot at any size (tests); RMG_NO_STAGING disables it.
inline bool staging_plan(const void *src, double total_bytes, size_t row_bytes, int n, int &rows_per_slot)
{
    const int forced_rows = std::getenv("RMG_STAGING_ROWS") ? std::atoi(std::getenv("RMG_STAGING_ROWS")) : 0;
    rows_per_slot = forced_rows > 0 ? std::min(forced_rows, n)
        : (int)std::max<int64_t>(1, std::min<int64_t>(n, (128ll << 20) / (int64_t)std::max<size_t>(row_bytes, 1)));
    return total_bytes > 0 && !is_pinned_host(src) && !std::getenv("RMG_NO_STAGING") &&
           (forced_rows > 0 || total_bytes >= 256.0 * 1024 * 1024);
}

// n rows of row_bytes bytes (source pitch spitch) -> device rows of pitch dpitch, queued on `up`, a stream of `device`
// (made current).  ring == nullptr: one strided copy (pinned or small sources); else all host threads copy groups of
// rows into the next free pinned slot and the DMA engine drains the slots behind them.
inline cudaError_t upload_rows(PinnedRing *ring, int rows_per_slot, char *dst, size_t dpitch, const char *src, size_t spitch,
                               size_t row_bytes, int n, cudaStream_t up, int device)
{
    if (row_bytes == 0 || n <= 0) return cudaSuccess;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return e;
    if (!ring) return cudaMemcpy2DAsync(dst, dpitch, src, spitch, row_bytes, (size_t)n, cudaMemcpyHostToDevice, up);
    rows_per_slot = (int)std::max<size_t>(1, std::min<size_t>((size_t)rows_per_slot, ring->slot_bytes / row_bytes));
    for (int r0 = 0; r0 < n; r0 += rows_per_slot) {
        const int rows = std::min(rows_per_slot, n - r0);
        int slot = 0;
        e = ring->acquire(slot);
        if (e != cudaSuccess) return e;
        char *buf = ring->buf[slot];
        const int pieces = rows * 8;                   // 8 column pieces per row keep every thread busy
        parallel_for(pieces, [&](int i0, int i1) {
            for (int i = i0; i < i1; ++i) {
                const int r = i / 8, part = i % 8;
                const size_t a0 = row_bytes * part / 8, a1 = row_bytes * (part + 1) / 8;
                staging_copy(buf + (size_t)r * row_bytes + a0, src + (size_t)(r0 + r) * spitch + a0, a1 - a0);
            }
        });
        e = cudaMemcpy2DAsync(dst + (size_t)r0 * dpitch, dpitch, buf, row_bytes, row_bytes, (size_t)rows, cudaMemcpyHostToDevice, up);
        if (e != cudaSuccess) return e;
        e = ring->mark(slot, device, up);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// Volumes in their stored type (rmg_lm_fit_stored): Y is then V x n samples of `dtype`, not doubles.
struct StoredSpec {
    int dtype;
    const double *scale, *offset;   // host, [nslices x n], slice fastest (u16 only)
    double voxel_min;
    int64_t slice_len;
    int64_t v_offset = 0;        // global index of this call's first voxel (a shard of a larger volume)
    int64_t nslices_total = 0;   // rows of scale / offset (0: derive from this call's V)
};

inline int check_stored_args(int dtype, const double *scale, const double *offset, double voxel_min, int64_t slice_len,
                             int p, char *err, size_t errlen)
{
    if (dtype != RMG_STORED_U16 && dtype != RMG_STORED_F32)
        return fail(err, errlen, RMG_ERR_INVALID, "dtype must be RMG_STORED_U16 or RMG_STORED_F32 (got %d)", dtype);
    if (p > 16) return fail(err, errlen, RMG_ERR_INVALID, "stored-type fits support at most 16 design columns (got %d)", p);
    if (dtype == RMG_STORED_U16) {
        if (!scale || !offset) return fail(err, errlen, RMG_ERR_INVALID, "scale / offset are NULL");
        if (slice_len < 1) return fail(err, errlen, RMG_ERR_INVALID, "slice_len must be >= 1");
        if (!(std::fabs(voxel_min) <= 2147483648.0) || voxel_min != std::floor(voxel_min))
            return fail(err, errlen, RMG_ERR_INVALID, "voxel_min must be an integer (got %g)", voxel_min);
    }
    return RMG_OK;
}

// Device-resident design constants, cached per (device, X).  A mincLm call, every chunk of
// the host pipeline and repeated calls with the same model matrix share one entry, so the
// steady-state cost of a call is the kernel launch alone: no re-factorisation, no upload and
// no allocation on the caller's stream.  Entries are immutable once `ready` has fired; consumers
// make their stream wait on it.
struct CachedDesign {
    int device = -1;
    std::vector<double> X;   // n x p, the key ...
    std::vector<int32_t> asgn;   // ... together with the term assignment (empty = lm epilogue)
    std::vector<double> w;       // ... and the observation weights (empty = unweighted)
    Design D;
    DevDesign *dd = nullptr;
    double *qt = nullptr;
    double *rowscale = nullptr;  // device copy of sqrt(w) (weighted fits)
    cudaEvent_t ready = nullptr;
    uint64_t last_use = 0;
    ~CachedDesign()
    {
        // cudaFree synchronises the device: no kernel can still be reading the buffers
        int cur = 0;
        if (device >= 0 && cudaGetDevice(&cur) == cudaSuccess) {
            cudaSetDevice(device);
            if (dd) cudaFree(dd);
            if (qt) cudaFree(qt);
            if (rowscale) cudaFree(rowscale);
            if (ready) cudaEventDestroy(ready);
            cudaSetDevice(cur);
        }
    }
};

std::shared_ptr<CachedDesign> design_cache_get(int device, const double *X, int n, int p, const int32_t *asgn,
                                               int &rc, char *err, size_t errlen, const double *w = nullptr);
LmPlan plan_from_env(int device);

// Factor X and translate the reference's error contract.  0 or RMG_ERR_*.
int factor_or_fail(const double *X, int n, int p, Design &D, char *err, size_t errlen, const double *w = nullptr);

}  // namespace rmg
