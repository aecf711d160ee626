// capi_common.hpp -- helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
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

// Runs fn(begin, end) over [0, count) on up to hardware_concurrency host threads (host-side
// preparation of the permutation test: factorisations and operand packing).
// Host worker threads for this process: the cores are shared with the other ranks of the box (one
// process per GPU), so take hardware_concurrency / visible GPUs, at most 16 (RMG_HOST_THREADS overrides).
int host_threads();

template <class Fn>
inline void parallel_for(int count, Fn fn)
{
    int nt = host_threads();
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

// Pageable host memory (R's heap, plain numpy) reaches only ~9.5 GB/s through cudaMemcpy2DAsync
// (the driver stages it single-threaded); pinned memory reaches the PCIe rate (55 GB/s).  For
// pageable Y the rows of a chunk are therefore copied by all host threads into a small ring of
// pinned staging buffers, and the DMA engine drains the ring while the next buffer is being filled.
struct PinnedRing {
    static constexpr int kSlots = 4;
    char *buf[kSlots] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t done[kSlots] = {nullptr, nullptr, nullptr, nullptr};
    bool busy[kSlots] = {false, false, false, false};
    size_t slot_bytes = 0;
    int next = 0;
    cudaError_t init(size_t bytes)
    {
        slot_bytes = bytes;
        for (int s = 0; s < kSlots; ++s) {
            cudaError_t e = cudaHostAlloc(reinterpret_cast<void **>(&buf[s]), bytes, cudaHostAllocDefault);
            if (e != cudaSuccess) return e;
            e = cudaEventCreateWithFlags(&done[s], cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    // next free slot (waits for the DMA that last read it)
    cudaError_t acquire(int &slot)
    {
        slot = next;
        next = (next + 1) % kSlots;
        if (busy[slot]) {
            cudaError_t e = cudaEventSynchronize(done[slot]);
            if (e != cudaSuccess) return e;
            busy[slot] = false;
        }
        return cudaSuccess;
    }
    void release_all()
    {
        for (int s = 0; s < kSlots; ++s) {
            if (done[s]) cudaEventDestroy(done[s]);
            if (buf[s]) cudaFreeHost(buf[s]);
            buf[s] = nullptr;
            done[s] = nullptr;
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

// n rows of row_bytes bytes (source pitch spitch) -> device rows of pitch dpitch, queued on `up`.  ring == nullptr:
// one strided copy (pinned or small sources); else all host threads copy groups of rows into the next free pinned
// slot and the DMA engine drains the slots behind them.
inline cudaError_t upload_rows(PinnedRing *ring, int rows_per_slot, char *dst, size_t dpitch, const char *src, size_t spitch,
                               size_t row_bytes, int n, cudaStream_t up)
{
    if (row_bytes == 0 || n <= 0) return cudaSuccess;
    if (!ring) return cudaMemcpy2DAsync(dst, dpitch, src, spitch, row_bytes, (size_t)n, cudaMemcpyHostToDevice, up);
    for (int r0 = 0; r0 < n; r0 += rows_per_slot) {
        const int rows = std::min(rows_per_slot, n - r0);
        int slot = 0;
        cudaError_t e = ring->acquire(slot);
        if (e != cudaSuccess) return e;
        char *buf = ring->buf[slot];
        const int pieces = rows * 8;                   // 8 column pieces per row keep every thread busy
        parallel_for(pieces, [&](int i0, int i1) {
            for (int i = i0; i < i1; ++i) {
                const int r = i / 8, part = i % 8;
                const size_t a0 = row_bytes * part / 8, a1 = row_bytes * (part + 1) / 8;
                std::memcpy(buf + (size_t)r * row_bytes + a0, src + (size_t)(r0 + r) * spitch + a0, a1 - a0);
            }
        });
        e = cudaMemcpy2DAsync(dst + (size_t)r0 * dpitch, dpitch, buf, row_bytes, row_bytes, (size_t)rows, cudaMemcpyHostToDevice, up);
        if (e != cudaSuccess) return e;
        e = cudaEventRecord(ring->done[slot], up);
        if (e != cudaSuccess) return e;
        ring->busy[slot] = true;
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

// Factor X and translate the reference's error contract.  0 or RMG_ERR_*.
int factor_or_fail(const double *X, int n, int p, Design &D, char *err, size_t errlen, const double *w = nullptr);

}  // namespace rmg
