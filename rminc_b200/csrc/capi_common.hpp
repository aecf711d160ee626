// capi_common.hpp -- helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
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

// Factor X and translate the reference's error contract.  0 or RMG_ERR_*.
int factor_or_fail(const double *X, int n, int p, Design &D, char *err, size_t errlen, const double *w = nullptr);

}  // namespace rmg
