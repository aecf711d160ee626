// Micro-benchmark: scalar FP64 issue rate by OPERAND PATTERN on sm_100a, register-only.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_operand_bench dfma_operand_bench.cu
// The stored-type fit's inner loop is   acc[i][k] = fma(q[k], y[i], acc[i][k])   (+ fma(y[i], y[i], yy[i])) with 4 voxels i and
// KQ design columns k: every DFMA reads a different accumulator pair and shares q[k] or y[i] with its neighbours -- three
// 64-bit register operands, of which the operand-reuse cache can hold some.  dfma_bench.cu's  x = fma(x, a, b)  has ONE
// varying operand.  This program measures what the pipe sustains for the accumulate pattern: if it is well above 2.2 cycles
// per warp-wide DFMA, register-file bandwidth, not the kernel's schedule, sets the stored-type fit's floor.
#include <cstdio>
#include <cuda_runtime.h>

// PATTERN 0: x[i] = fma(x[i], a, b), 20 chains          (one varying operand)
// PATTERN 1: 4 x (KQ + 1) accumulators, q / y in registers, refreshed by integer ops every round (three register operands)
// PATTERN 2: the same, y[i] produced by a DFMA from w[i] (the conversion) as in the kernel
template <int PATTERN, int KQ>
__global__ void k(double *out, int iters, double a, double b, int flip)
{
    double acc[4][KQ + 1], y[4], q[KQ], x[20];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        y[i] = 1.0 + threadIdx.x * 1e-9 + i * 1e-3;
#pragma unroll
        for (int kk = 0; kk <= KQ; ++kk) acc[i][kk] = 0.0;
    }
#pragma unroll
    for (int kk = 0; kk < KQ; ++kk) q[kk] = 0.5 + kk * 1e-3;
#pragma unroll
    for (int i = 0; i < 20; ++i) x[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (PATTERN == 0) {
#pragma unroll
                for (int i = 0; i < 20; ++i) x[i] = fma(x[i], a, b);
                continue;
            }
            // "new subject": q and the samples change (integer work only, like loads + logic ops)
#pragma unroll
            for (int kk = 0; kk < KQ; ++kk) q[kk] = __hiloint2double(__double2hiint(q[kk]) ^ flip, __double2loint(q[kk]));
            double t[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                y[i] = __hiloint2double(__double2hiint(y[i]) ^ flip, __double2loint(y[i]));
                t[i] = PATTERN == 2 ? fma(a, y[i], b) : y[i];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[i][KQ] = fma(t[i], t[i], acc[i][KQ]);
#pragma unroll
                for (int kk = 0; kk < KQ; ++kk) acc[i][kk] = fma(q[kk], t[i], acc[i][kk]);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int kk = 0; kk <= KQ; ++kk) s += acc[i][kk];
#pragma unroll
    for (int i = 0; i < 20; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int PATTERN, int KQ>
void run(int warps_per_sm, const char *what)
{
    double *out;
    cudaMalloc(&out, 148 * 1024 * 8);
    const int iters = 5000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<PATTERN, KQ><<<148, warps_per_sm * 32>>>(out, 50, 0.999999, 1e-7, 0);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<PATTERN, KQ><<<148, warps_per_sm * 32>>>(out, iters, 0.999999, 1e-7, 0);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double per_round = PATTERN == 0 ? 20.0 : 4.0 * (KQ + 1) + (PATTERN == 2 ? 4.0 : 0.0);
    const double per_warp = (double)iters * 4 * per_round;
    const double cyc = ms * 1e-3 * 1.965e9;
    printf("%-58s warps/SM=%2d : %5.2f cycles per FP64 instruction and scheduler\n", what, warps_per_sm, cyc / (per_warp * warps_per_sm / 4.0));
    cudaFree(out);
}

int main()
{
    for (int w : {4, 12, 16}) {
        run<0, 4>(w, "x = fma(x, a, b), 20 chains");
        run<1, 4>(w, "acc[i][k] = fma(q[k], y[i], acc[i][k]), 4 x (4 + 1)");
        run<2, 4>(w, "... with y[i] = fma(A, w[i], C) first (the kernel's mix)");
        run<1, 7>(w, "acc[i][k] = fma(q[k], y[i], acc[i][k]), 4 x (7 + 1)");
    }
    return 0;
}
