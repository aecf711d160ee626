// Micro-benchmark: DFMA issue rate when integer / shared-memory instructions are interleaved (sm_100a, register-only
// except for the LDS variant).  MIX = number of extra instructions per 7 DFMAs; KIND 0 = LOP3 (ALU pipe), 1 = IMAD (FMA
// pipe), 2 = LDS.64.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_mix_bench dfma_mix_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MIX, int KIND>
__global__ void k(double *out, int iters, double a, double b, unsigned m)
{
    __shared__ double sh[512];
    sh[threadIdx.x % 512] = threadIdx.x;
    __syncthreads();
    double x[8];
    unsigned u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-9 + i; u[i] = threadIdx.x + i; }
    double acc = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 7; ++i) x[i] = fma(x[i], a, b);
#pragma unroll
            for (int i = 0; i < MIX; ++i) {
                if (KIND == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(m), "r"(u[(i + 1) & 7]));
                else if (KIND == 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(m), "r"(u[(i + 1) & 7]));
                else {
                    double t;
                    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(t) : "r"((unsigned)((u[i] & 511u) * 8u)));
                    acc += 0;   // keep t alive through the asm volatile only
                    u[i] += (unsigned)__double2hiint(t) & 1u;
                }
            }
        }
    }
    double s = acc;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i] + u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MIX, int KIND>
void run(int warps_per_sm)
{
    double *out;
    cudaMalloc(&out, 148 * 1024 * 8);
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MIX, KIND><<<148, warps_per_sm * 32>>>(out, 100, 0.999999, 1e-7, 0x9e3779b9u);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MIX, KIND><<<148, warps_per_sm * 32>>>(out, iters, 0.999999, 1e-7, 0x9e3779b9u);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double dfma_per_sched = (double)iters * 8 * 7 * (warps_per_sm / 4.0);
    printf("kind=%d mix=%d per 7 DFMA, warps/SM=%2d : %5.2f cycles per DFMA per scheduler\n", KIND, MIX, warps_per_sm,
           ms * 1e-3 * 1.965e9 / dfma_per_sched);
    cudaFree(out);
}

int main()
{
    run<0, 0>(12);
    run<2, 0>(12); run<4, 0>(12); run<7, 0>(12);
    run<2, 1>(12); run<4, 1>(12); run<7, 1>(12);
    run<1, 2>(12); run<2, 2>(12);
    run<4, 0>(16); run<4, 1>(16);
    return 0;
}
