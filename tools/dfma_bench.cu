// Micro-benchmark: scalar FP64 (DFMA) issue rate and dependent-issue latency on sm_100a, register-only.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_bench dfma_bench.cu
// Each thread runs NCH independent chains of dependent fused multiply-adds; with one warp per scheduler and NCH = 1
// the time per DFMA is the dependent-issue latency, with many chains / warps it is the pipe's issue interval.
#include <cstdio>
#include <cuda_runtime.h>

template <int NCH>
__global__ void k(double *out, int iters, double a, double b)
{
    double x[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) x[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < NCH; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NCH; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NCH>
void run(int warps_per_sm)
{
    double *out;
    cudaMalloc(&out, 148 * 1024 * 8);
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NCH><<<148, warps_per_sm * 32>>>(out, 100, 0.999999, 1e-7);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<NCH><<<148, warps_per_sm * 32>>>(out, iters, 0.999999, 1e-7);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double per_warp = (double)iters * 8 * NCH;                        // DFMAs issued by one warp
    const double cyc = ms * 1e-3 * 1.965e9;
    const double warps_per_smsp = warps_per_sm / 4.0;
    printf("chains=%2d warps/SM=%2d : %6.2f cycles per DFMA per warp, %6.2f cycles per DFMA per scheduler, %6.2f TFLOP/s\n", NCH,
           warps_per_sm, cyc / per_warp, cyc / (per_warp * (warps_per_smsp < 1 ? 1 : warps_per_smsp)),
           2.0 * per_warp * 32 * warps_per_sm * 148 / (ms * 1e-3) / 1e12);
    cudaFree(out);
}

int main()
{
    run<1>(4); run<2>(4); run<4>(4); run<8>(4); run<16>(4);
    run<1>(8); run<2>(8); run<4>(8);
    run<1>(12); run<2>(12); run<4>(12); run<8>(12);
    run<1>(16); run<4>(16);
    run<4>(32);
    return 0;
}
