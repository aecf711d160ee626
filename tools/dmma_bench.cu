// Micro-benchmark: FP64 mma.sync shapes on sm_100a (register-only, no memory traffic).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_bench dmma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE, int NACC>
__global__ void k(double *out, int iters)
{
    double acc[NACC][4];
    for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = threadIdx.x * 1e-9 + i;
    double a[8], b[4];
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + threadIdx.x * 1e-6 + i;
    for (int i = 0; i < 4; ++i) b[i] = 0.5 - threadIdx.x * 1e-6 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (SHAPE == 0)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(acc[i][0]), "+d"(acc[i][1]) : "d"(a[0]), "d"(b[0]));
            else if (SHAPE == 1)
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+d"(acc[i][0]), "+d"(acc[i][1]), "+d"(acc[i][2]), "+d"(acc[i][3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
            else if (SHAPE == 2)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+d"(acc[i][0]), "+d"(acc[i][1]), "+d"(acc[i][2]), "+d"(acc[i][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                             : "+d"(acc[i][0]), "+d"(acc[i][1]), "+d"(acc[i][2]), "+d"(acc[i][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                               "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
        }
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SHAPE, int NACC>
void run(const char *name, double flops_per_mma, int warps_per_sm)
{
    double *out;
    cudaMalloc(&out, 148 * 1024 * 8);
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<SHAPE, NACC><<<148, warps_per_sm * 32>>>(out, 100);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<SHAPE, NACC><<<148, warps_per_sm * 32>>>(out, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double total = (double)iters * NACC * warps_per_sm * 148 * flops_per_mma;
    printf("%-10s acc=%2d warps/SM=%2d : %8.2f TFLOP/s  (%.1f cycles/mma/SMSP at 1.965 GHz)\n", name, NACC, warps_per_sm,
           total / (ms * 1e-3) / 1e12, ms * 1e-3 * 1.965e9 / ((double)iters * NACC * warps_per_sm / 4.0));
    cudaFree(out);
}

int main()
{
    for (int w : {4, 8, 16}) {
        if (w == 4) { run<0, 16>("m8n8k4", 2.0 * 8 * 8 * 4, 4); run<1, 16>("m16n8k4", 2.0 * 16 * 8 * 4, 4); run<2, 16>("m16n8k8", 2.0 * 16 * 8 * 8, 4); run<3, 16>("m16n8k16", 2.0 * 16 * 8 * 16, 4); }
        if (w == 8) { run<0, 16>("m8n8k4", 2.0 * 8 * 8 * 4, 8); run<1, 16>("m16n8k4", 2.0 * 16 * 8 * 4, 8); run<2, 16>("m16n8k8", 2.0 * 16 * 8 * 8, 8); run<3, 16>("m16n8k16", 2.0 * 16 * 8 * 16, 8); }
        if (w == 16) { run<0, 8>("m8n8k4", 2.0 * 8 * 8 * 4, 16); run<3, 8>("m16n8k16", 2.0 * 16 * 8 * 16, 16); }
    }
    return 0;
}
