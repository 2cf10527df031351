// Micro-benchmark: peak rate of mma.sync.m8n8k4.f64 (DMMA) on this GPU, register operands, 8 independent accumulators/warp.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, int iters) {
    double c[8][2];
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0; for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    double* out; cudaMalloc(&out, 148 * 8 * 256 * 8);
    for (int warps_per_sm : {4, 8, 16, 32}) {
        const int threads = 256, blocks = 148 * warps_per_sm * 32 / threads;
        const int iters = 20000;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<<<blocks, threads>>>(out, 100);
        cudaEventRecord(e0);
        k<<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 8 * 8 * 4 * 8.0 * iters * (blocks * threads / 32);
        printf("warps/SM=%2d  %.3f ms  %.1f TFLOP/s\n", warps_per_sm, ms, flops / ms / 1e9);
    }
    return 0;
}
