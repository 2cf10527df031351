// Micro-benchmark: does a warp-wide FP64 instruction on sm_100 occupy the issue slot for 1 or 2 cycles?
// Loop body = ND independent DFMAs + NA independent integer ops per iteration; time vs NA tells.
#include <cstdio>
#include <cuda_runtime.h>
template <int ND, int NA>
__global__ void k(double* out, int iters, double a, int ia) {
    double d[8];
    int x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { d[i] = threadIdx.x * 1e-3 + i; x[i] = threadIdx.x + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ND; ++i) d[i % 8] = fma(d[i % 8], a, 1e-9);
#pragma unroll
        for (int i = 0; i < NA; ++i) x[i % 8] = (x[i % 8] ^ ia) + (x[(i + 1) % 8] >> 1);
    }
    double s = 0; int xs = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += d[i]; xs += x[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + xs;
}
template <int ND, int NA>
void run(double* out) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000, blocks = 148 * 8, threads = 256;   // 64 warps/SM
    k<ND, NA><<<blocks, threads>>>(out, 10, 1.0000001, 3);
    cudaEventRecord(e0);
    k<ND, NA><<<blocks, threads>>>(out, iters, 1.0000001, 3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // cycles per iteration per SM sub-partition: 16 warps per SMSP
    double cyc = ms * 1e-3 * 1.965e9 / iters / 16.0;
    printf("ND=%2d NA=%3d  %.3f ms  %.1f cycles/iter/warp-slot (ND*2=%d, ND+2NA~=%d)\n", ND, NA, ms, cyc, ND * 2, ND + 2 * NA);
}
int main() {
    double* out; cudaMalloc(&out, 148 * 8 * 256 * 8);
    run<40, 0>(out); run<40, 10>(out); run<40, 20>(out); run<40, 40>(out); run<40, 60>(out); run<0, 40>(out); run<0, 80>(out);
    run<20, 40>(out);
    return 0;
}
