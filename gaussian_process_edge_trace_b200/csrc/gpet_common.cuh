// Shared helpers for the libgpet_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "gpet_b200.h"

namespace gpet {

void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return GPET_ERR_CUDA;
    }
    return GPET_OK;
}

#define GPET_REQUIRE(cond, ...)            \
    do {                                   \
        if (!(cond)) {                     \
            gpet::set_error(__VA_ARGS__);  \
            return GPET_ERR_INVALID;       \
        }                                  \
    } while (0)

#define GPET_SUPPORTED(cond, ...)          \
    do {                                   \
        if (!(cond)) {                     \
            gpet::set_error(__VA_ARGS__);  \
            return GPET_ERR_UNSUPPORTED;   \
        }                                  \
    } while (0)

// gpet_sym_eig_f64 with the solver given explicitly (gpet_factor.cu): jt == 0 Householder + QL, else the parallel cyclic
// Jacobi kernel with jt threads per matrix
int sym_eig_run(double* Mr, int B, int rp, double* d, double* Q, int32_t* sweeps, void* work, void* stream, int jt);

constexpr int kNumSMs = 148;  // B200
extern int g_tune[GPET_TUNE_COUNT];

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// float32 min-max normalisation exactly as numpy does it in gpet_utils.normalise: both steps are
// correctly-rounded float32 operations (no FMA, no reciprocal).
__device__ __forceinline__ float normalise_f32(float a, float mn, float range) {
    return __fdiv_rn(__fsub_rn(a, mn), range);
}

// minmax scratch layout: [2*b] = bits of min (float >= 0 or +inf), [2*b+1] = bits of max.
// All normalised maps here are >= 0 (clipped gradient, densities), so the int ordering of the bit
// patterns equals the float ordering.
__device__ __forceinline__ void atomic_minmax_nonneg(uint32_t* mm, float vmin, float vmax) {
    atomicMin(mm, __float_as_uint(vmin));
    atomicMax(mm + 1, __float_as_uint(vmax));
}

}  // namespace gpet
