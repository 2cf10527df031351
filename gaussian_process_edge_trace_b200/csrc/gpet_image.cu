// Image-space kernels: gradient stencil (comp_grad_img), float32 min-max normalise, 9x9 Gaussian
// blur (KDEpy FFTKDE restated as a direct separable convolution), gradient-image KDE, transpose.
//
// Reference seams: gpet_utils.py:95-119 (comp_grad_img), :65-91 (normalise), gpet.py:503-529
// (kernel_density_estimate, gradient branch).
#include "gpet_common.cuh"
#include <type_traits>

namespace gpet {

// ------------------------------------------------------------------------------------------------
// K1: stencil. One CTA computes a TW x TH output tile from an edge-replicated shared-memory tile.
// Accumulation order = scipy.ndimage NI_Correlate: raster order over the flipped kernel, zero taps
// skipped, separate multiply and add (no FMA) so the fp64 result is bit-identical to the reference.
// ------------------------------------------------------------------------------------------------
constexpr int ST_TW = 64, ST_TH = 32, ST_THREADS = 256;

// EXACT: fp64 tile, taps accumulated in scipy.ndimage's raster order with separate multiply and add - the fp64 map is
//        bit-identical to scipy (FP64-pipe bound: 100 D-ops per pixel for the 11 x 5 filter).
// !EXACT (opt-in, comp_grad_img(exact=False)): float32 tile and fused float32 multiply-adds - within ~1e-6 relative of
//        the exact map (north_star bar for the stencil: 1e-4), HBM bound (8 B read + 4 B written per pixel).
// Tile fill: 128-bit loads (two pixels) wherever the pair is inside the image and 16-byte aligned, clamped scalar loads on
// the replicated border.
template <bool EXACT>
__global__ void __launch_bounds__(ST_THREADS)
stencil_kernel(const double* __restrict__ img, int M, int N, const double* __restrict__ taps, int kh, int kw,
               float* __restrict__ out, uint32_t* __restrict__ minmax) {
    using T = typename std::conditional<EXACT, double, float>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* smem = reinterpret_cast<T*>(smem_raw);
    const int tw = (ST_TW + kw - 1 + 1) & ~1, th = ST_TH + kh - 1;       // even row length: pairs never straddle rows
    T* tile = smem;                 // th x tw
    T* ftap = smem + th * tw;       // kh*kw, flipped
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * ST_TW, y0 = blockIdx.y * ST_TH;
    const double* src = img + (size_t)b * M * N;
    for (int i = threadIdx.x; i < kh * kw; i += ST_THREADS) ftap[i] = (T)taps[kh * kw - 1 - i];
    const int ry = kh / 2, rx = kw / 2;
    const bool rows_aligned = ((N & 1) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    for (int i = threadIdx.x; i < th * (tw / 2); i += ST_THREADS) {
        const int ty = i / (tw / 2), tx = 2 * (i - ty * (tw / 2));
        const int gy = min(max(y0 + ty - ry, 0), M - 1);
        const int gx = x0 + tx - rx;
        double v0, v1;
        if (rows_aligned && (gx & 1) == 0 && gx >= 0 && gx + 1 < N) {
            const double2 v = __ldg(reinterpret_cast<const double2*>(src + (size_t)gy * N + gx));
            v0 = v.x;
            v1 = v.y;
        } else {
            v0 = __ldg(src + (size_t)gy * N + min(max(gx, 0), N - 1));
            v1 = __ldg(src + (size_t)gy * N + min(max(gx + 1, 0), N - 1));
        }
        tile[ty * tw + tx] = (T)v0;
        tile[ty * tw + tx + 1] = (T)v1;
    }
    __syncthreads();
    const int lx = threadIdx.x % ST_TW, ly0 = threadIdx.x / ST_TW;  // 4 row groups
    float vmin = __int_as_float(0x7f800000), vmax = 0.0f;
    // The taps of one pixel are a dependent chain by construction (scipy's order), so the instruction-level parallelism
    // comes from the thread's ST_PX pixels (rows ly0, ly0 + 4, ...): the tap loop is outermost and every pixel keeps its
    // own accumulator - same operations per pixel, in the same order.
    constexpr int ST_ROWGROUPS = ST_THREADS / ST_TW, ST_PX = ST_TH / ST_ROWGROUPS;
    T acc[ST_PX];
#pragma unroll
    for (int p = 0; p < ST_PX; ++p) acc[p] = (T)0;
    const T* base = tile + ly0 * tw + lx;
    for (int a = 0; a < kh; ++a) {
        for (int c = 0; c < kw; ++c) {
            const T t = ftap[a * kw + c];
            if (t != (T)0) {
#pragma unroll
                for (int p = 0; p < ST_PX; ++p) {
                    if constexpr (EXACT)
                        acc[p] = __dadd_rn(acc[p], __dmul_rn(base[(p * ST_ROWGROUPS + a) * tw + c], t));
                    else
                        acc[p] = fmaf(base[(p * ST_ROWGROUPS + a) * tw + c], t, acc[p]);
                }
            }
        }
    }
#pragma unroll
    for (int p = 0; p < ST_PX; ++p) {
        const int gy = y0 + ly0 + p * ST_ROWGROUPS, gx = x0 + lx;
        if (gy >= M || gx >= N) continue;
        float v;
        if constexpr (EXACT) {
            double r = acc[p];
            if (r < 0.0) r = 0.0;
            v = __double2float_rn(r);
        } else {
            v = acc[p] < 0.0f ? 0.0f : acc[p];
        }
        out[((size_t)b * M + gy) * N + gx] = v;
        vmin = fminf(vmin, v + 0.0f);
        vmax = fmaxf(vmax, v + 0.0f);
    }
    // one atomic pair per CTA
    __shared__ float smin[ST_THREADS / 32], smax[ST_THREADS / 32];
    vmin = warp_min(vmin);
    vmax = warp_max(vmax);
    if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = vmin; smax[threadIdx.x >> 5] = vmax; }
    __syncthreads();
    if (threadIdx.x < 32) {
        vmin = threadIdx.x < ST_THREADS / 32 ? smin[threadIdx.x] : __int_as_float(0x7f800000);
        vmax = threadIdx.x < ST_THREADS / 32 ? smax[threadIdx.x] : 0.0f;
        vmin = warp_min(vmin);
        vmax = warp_max(vmax);
        if (threadIdx.x == 0) atomic_minmax_nonneg(minmax + 2 * b, vmin, vmax);
    }
}

__global__ void init_minmax_kernel(uint32_t* minmax, int B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) {
        minmax[2 * i] = 0x7f800000u;  // +inf
        minmax[2 * i + 1] = 0u;       // +0
    }
}

__global__ void minmax_f32_kernel(const float* __restrict__ img, size_t per_image, uint32_t* __restrict__ minmax) {
    const int b = blockIdx.y;
    const float* p = img + (size_t)b * per_image;
    float vmin = __int_as_float(0x7f800000), vmax = 0.0f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_image; i += (size_t)gridDim.x * blockDim.x) {
        float v = p[i] + 0.0f;
        vmin = fminf(vmin, v);
        vmax = fmaxf(vmax, v);
    }
    // one atomic pair per CTA (a pair per warp made 3920 same-address atomics per image: 2.8 ms per 1250 images)
    __shared__ float smin[32], smax[32];
    vmin = warp_min(vmin);
    vmax = warp_max(vmax);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = (blockDim.x + 31) >> 5;
    if (lane == 0) { smin[warp] = vmin; smax[warp] = vmax; }
    __syncthreads();
    if (warp == 0) {
        vmin = lane < nwarps ? smin[lane] : __int_as_float(0x7f800000);
        vmax = lane < nwarps ? smax[lane] : 0.0f;
        vmin = warp_min(vmin);
        vmax = warp_max(vmax);
        if (lane == 0) atomic_minmax_nonneg(minmax + 2 * b, vmin, vmax);
    }
}

__global__ void normalise_f32_kernel(float* __restrict__ img, size_t per_image, const uint32_t* __restrict__ minmax) {
    const int b = blockIdx.y;
    float* p = img + (size_t)b * per_image;
    const float mn = __uint_as_float(minmax[2 * b]);
    const float range = __fsub_rn(__uint_as_float(minmax[2 * b + 1]), mn);
    // an image that is already min-max normalised (comp_grad_img's own output handed to GP_Edge_Tracing, gpet.py:97):
    // a - 0 and a / 1 are exact identities in float32, so neither the read nor the write is needed
    if (mn == 0.0f && range == 1.0f) return;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_image; i += (size_t)gridDim.x * blockDim.x)
        p[i] = normalise_f32(p[i], mn, range);
}

__global__ void normalise_copy_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t per_image,
                                          const uint32_t* __restrict__ minmax) {
    const int b = blockIdx.y;
    const float mn = __uint_as_float(minmax[2 * b]);
    const float range = __fsub_rn(__uint_as_float(minmax[2 * b + 1]), mn);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_image; i += (size_t)gridDim.x * blockDim.x)
        dst[(size_t)b * per_image + i] = normalise_f32(src[(size_t)b * per_image + i], mn, range);
}

// ------------------------------------------------------------------------------------------------
// 9x9 Gaussian blur (sigma = 1 px, offsets -4..4, zero outside the image), separable, fp64.
// Source element type is a template parameter: double (already weighted) or uint64 fixed point
// (the density splat, scale 2^-60).  dst = float32(scale[b] * blur) and per-image min/max.
// ------------------------------------------------------------------------------------------------
constexpr int BL_TW = 64, BL_TH = 32, BL_R = 4, BL_THREADS = 256;
__constant__ double c_gauss1d[9];  // exp(-d^2/2), d = -4..4

template <typename T>
__device__ __forceinline__ double to_f64(T v);
template <>
__device__ __forceinline__ double to_f64<double>(double v) { return v; }
template <>
__device__ __forceinline__ double to_f64<unsigned long long>(unsigned long long v) {
    return __ull2double_rn(v) * 8.6736173798840355e-19;  // 2^-60, exact scaling
}

template <typename T>
__global__ void __launch_bounds__(BL_THREADS)
blur9_kernel(const T* __restrict__ src, int M, int N, const double* __restrict__ scale, float* __restrict__ dst,
             uint32_t* __restrict__ minmax) {
    __shared__ double tile[(BL_TH + 2 * BL_R) * (BL_TW + 2 * BL_R)];
    __shared__ double vert[BL_TH * (BL_TW + 2 * BL_R)];
    constexpr int tw = BL_TW + 2 * BL_R, th = BL_TH + 2 * BL_R;
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * BL_TW, y0 = blockIdx.y * BL_TH;
    const T* s = src + (size_t)b * M * N;
    // all global loads of the tile are issued before the first one is converted and stored (the loop form waited a
    // DRAM round trip per element: 74 % of the kernel's stall samples)
    constexpr int NL = (th * tw + BL_THREADS - 1) / BL_THREADS;
    T raw[NL];
#pragma unroll
    for (int k = 0; k < NL; ++k) {
        const int i = threadIdx.x + k * BL_THREADS;
        const int ty = i / tw, tx = i - ty * tw;
        const int gy = y0 + ty - BL_R, gx = x0 + tx - BL_R;
        raw[k] = (i < th * tw && gy >= 0 && gy < M && gx >= 0 && gx < N) ? s[(size_t)gy * N + gx] : T(0);
    }
    bool nz = false;
#pragma unroll
    for (int k = 0; k < NL; ++k) {
        const int i = threadIdx.x + k * BL_THREADS;
        if (i < th * tw) tile[i] = to_f64<T>(raw[k]);
        nz |= (raw[k] != T(0));
    }
    // the curve density is non-zero only in a band around the kept curves: a tile whose halo is all zero blurs to
    // exact zeros (0 * tap sums to +0.0, as in the general path), so only the store and the min/max update remain
    if (!__syncthreads_or(nz)) {
        for (int i = threadIdx.x; i < BL_TH * BL_TW; i += BL_THREADS) {
            int ty = i / BL_TW, tx = i - ty * BL_TW;
            int gy = y0 + ty, gx = x0 + tx;
            if (gy < M && gx < N) dst[((size_t)b * M + gy) * N + gx] = 0.0f;
        }
        if (threadIdx.x == 0) atomic_minmax_nonneg(minmax + 2 * b, 0.0f, 0.0f);
        return;
    }
    for (int i = threadIdx.x; i < BL_TH * tw; i += BL_THREADS) {
        int ty = i / tw, tx = i - ty * tw;
        double a = 0.0;
#pragma unroll
        for (int d = 0; d < 9; ++d) a = fma(tile[(ty + d) * tw + tx], c_gauss1d[d], a);
        vert[i] = a;
    }
    __syncthreads();
    const double sc = scale[b];
    float vmin = __int_as_float(0x7f800000), vmax = 0.0f;
    for (int i = threadIdx.x; i < BL_TH * BL_TW; i += BL_THREADS) {
        int ty = i / BL_TW, tx = i - ty * BL_TW;
        int gy = y0 + ty, gx = x0 + tx;
        if (gy >= M || gx >= N) continue;
        double a = 0.0;
#pragma unroll
        for (int d = 0; d < 9; ++d) a = fma(vert[ty * tw + tx + d], c_gauss1d[d], a);
        const float v = __double2float_rn(a * sc);
        dst[((size_t)b * M + gy) * N + gx] = v;
        vmin = fminf(vmin, v + 0.0f);
        vmax = fmaxf(vmax, v + 0.0f);
    }
    vmin = warp_min(vmin);
    vmax = warp_max(vmax);
    if ((threadIdx.x & 31) == 0) atomic_minmax_nonneg(minmax + 2 * b, vmin, vmax);
}

template __global__ void blur9_kernel<double>(const double*, int, int, const double*, float*, uint32_t*);
template __global__ void blur9_kernel<unsigned long long>(const unsigned long long*, int, int, const double*, float*,
                                                          uint32_t*);

// __constant__ memory is per device: remember which devices of this process hold the taps
static bool g_gauss_ready[64] = {};
int ensure_gauss_taps() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && g_gauss_ready[dev]) return GPET_OK;
    double h[9];
    for (int d = -4; d <= 4; ++d) h[d + 4] = exp(-0.5 * (double)(d * d));
    cudaError_t e = cudaMemcpyToSymbol(c_gauss1d, h, sizeof(h));
    if (e != cudaSuccess) {
        set_error("cudaMemcpyToSymbol(c_gauss1d): %s", cudaGetErrorString(e));
        return GPET_ERR_CUDA;
    }
    if (dev >= 0 && dev < 64) g_gauss_ready[dev] = true;
    return GPET_OK;
}

int launch_blur9_u64(const unsigned long long* src, int B, int M, int N, const double* scale, float* dst,
                     uint32_t* minmax, cudaStream_t st) {
    int rc = ensure_gauss_taps();
    if (rc) return rc;
    dim3 grid((N + BL_TW - 1) / BL_TW, (M + BL_TH - 1) / BL_TH, B);
    blur9_kernel<unsigned long long><<<grid, BL_THREADS, 0, st>>>(src, M, N, scale, dst, minmax);
    return check_launch("blur9_kernel<u64>");
}

// ------------------------------------------------------------------------------------------------
// Gradient-image KDE: weights = G over pixels with G > 1e-3, normalised by their sum (KDEpy), then
// blur * 1/(2 pi), float32 min-max normalise.  The sum is a fixed-order two-stage reduction
// (deterministic).
// ------------------------------------------------------------------------------------------------
constexpr int GK_PARTS = 256;

__global__ void gradkde_prepare_kernel(const float* __restrict__ grad, size_t per_image, double* __restrict__ wsrc,
                                       double* __restrict__ partial) {
    const int b = blockIdx.y;
    __shared__ double red[8];
    double acc = 0.0;
    // contiguous chunk per block, strided inside: fixed order for a fixed launch shape
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_image; i += (size_t)gridDim.x * blockDim.x) {
        const double g = (double)grad[(size_t)b * per_image + i];
        const double v = (g > 1e-3) ? g : 0.0;
        wsrc[(size_t)b * per_image + i] = v;
        acc += v;
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
        partial[(size_t)b * GK_PARTS + blockIdx.x] = t;
    }
}

__global__ void gradkde_scale_kernel(const double* __restrict__ partial, double* __restrict__ scale, int B) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double t = 0.0;
    for (int i = 0; i < GK_PARTS; ++i) t += partial[(size_t)b * GK_PARTS + i];
    scale[b] = (1.0 / t) * 0.15915494309189535;  // 1/sum(weights) * 1/(2 pi)
}

// ------------------------------------------------------------------------------------------------
// Transpose [B][M][N] -> [B][N][M+2] (float32), 32x32 shared-memory tiles.  Every column carries one guard entry
// at each end that repeats its first / last value: dst[b][x][r+1] = src[b][clamp(r, 0, M-1)][x], r = -1 .. M.  The
// scoring gather then clamps only the integer row (to [-1, M-1]); both taps of a clamped point read the same value,
// so the interpolation weight needs no clamping (gpet_score.cu).
// ------------------------------------------------------------------------------------------------
__global__ void transpose_f32_kernel(const float* __restrict__ src, int M, int N, float* __restrict__ dst) {
    __shared__ float t[32][33];
    const int b = blockIdx.z;
    const float* s = src + (size_t)b * M * N;
    const int Mp = M + 2;
    float* d = dst + (size_t)b * Mp * N;
    int x = blockIdx.x * 32 + threadIdx.x;
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        int y = blockIdx.y * 32 + k;
        if (x < N && y < M) t[k][threadIdx.x] = s[(size_t)y * N + x];
    }
    __syncthreads();
    int y = blockIdx.y * 32 + threadIdx.x;
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        int xx = blockIdx.x * 32 + k;
        if (xx < N && y < M) {
            const float v = t[threadIdx.x][k];
            d[(size_t)xx * Mp + y + 1] = v;
            if (y == 0) d[(size_t)xx * Mp] = v;
            if (y == M - 1) d[(size_t)xx * Mp + M + 1] = v;
        }
    }
}

}  // namespace gpet

using namespace gpet;

static int comp_grad_img_impl(const double* img, int B, int M, int N, const double* taps, int kh, int kw, float* out,
                              uint32_t* minmax, bool exact, cudaStream_t st) {
    GPET_REQUIRE(img && taps && out && minmax, "gpet_comp_grad_img: null pointer");
    GPET_REQUIRE(B > 0 && M > 0 && N > 0 && kh > 0 && kw > 0, "gpet_comp_grad_img: bad shape");
    GPET_SUPPORTED((kh & 1) && (kw & 1) && kh <= 31 && kw <= 31, "gpet_comp_grad_img: kernel must be odd-sized <= 31x31");
    const size_t elems = (size_t)(ST_TH + kh - 1) * ((ST_TW + kw) & ~1) + (size_t)kh * kw;
    const size_t smem = elems * (exact ? sizeof(double) : sizeof(float));
    if (smem > 48 * 1024) {
        cudaError_t e = exact ? cudaFuncSetAttribute(stencil_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                              : cudaFuncSetAttribute(stencil_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("stencil smem attribute: %s", cudaGetErrorString(e));
            return GPET_ERR_CUDA;
        }
    }
    init_minmax_kernel<<<(B + 255) / 256, 256, 0, st>>>(minmax, B);
    dim3 grid((N + ST_TW - 1) / ST_TW, (M + ST_TH - 1) / ST_TH, B);
    if (exact) stencil_kernel<true><<<grid, ST_THREADS, smem, st>>>(img, M, N, taps, kh, kw, out, minmax);
    else stencil_kernel<false><<<grid, ST_THREADS, smem, st>>>(img, M, N, taps, kh, kw, out, minmax);
    int rc = check_launch("stencil_kernel");
    if (rc) return rc;
    const size_t per = (size_t)M * N;
    dim3 g2((unsigned)min((size_t)1024, (per + 1023) / 1024), B);
    normalise_f32_kernel<<<g2, 256, 0, st>>>(out, per, minmax);
    return check_launch("normalise_f32_kernel");
}

extern "C" int gpet_comp_grad_img_f64(const double* img, int B, int M, int N, const double* taps, int kh, int kw,
                                      float* out, uint32_t* minmax, void* stream) {
    return comp_grad_img_impl(img, B, M, N, taps, kh, kw, out, minmax, true, (cudaStream_t)stream);
}

extern "C" int gpet_comp_grad_img_fast_f32(const double* img, int B, int M, int N, const double* taps, int kh, int kw,
                                           float* out, uint32_t* minmax, void* stream) {
    return comp_grad_img_impl(img, B, M, N, taps, kh, kw, out, minmax, false, (cudaStream_t)stream);
}

extern "C" int gpet_normalise_f32(float* img, int B, int M, int N, uint32_t* minmax, void* stream) {
    GPET_REQUIRE(img && minmax && B > 0 && M > 0 && N > 0, "gpet_normalise_f32: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t per = (size_t)M * N;
    init_minmax_kernel<<<(B + 255) / 256, 256, 0, st>>>(minmax, B);
    dim3 g2((unsigned)min((size_t)1024, (per + 1023) / 1024), B);
    minmax_f32_kernel<<<g2, 256, 0, st>>>(img, per, minmax);
    normalise_f32_kernel<<<g2, 256, 0, st>>>(img, per, minmax);
    return check_launch("gpet_normalise_f32");
}

extern "C" int64_t gpet_grad_kde_workspace_bytes(int B, int M, int N) {
    return (int64_t)B * M * N * 8 + (int64_t)B * GK_PARTS * 8 + (int64_t)B * 8 + (int64_t)B * 8 + 256;
}

extern "C" int gpet_grad_kde_f32(const float* grad, int B, int M, int N, float* grad_kde, void* work, void* stream) {
    GPET_REQUIRE(grad && grad_kde && work && B > 0 && M > 0 && N > 0, "gpet_grad_kde_f32: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = ensure_gauss_taps();
    if (rc) return rc;
    const size_t per = (size_t)M * N;
    double* wsrc = (double*)work;
    double* partial = wsrc + (size_t)B * per;
    double* scale = partial + (size_t)B * GK_PARTS;
    uint32_t* minmax = (uint32_t*)(scale + B);
    dim3 gp(GK_PARTS, B);
    gradkde_prepare_kernel<<<gp, 256, 0, st>>>(grad, per, wsrc, partial);
    gradkde_scale_kernel<<<(B + 127) / 128, 128, 0, st>>>(partial, scale, B);
    init_minmax_kernel<<<(B + 255) / 256, 256, 0, st>>>(minmax, B);
    dim3 grid((N + BL_TW - 1) / BL_TW, (M + BL_TH - 1) / BL_TH, B);
    blur9_kernel<double><<<grid, BL_THREADS, 0, st>>>(wsrc, M, N, scale, grad_kde, minmax);
    dim3 g2((unsigned)min((size_t)1024, (per + 1023) / 1024), B);
    normalise_f32_kernel<<<g2, 256, 0, st>>>(grad_kde, per, minmax);
    return check_launch("gpet_grad_kde_f32");
}

extern "C" int gpet_transpose_f32(const float* src, int B, int M, int N, float* dst, void* stream) {
    GPET_REQUIRE(src && dst && B > 0 && M > 0 && N > 0, "gpet_transpose_f32: bad argument");
    dim3 grid((N + 31) / 32, (M + 31) / 32, B), block(32, 8);
    transpose_f32_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(src, M, N, dst);
    return check_launch("transpose_f32_kernel");
}

extern "C" int gpet_kde_normalised_f32(const float* dens, const uint32_t* minmax, int B, int M, int N, float* kde,
                                       void* stream) {
    GPET_REQUIRE(dens && minmax && kde && B > 0 && M > 0 && N > 0, "gpet_kde_normalised_f32: bad argument");
    const size_t per = (size_t)M * N;
    dim3 g2((unsigned)min((size_t)1024, (per + 1023) / 1024), B);
    normalise_copy_f32_kernel<<<g2, 256, 0, (cudaStream_t)stream>>>(dens, kde, per, minmax);
    return check_launch("normalise_copy_f32_kernel");
}
