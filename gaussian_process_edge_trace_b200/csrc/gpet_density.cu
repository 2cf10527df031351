// Kernel density of the kept posterior curves and per-bin candidate selection.
//
// Reference seams: gpet.py:485-500, 514-527 (kernel_density_estimate of best curves: KDEpy.FFTKDE = weights
// normalised to 1, linear binning on the integer lattice, 9x9 Gaussian, float32 min-max normalise);
// gpet.py:622-662 and :532-618 (get_best_pixels / compute_new_obs) in the collapsed per-bin form of
// SURVEY.md A.3.
//
// Determinism: the splat accumulates in 64-bit fixed point (scale 2^60), so the result does not depend on the
// order in which atomics land (and is more accurate than the fp64 running sum it replaces).
#include "gpet_common.cuh"

namespace gpet {

int launch_blur9_u64(const unsigned long long* src, int B, int M, int N, const double* scale, float* dst,
                     uint32_t* minmax, cudaStream_t st);
__global__ void init_minmax_kernel(uint32_t* minmax, int B);

constexpr double FX_SCALE = 1152921504606846976.0;  // 2^60

__global__ void __launch_bounds__(256)
density_splat_kernel(const double* __restrict__ Y, const int32_t* __restrict__ idx, const double* __restrict__ wts, int n,
                     int S, int Kp, int M, int N, int x_st, unsigned long long* __restrict__ grid,
                     int32_t* __restrict__ n_out) {
    const int b = blockIdx.y;
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < (long long)Kp * n) {
        const int j = (int)(p / Kp), c = (int)(p - (long long)j * Kp);  // point order of the reference: column, then curve
        const int s = idx[(size_t)b * Kp + c];
        if (s < 0) return;                   // sample-sharded runs: this kept curve lives on another rank
        const double y = Y[((size_t)b * n + j) * S + s];
        if (!(y < 0.0 || y > (double)(M - 1))) {  // gpet.py:498-500
            const double w = wts[(size_t)b * Kp + c];
            // KDEpy linear binning on the lattice y in [-1 .. M]: t = (y - (-1)) / 1
            const double ty = y + 1.0;
            const double iy = floor(ty);
            const double fy = ty - iy;
            const int r0 = (int)iy - 1;  // image row of the lower tap
            unsigned long long* cell = grid + ((size_t)b * M + r0) * N + (x_st + j);
            atomicAdd(cell, __double2ull_rn(((1.0 - fy) * w) * FX_SCALE));
            if (fy > 0.0 && r0 + 1 < M) atomicAdd(cell + N, __double2ull_rn((fy * w) * FX_SCALE));
        } else {
            atomicAdd(n_out + (size_t)b * Kp + c, 1);  // dropped point (rare): exact integer count per curve
        }
    }
}

// KDEpy renormalises the point weights by their sum over the in-domain points:
// total = sum_c w_c * (n - dropped_c), accumulated in a fixed order by one thread per trace.
__global__ void density_scale_kernel(const double* __restrict__ wts, const int32_t* __restrict__ n_out, int n, int Kp,
                                     double* __restrict__ scale, int B) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double t = 0.0;
    for (int c = 0; c < Kp; ++c) t += wts[(size_t)b * Kp + c] * (double)(n - n_out[(size_t)b * Kp + c]);
    scale[b] = (1.0 / t) * 0.15915494309189535;  // / sum(weights) / (2 pi)
}

// ---- selection ------------------------------------------------------------------------------------
// col_bin[x] >= 0: bin of a candidate column; < 0: -(bin+1) of a column whose new pixels are excluded
// (fix_endpoints).  group_cols[g] .. group_cols[g+1]-1 = columns of CTA g; no bin straddles two groups.
constexpr int SEL_THREADS = 256, SEL_MAX_W = 64;

__device__ __forceinline__ double pixel_score(double kde, double gk) {
    // gpet.py:582  1/3 * (iv*gv + iv + gv), numpy evaluation order, no FMA contraction
    return __dmul_rn(1.0 / 3.0, __dadd_rn(__dadd_rn(__dmul_rn(kde, gk), kde), gk));
}

__global__ void __launch_bounds__(SEL_THREADS)
select_kernel(const float* __restrict__ dens, const uint32_t* __restrict__ minmax, const float* __restrict__ grad_kde,
              const int32_t* __restrict__ img_index, const int32_t* __restrict__ bands,
              int M, int N, const int32_t* __restrict__ col_bin, const int32_t* __restrict__ group_cols,
              const int32_t* __restrict__ old_yx, const int32_t* __restrict__ n_old, int max_old, int nb,
              double* __restrict__ bin_score, int32_t* __restrict__ bin_pos) {
    __shared__ unsigned long long best_s[SEL_MAX_W];
    __shared__ unsigned int best_p[SEL_MAX_W];
    const int b = blockIdx.y, g = blockIdx.x, tid = threadIdx.x;
    const int c0 = group_cols[g], c1 = group_cols[g + 1];
    const int W = c1 - c0;
    if (W <= 0) return;
    // band-limited densities (gpet_density_bands_f64): only rows [r_lo, r_hi) of this group's columns are stored, every
    // other pixel of these columns is exactly zero
    const int r_lo = bands ? bands[((size_t)b * gridDim.x + g) * 2] : 0;
    const int r_hi = bands ? bands[((size_t)b * gridDim.x + g) * 2 + 1] : M;
    int cb0 = col_bin[c0];
    const int bin0 = cb0 >= 0 ? cb0 : -(cb0 + 1);
    if (tid < SEL_MAX_W) {
        best_s[tid] = 0ull;
        best_p[tid] = 0xffffffffu;
    }
    __syncthreads();
    const float mn = __uint_as_float(minmax[2 * b]);
    const float range = __fsub_rn(__uint_as_float(minmax[2 * b + 1]), mn);
    const float* db = dens + (size_t)b * M * N;
    const float* gb = grad_kde + (size_t)(img_index ? img_index[b] : b) * M * N;
    const int rows_per_pass = SEL_THREADS / W;
    double my_s = -1.0;
    unsigned int my_p = 0xffffffffu;
    int my_bin = -1;
    if (tid < rows_per_pass * W) {
        const int x = c0 + tid % W;
        const int cb = col_bin[x];
        if (cb >= 0) {
            my_bin = cb - bin0;
            // four rows per trip, their density loads issued together (one dependent DRAM round trip per row made the
            // kernel latency bound).  Most pixels have no density at all: a pixel whose un-normalised excess over the
            // minimum is below half the threshold cannot pass kde > 1e-3, so the exact float32 division is skipped.
            const float skip_below = 0.5e-3f * range;
            for (int y = r_lo + tid / W; y < r_hi; y += 4 * rows_per_pass) {
                float v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int yy = y + k * rows_per_pass;
                    v[k] = (yy < r_hi) ? db[(size_t)yy * N + x] : mn;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int yy = y + k * rows_per_pass;
                    if (!(__fsub_rn(v[k], mn) > skip_below)) continue;
                    const double kde = (double)normalise_f32(v[k], mn, range);
                    if (kde > 1e-3) {
                        const double s = pixel_score(kde, (double)gb[(size_t)yy * N + x]);
                        if (s > my_s) {  // rows ascend: the first maximum is kept
                            my_s = s;
                            my_p = (unsigned int)(max_old + yy * N + x);
                        }
                    }
                }
            }
        }
    }
    // previously accepted observations (gpet.py:568-574): rescored, not subject to the column filter; any number of
    // them (thread tid takes observations tid, tid + 256, ...), scored again in the position pass
    const int n_old_b = n_old[b];
    auto old_score = [&](int o, int& bin) -> double {
        const int y = old_yx[((size_t)b * max_old + o) * 2], x = old_yx[((size_t)b * max_old + o) * 2 + 1];
        bin = -1;
        if (x < c0 || x >= c1 || y < r_lo || y >= r_hi) return -1.0;   // outside the band: density 0, kde <= 0
        const double kde = (double)normalise_f32(db[(size_t)y * N + x], mn, range);
        if (!(kde > 1e-3)) return -1.0;
        const int cb = col_bin[x];
        bin = (cb >= 0 ? cb : -(cb + 1)) - bin0;
        return pixel_score(kde, (double)gb[(size_t)y * N + x]);
    };
    // scores are >= 0, so the bit pattern orders like the value
    if (my_bin >= 0 && my_s >= 0.0) atomicMax(&best_s[my_bin], (unsigned long long)__double_as_longlong(my_s) + 1ull);
    for (int o = tid; o < n_old_b; o += SEL_THREADS) {
        int ob;
        const double os = old_score(o, ob);
        if (ob >= 0) atomicMax(&best_s[ob], (unsigned long long)__double_as_longlong(os) + 1ull);
    }
    __syncthreads();
    if (my_bin >= 0 && my_s >= 0.0 && best_s[my_bin] == (unsigned long long)__double_as_longlong(my_s) + 1ull)
        atomicMin(&best_p[my_bin], my_p);
    for (int o = tid; o < n_old_b; o += SEL_THREADS) {
        int ob;
        const double os = old_score(o, ob);
        if (ob >= 0 && best_s[ob] == (unsigned long long)__double_as_longlong(os) + 1ull)
            atomicMin(&best_p[ob], (unsigned int)o);
    }
    __syncthreads();
    int cb1 = col_bin[c1 - 1];
    const int bin1 = cb1 >= 0 ? cb1 : -(cb1 + 1);
    for (int k = tid; k <= bin1 - bin0; k += SEL_THREADS) {
        const int bin = bin0 + k;
        if (bin < nb) {
            const unsigned long long v = best_s[k];
            bin_score[(size_t)b * nb + bin] = v ? __longlong_as_double((long long)(v - 1ull)) : -1.0;
            bin_pos[(size_t)b * nb + bin] = v ? (int32_t)best_p[k] : -1;
        }
    }
}

}  // namespace gpet

using namespace gpet;

extern "C" int64_t gpet_density_workspace_bytes(int B, int M, int N, int Kp) {
    return (int64_t)B * M * N * 8 + (int64_t)B * 8 + (int64_t)B * Kp * 4 + 256;
}

// Workspace layout: u64 grid[B][M][N] | f64 scale[B] | i32 n_out[B][Kp].  The first two entry points are the two
// halves of gpet_density_f64; a sample-sharded run all-reduces grid and n_out (exact integer sums) between them.
extern "C" int gpet_density_splat_f64(const double* Y, const int32_t* idx, const double* wts, int B, int n, int S, int Kp,
                                      int M, int N, int x_st, void* work, void* stream) {
    GPET_REQUIRE(Y && idx && wts && work, "gpet_density_splat_f64: null pointer");
    GPET_REQUIRE(B > 0 && n > 0 && S > 0 && Kp > 0 && M > 1 && N > 0 && x_st >= 0 && x_st + n <= N,
                 "gpet_density_splat_f64: bad shape");
    GPET_SUPPORTED(B <= 65535, "gpet_density_splat_f64: B too large for one launch");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* grid = (unsigned long long*)work;
    double* scale = (double*)(grid + (size_t)B * M * N);
    int32_t* n_out = (int32_t*)(scale + B);
    cudaError_t e = cudaMemsetAsync(work, 0, ((size_t)B * M * N + B) * 8 + (size_t)B * Kp * 4, st);
    if (e != cudaSuccess) {
        set_error("density memset: %s", cudaGetErrorString(e));
        return GPET_ERR_CUDA;
    }
    const long long pts = (long long)Kp * n;
    dim3 g1((unsigned)((pts + 255) / 256), B);
    density_splat_kernel<<<g1, 256, 0, st>>>(Y, idx, wts, n, S, Kp, M, N, x_st, grid, n_out);
    return check_launch("density_splat_kernel");
}

extern "C" int gpet_density_finish_f64(const double* wts, int B, int n, int Kp, int M, int N, float* dens,
                                       uint32_t* minmax, void* work, void* stream) {
    GPET_REQUIRE(wts && dens && minmax && work, "gpet_density_finish_f64: null pointer");
    GPET_REQUIRE(B > 0 && n > 0 && Kp > 0 && M > 1 && N > 0, "gpet_density_finish_f64: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* grid = (unsigned long long*)work;
    double* scale = (double*)(grid + (size_t)B * M * N);
    int32_t* n_out = (int32_t*)(scale + B);
    density_scale_kernel<<<(B + 127) / 128, 128, 0, st>>>(wts, n_out, n, Kp, scale, B);
    init_minmax_kernel<<<(B + 255) / 256, 256, 0, st>>>(minmax, B);
    int rc = launch_blur9_u64(grid, B, M, N, scale, dens, minmax, st);
    if (rc) return rc;
    return check_launch("gpet_density_finish_f64");
}

extern "C" int gpet_density_f64(const double* Y, const int32_t* idx, const double* wts, int B, int n, int S, int Kp, int M,
                                int N, int x_st, float* dens, uint32_t* minmax, void* work, void* stream) {
    GPET_REQUIRE(dens && minmax, "gpet_density_f64: null pointer");
    int rc = gpet_density_splat_f64(Y, idx, wts, B, n, S, Kp, M, N, x_st, work, stream);
    if (rc) return rc;
    return gpet_density_finish_f64(wts, B, n, Kp, M, N, dens, minmax, work, stream);
}

extern "C" int gpet_select_f64(const float* dens, const uint32_t* minmax, const float* grad_kde, const int32_t* img_index,
                               int B, int M, int N,
                               const int32_t* col_bin, const int32_t* group_cols, int n_groups, const int32_t* old_yx,
                               const int32_t* n_old, int max_old, int nb, double* bin_score, int32_t* bin_pos,
                               void* stream) {
    GPET_REQUIRE(dens && minmax && grad_kde && col_bin && group_cols && old_yx && n_old && bin_score && bin_pos,
                 "gpet_select_f64: null pointer");
    GPET_REQUIRE(B > 0 && M > 0 && N > 0 && n_groups > 0 && nb > 0 && max_old >= 0, "gpet_select_f64: bad shape");
    GPET_SUPPORTED(B <= 65535, "gpet_select_f64: B too large for one launch");
    GPET_SUPPORTED((long long)M * N + max_old < 0x7fffffffLL, "gpet_select_f64: image too large for 32-bit positions");
    dim3 grid(n_groups, B);
    select_kernel<<<grid, SEL_THREADS, 0, (cudaStream_t)stream>>>(dens, minmax, grad_kde, img_index, nullptr, M, N, col_bin,
                                                                 group_cols, old_yx, n_old, max_old, nb, bin_score, bin_pos);
    return check_launch("select_kernel");
}

extern "C" int gpet_select_bands_f64(const float* dens, const uint32_t* minmax, const float* grad_kde,
                                     const int32_t* img_index, const int32_t* bands, int B, int M, int N,
                                     const int32_t* col_bin, const int32_t* group_cols, int n_groups, const int32_t* old_yx,
                                     const int32_t* n_old, int max_old, int nb, double* bin_score, int32_t* bin_pos,
                                     void* stream) {
    GPET_REQUIRE(dens && minmax && grad_kde && bands && col_bin && group_cols && old_yx && n_old && bin_score && bin_pos,
                 "gpet_select_bands_f64: null pointer");
    GPET_REQUIRE(B > 0 && M > 0 && N > 0 && n_groups > 0 && nb > 0 && max_old >= 0, "gpet_select_bands_f64: bad shape");
    GPET_SUPPORTED(B <= 65535, "gpet_select_bands_f64: B too large for one launch");
    GPET_SUPPORTED((long long)M * N + max_old < 0x7fffffffLL, "gpet_select_bands_f64: image too large for 32-bit positions");
    dim3 grid(n_groups, B);
    select_kernel<<<grid, SEL_THREADS, 0, (cudaStream_t)stream>>>(dens, minmax, grad_kde, img_index, bands, M, N, col_bin,
                                                                 group_cols, old_yx, n_old, max_old, nb, bin_score, bin_pos);
    return check_launch("select_kernel (bands)");
}
