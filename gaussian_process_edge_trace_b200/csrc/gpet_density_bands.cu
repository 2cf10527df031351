// Band-limited kernel density of the kept posterior curves: the whole of kernel_density_estimate (gpet.py:455-529)
// for one column group of one trace inside ONE CTA's shared memory.
//
// The general path (gpet_density.cu) keeps a 64-bit fixed-point grid of the whole image in HBM: memset (8 B/px), global
// atomics, a blur pass (8 B/px read, 4 B/px written) and the selection pass (8 B/px read) - 28 B per pixel and iteration
// although the kept curves cover a narrow band of rows.  Here a CTA owns the output columns [c0, c1) of one trace (a
// column group of gpet_select_f64: whole bins, a few tens of columns):
//   1. it reads the points of the kept curves that reach those columns through the 9-tap blur (columns c0-4 .. c1+3) from
//      a compact copy Yk[b][c][j] of the kept curves and finds the band of image rows they touch;
//   2. the 64-bit fixed-point histogram of that band (+4 rows / columns of halo) lives in shared memory (same linear
//      binning, same fixed-point sums as density_splat_kernel: order independent, bit-identical);
//   3. vertical and horizontal 9-tap passes run IN PLACE with a register window (one shared-memory load per output),
//      the same fused multiply-add chains as blur9_kernel, so every float32 density is bit-identical to the general path;
//   4. only the band rows are written to HBM (float32) together with the band itself; everything outside a band is
//      exactly zero by construction and is never stored or read again (gpet_select_f64 with `bands`).
// Per iteration this moves 4 B per BAND pixel instead of 28 B per image pixel.
#include <limits.h>
#include <math.h>

#include "gpet_common.cuh"

namespace gpet {

constexpr double DB_FX_SCALE = 1152921504606846976.0;  // 2^60 (as in gpet_density.cu)
constexpr int DB_MAX_W = 56;                            // widest column group (output columns of one CTA)

struct GaussTaps {
    double g[9];  // exp(-d^2/2), d = -4..4 (computed on the host like blur9_kernel's constant table)
};

__device__ __forceinline__ double fx_to_f64(unsigned long long v) {
    return __ull2double_rn(v) * 8.6736173798840355e-19;  // 2^-60, exact scaling (blur9_kernel's to_f64)
}

// linear binning of one point (KDEpy on the integer lattice, see density_splat_kernel): lower row r0, upper weight fy
__device__ __forceinline__ bool bin_point(double y, int M, int& r0, double& fy) {
    if (y < 0.0 || y > (double)(M - 1)) return false;  // dropped (gpet.py:498-500)
    const double ty = y + 1.0;
    const int iy = __double2int_rd(ty);
    fy = ty - (double)iy;
    r0 = iy - 1;
    return r0 >= 0 && r0 < M;  // false only for NaN
}

// Compact copy of the kept curves, Yk[b][c][j] = Y[b][j][idx[b][c]] (j contiguous: a CTA of the band kernel reads runs of
// consecutive columns of one curve), the exact count of dropped points per curve (gpet.py:498-500: points outside
// [0, M-1] are removed before the KDE) and, per curve column, the first and last image row the kept curves put weight on
// (colband[b][j] = {first, last}; {INT_MAX, -1} for a column without points) and, per point, its lower lattice row
// R0k[b][c][j] (int16; -1: no weight).  One CTA per 32 columns of one trace:
// lanes across the columns, warps across the kept curves.
constexpr int KG_THREADS = 256, KG_UNROLL = 4;
__global__ void __launch_bounds__(KG_THREADS)
keep_gather_kernel(const double* __restrict__ Y, const int32_t* __restrict__ idx, int n, int S, int Kp, int M,
                   double* __restrict__ Yk, int16_t* __restrict__ R0k, int32_t* __restrict__ n_out,
                   int32_t* __restrict__ colband) {
    __shared__ int s_lo[32], s_hi[32];
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = KG_THREADS / 32;
    const int j = blockIdx.x * 32 + lane;
    if (threadIdx.x < 32) {
        s_lo[threadIdx.x] = INT_MAX;
        s_hi[threadIdx.x] = -1;
    }
    __syncthreads();
    int lo = INT_MAX, hi = -1;
    const double* yrow = Y + ((size_t)b * n + min(j, n - 1)) * S;
    const int32_t* ib = idx + (size_t)b * Kp;
    for (int cb = warp; cb < Kp; cb += NW * KG_UNROLL) {
        double y[KG_UNROLL];
#pragma unroll
        for (int u = 0; u < KG_UNROLL; ++u) {       // the scattered loads of several curves are in flight together
            const int c = cb + u * NW;
            y[u] = (c < Kp && j < n) ? yrow[ib[c]] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < KG_UNROLL; ++u) {
            const int c = cb + u * NW;
            if (c >= Kp) break;                     // warp-uniform
            bool dropped = false;
            if (j < n) {
                int r0;
                double fy;
                int16_t r16 = -1;
                if (bin_point(y[u], M, r0, fy)) {
                    r16 = (int16_t)r0;
                    lo = min(lo, r0);
                    hi = max(hi, (fy > 0.0 && r0 + 1 < M) ? r0 + 1 : r0);
                } else {
                    dropped = (y[u] < 0.0 || y[u] > (double)(M - 1));
                }
                Yk[((size_t)b * Kp + c) * n + j] = y[u];
                R0k[((size_t)b * Kp + c) * n + j] = r16;
            }
            const unsigned int d = __ballot_sync(0xffffffffu, dropped);
            if (lane == 0 && d) atomicAdd(n_out + (size_t)b * Kp + c, __popc(d));
        }
    }
    if (hi >= 0) {
        atomicMin(&s_lo[lane], lo);
        atomicMax(&s_hi[lane], hi);
    }
    __syncthreads();
    if (threadIdx.x < 32 && j < n) {
        colband[((size_t)b * n + j) * 2] = s_lo[lane];
        colband[((size_t)b * n + j) * 2 + 1] = s_hi[lane];
    }
}

// KDEpy renormalises the point weights by their sum over the in-domain points (density_scale_kernel's fixed order);
// also resets the per-trace min/max.
__global__ void density_bands_prep_kernel(const double* __restrict__ wts, const int32_t* __restrict__ n_out, int n, int Kp,
                                          double* __restrict__ scale, uint32_t* __restrict__ minmax, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double t = 0.0;
    for (int c = 0; c < Kp; ++c) t += wts[(size_t)b * Kp + c] * (double)(n - n_out[(size_t)b * Kp + c]);
    scale[b] = (1.0 / t) * 0.15915494309189535;  // / sum(weights) / (2 pi)
    minmax[2 * b] = 0x7f800000u;                   // +inf
    minmax[2 * b + 1] = 0u;                        // +0
}

// One CTA = one column group of one trace; the histogram of the whole band lives in shared memory.
template <int DB_THREADS>
__global__ void __launch_bounds__(DB_THREADS, 1024 / DB_THREADS)
density_band_kernel(const double* __restrict__ Yk, const int16_t* __restrict__ R0k, const double* __restrict__ wts,
                    const double* __restrict__ scale, const int32_t* __restrict__ colband,
                    const int32_t* __restrict__ group_cols, int n, int Kp, int M, int N, int x_st, int HC,
                    float* __restrict__ dens, uint32_t* __restrict__ minmax, int32_t* __restrict__ bands,
                    const GaussTaps taps) {
    extern __shared__ double H[];  // [R + 8][HC]: fixed-point histogram -> vertical pass -> horizontal pass, in place
    __shared__ int s_rmin, s_rmax;
    __shared__ float s_vmin[DB_THREADS / 32], s_vmax[DB_THREADS / 32];
    const int b = blockIdx.y, g = blockIdx.x, tid = threadIdx.x;
    const int c0 = group_cols[g], c1 = group_cols[g + 1], W = c1 - c0;
    int32_t* band = bands + ((size_t)b * gridDim.x + g) * 2;
    // curve columns whose points reach the output columns [c0, c1) through the blur: x = x_st + j in [c0 - 4, c1 + 4)
    const int j0 = max(0, c0 - 4 - x_st), j1 = min(n, c1 + 4 - x_st);
    const int nj = (W > 0 && j1 > j0) ? j1 - j0 : 0;
    const int warp = tid >> 5, lane = tid & 31;
    constexpr int NW = DB_THREADS / 32;
    if (tid == 0) {
        s_rmin = INT_MAX;
        s_rmax = -1;
    }
    __syncthreads();
    // band of this group = union of its curve columns' bands (keep_gather_kernel)
    if (tid < nj) {
        const int lo = colband[((size_t)b * n + j0 + tid) * 2], hi = colband[((size_t)b * n + j0 + tid) * 2 + 1];
        if (hi >= 0) {
            atomicMin(&s_rmin, lo);
            atomicMax(&s_rmax, hi);
        }
    }
    __syncthreads();
    const int rmin = s_rmin, rmax = s_rmax;
    if (rmax < 0) {  // no point reaches this group: its columns are exactly zero
        if (tid == 0) {
            band[0] = 0;
            band[1] = 0;
            if (W > 0) atomic_minmax_nonneg(minmax + 2 * b, 0.0f, 0.0f);
        }
        return;
    }
    const int b0 = max(0, rmin - 4), b1 = min(M, rmax + 5);  // rows the vertical pass can make non-zero
    const int R = b1 - b0, RH = R + 8;                       // histogram rows b0 - 4 .. b1 + 3 (hist row = r - b0 + 4)
    const int HCu = W + 8;                                   // histogram columns c0 - 4 .. c1 + 3 (hist col = x - c0 + 4)
    unsigned long long* Hu = reinterpret_cast<unsigned long long*>(H);
    {
        ulonglong2* H2 = reinterpret_cast<ulonglong2*>(H);
        const int n2 = (RH * HC + 1) >> 1;                  // the allocation is (M + 8) * HC + 1 cells: rounding up is safe
        for (int i = tid; i < n2; i += DB_THREADS) H2[i] = make_ulonglong2(0ull, 0ull);
    }
    __syncthreads();
    // ---- linear binning into the shared-memory histogram (64-bit fixed point: order independent).  A 64-bit add is two
    //      native 32-bit shared-memory atomics: the low word's returned old value tells whether THIS add carried, and the
    //      carry joins the high word's add - every carry is counted exactly once, so the final cells are the exact sums
    //      (they are read after the barrier).  The 64-bit shared atomic itself compiles to a compare-and-swap loop.
    {
        unsigned int* H32 = reinterpret_cast<unsigned int*>(H);
        const int hoff = x_st + j0 - c0 + 4;                 // histogram column of curve column j0
        auto add64 = [&](int cell, unsigned long long v) {
            const unsigned int lo = (unsigned int)v, hi = (unsigned int)(v >> 32);
            const unsigned int old = atomicAdd(H32 + 2 * cell, lo);
            const unsigned int carry = (old + lo < old) ? 1u : 0u;
            if (hi + carry) atomicAdd(H32 + 2 * cell + 1, hi + carry);
        };
        // flat point index p = c * nj + jj, consecutive lanes = consecutive columns of a curve (coalesced reads; two lanes
        // of a warp share a histogram column only when a group has fewer than 32 curve columns)
        const int npts = nj * Kp;
        if (tid < npts) {
            const int dq = DB_THREADS / nj, dr = DB_THREADS - dq * nj;     // p += DB_THREADS  <=>  (c, jj) += (dq, dr)
            int c = tid / nj, jj = tid - c * nj;
            const double* yb = Yk + (size_t)b * Kp * n + j0;
            const int16_t* rb = R0k + (size_t)b * Kp * n + j0;
            const double* wb = wts + (size_t)b * Kp;
            // the loads of up to four points of a thread are issued together (two dependent global round trips per point
            // otherwise: the row index, then the curve value)
            for (int p = tid; p < npts; p += 4 * DB_THREADS) {
                int r0[4], cc[4], cl[4];
                double yv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const bool ok = p + u * DB_THREADS < npts;
                    const int at = c * n + jj;
                    r0[u] = ok ? (int)rb[at] : -1;
                    yv[u] = ok ? yb[at] : 0.0;
                    cc[u] = c;
                    cl[u] = hoff + jj;
                    c += dq;
                    jj += dr;
                    if (jj >= nj) {
                        jj -= nj;
                        ++c;
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (r0[u] >= 0) {
                        const double fy = (yv[u] + 1.0) - (double)(r0[u] + 1);      // bin_point's upper weight
                        const double w = wb[cc[u]];
                        const int cell = (r0[u] - b0 + 4) * HC + cl[u];
                        add64(cell, __double2ull_rn(((1.0 - fy) * w) * DB_FX_SCALE));
                        if (fy > 0.0 && r0[u] + 1 < M) add64(cell + HC, __double2ull_rn((fy * w) * DB_FX_SCALE));
                    }
                }
            }
        }
    }
    __syncthreads();
    // ---- vertical 9-tap pass, in place: thread = (histogram column, chunk of rows), register window ------------------
    {
        const int nchunks = DB_THREADS / HCu;
        const int rows_per = (R + nchunks - 1) / nchunks;
        const int h = tid % HCu, ch = tid / HCu;
        const int a = 4 + ch * rows_per, e = min(4 + R, a + rows_per);  // histogram rows [a, e) of this thread
        const bool act = ch < nchunks && a < e;
        double w[9], bh[4];
        if (act) {
            // the rows above (a-4 .. a-1) and below (e .. e+3) belong to other threads, which overwrite them: read first
#pragma unroll
            for (int k = 0; k < 8; ++k) w[k] = fx_to_f64(Hu[(a - 4 + k) * HC + h]);
#pragma unroll
            for (int k = 0; k < 4; ++k) bh[k] = fx_to_f64(Hu[(e + k) * HC + h]);
        }
        __syncthreads();
        if (act) {
            const unsigned long long* hp = Hu + (a + 4) * HC + h;
            double* op = H + a * HC + h;
            int left = e - a - 4;                              // rows still to be read from the histogram itself
            for (int r = a; r < e; r += 9) {
#pragma unroll
                for (int u = 0; u < 9; ++u) {
                    if (r + u < e) {
                        double nv;
                        if (left > 0) {
                            nv = fx_to_f64(*hp);
                        } else {
                            nv = left == 0 ? bh[0] : (left == -1 ? bh[1] : (left == -2 ? bh[2] : bh[3]));
                        }
                        --left;
                        hp += HC;
                        w[(u + 8) % 9] = nv;
                        double acc = 0.0;
#pragma unroll
                        for (int d = 0; d < 9; ++d) acc = fma(w[(u + d) % 9], taps.g[d], acc);
                        *op = acc;
                        op += HC;
                    }
                }
            }
        }
    }
    __syncthreads();
    // ---- horizontal 9-tap pass, in place: one thread per band row (odd row stride => conflict-free) ------------------
    for (int rr = 4 + tid; rr < 4 + R; rr += DB_THREADS) {
        double* row = H + rr * HC;
        double w[9];
#pragma unroll
        for (int k = 0; k < 8; ++k) w[k] = row[k];
        for (int o = 4; o < 4 + W; o += 9) {
#pragma unroll
            for (int u = 0; u < 9; ++u) {
                const int oo = o + u;
                if (oo < 4 + W) {
                    w[(u + 8) % 9] = row[oo + 4];
                    double acc = 0.0;
#pragma unroll
                    for (int d = 0; d < 9; ++d) acc = fma(w[(u + d) % 9], taps.g[d], acc);
                    row[oo] = acc;
                }
            }
        }
    }
    __syncthreads();
    // ---- float32 cast, store of the band rows, min / max: one band row per warp trip, lanes across the columns -------
    const double sc = scale[b];
    float vmin = __int_as_float(0x7f800000), vmax = 0.0f;
    for (int xx = lane; xx < W; xx += 32) {
        const double* hp = H + (warp + 4) * HC + xx + 4;
        float* dp = dens + ((size_t)b * M + b0 + warp) * N + c0 + xx;
        for (int r = warp; r < R; r += NW) {
            const float v = __double2float_rn(*hp * sc);
            *dp = v;
            vmin = fminf(vmin, v + 0.0f);
            vmax = fmaxf(vmax, v + 0.0f);
            hp += NW * HC;
            dp += (size_t)NW * N;
        }
    }
    if (R < M) vmin = fminf(vmin, 0.0f);  // the rows outside the band are exactly zero
    vmin = warp_min(vmin);
    vmax = warp_max(vmax);
    if (lane == 0) {
        s_vmin[warp] = vmin;
        s_vmax[warp] = vmax;
    }
    __syncthreads();
    if (tid == 0) {
        for (int k = 1; k < NW; ++k) {
            vmin = fminf(vmin, s_vmin[k]);
            vmax = fmaxf(vmax, s_vmax[k]);
        }
        atomic_minmax_nonneg(minmax + 2 * b, vmin, vmax);
        band[0] = b0;
        band[1] = b1;
    }
}

// normalised kde map (float32) from band-limited densities, for inspection / tests: zero outside the bands
__global__ void __launch_bounds__(256)
kde_bands_kernel(const float* __restrict__ dens, const uint32_t* __restrict__ minmax, const int32_t* __restrict__ bands,
                 const int32_t* __restrict__ group_cols, int M, int N, float* __restrict__ kde) {
    const int b = blockIdx.y, g = blockIdx.x;
    const int c0 = group_cols[g], W = group_cols[g + 1] - c0;
    if (W <= 0) return;
    const int b0 = bands[((size_t)b * gridDim.x + g) * 2], b1 = bands[((size_t)b * gridDim.x + g) * 2 + 1];
    const float mn = __uint_as_float(minmax[2 * b]);
    const float range = __fsub_rn(__uint_as_float(minmax[2 * b + 1]), mn);
    for (int i = threadIdx.x; i < M * W; i += blockDim.x) {
        const int r = i / W, x = c0 + (i - r * W);
        const size_t at = ((size_t)b * M + r) * N + x;
        kde[at] = normalise_f32((r >= b0 && r < b1) ? dens[at] : 0.0f, mn, range);
    }
}

static int bands_hist_stride(int max_width) { return (max_width + 8) | 1; }  // odd row stride (doubles)

}  // namespace gpet

using namespace gpet;

static size_t bands_smem_bytes(int M, int max_width) {
    return ((size_t)(M + 8) * bands_hist_stride(max_width) + 1) * sizeof(double);
}

extern "C" int gpet_density_bands_supported(int M, int N, int max_width) {
    if (M < 2 || M > 32767 || N < 1 || max_width < 1 || max_width > DB_MAX_W) return 0;
    return bands_smem_bytes(M, max_width) <= (size_t)220 * 1024 ? 1 : 0;
}

extern "C" int64_t gpet_density_bands_workspace_bytes(int B, int n, int Kp) {
    // f64 Yk[B][Kp][n] | f64 scale[B] | i32 n_out[B][Kp] | i32 colband[B][n][2] | i16 R0k[B][Kp][n]
    return (int64_t)B * n * Kp * 8 + (int64_t)B * 8 + (int64_t)B * Kp * 4 + (int64_t)B * n * 8 + (int64_t)B * n * Kp * 2 + 256;
}

extern "C" int gpet_density_bands_f64(const double* Y, const int32_t* idx, const double* wts, int B, int n, int S, int Kp,
                                      int M, int N, int x_st, const int32_t* group_cols, int n_groups, int max_width,
                                      float* dens, uint32_t* minmax, int32_t* bands, void* work, void* stream) {
    GPET_REQUIRE(Y && idx && wts && group_cols && dens && minmax && bands && work, "gpet_density_bands_f64: null pointer");
    GPET_REQUIRE(B > 0 && n > 0 && S > 0 && Kp > 0 && M > 1 && N > 0 && x_st >= 0 && x_st + n <= N && n_groups > 0,
                 "gpet_density_bands_f64: bad shape");
    GPET_SUPPORTED(B <= 65535, "gpet_density_bands_f64: B too large for one launch");
    GPET_SUPPORTED(gpet_density_bands_supported(M, N, max_width),
                   "gpet_density_bands_f64: M=%d rows x %d columns per group do not fit shared memory", M, max_width);
    cudaStream_t st = (cudaStream_t)stream;
    double* Yk = (double*)work;
    double* scale = Yk + (size_t)B * n * Kp;
    int32_t* n_out = (int32_t*)(scale + B);
    cudaError_t e = cudaMemsetAsync(n_out, 0, (size_t)B * Kp * 4, st);
    if (e != cudaSuccess) {
        set_error("density bands memset: %s", cudaGetErrorString(e));
        return GPET_ERR_CUDA;
    }
    int32_t* colband = n_out + (size_t)B * Kp;
    int16_t* R0k = (int16_t*)(colband + (size_t)B * n * 2);
    GPET_SUPPORTED(M <= 32767, "gpet_density_bands_f64: M=%d rows (16-bit row indices)", M);
    keep_gather_kernel<<<dim3((n + 31) / 32, B), KG_THREADS, 0, st>>>(Y, idx, n, S, Kp, M, Yk, R0k, n_out, colband);
    density_bands_prep_kernel<<<(B + 127) / 128, 128, 0, st>>>(wts, n_out, n, Kp, scale, minmax, B);
    GaussTaps taps;
    for (int d = -4; d <= 4; ++d) taps.g[d + 4] = exp(-0.5 * (double)(d * d));
    const int HC = bands_hist_stride(max_width);
    const size_t smem = bands_smem_bytes(M, max_width);
    const bool two = smem <= (size_t)110 * 1024;      // two CTAs per SM hide each other's barriers
    e = two ? cudaFuncSetAttribute(density_band_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
            : cudaFuncSetAttribute(density_band_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("density_band_kernel smem attribute: %s", cudaGetErrorString(e));
        return GPET_ERR_CUDA;
    }
    if (two)
        density_band_kernel<512><<<dim3(n_groups, B), 512, smem, st>>>(Yk, R0k, wts, scale, colband, group_cols, n, Kp, M, N, x_st,
                                                                       HC, dens, minmax, bands, taps);
    else
    density_band_kernel<1024><<<dim3(n_groups, B), 1024, smem, st>>>(Yk, R0k, wts, scale, colband, group_cols, n, Kp, M, N, x_st,
                                                                     HC,
                                                                     dens, minmax, bands, taps);
    return check_launch("density_band_kernel");
}

extern "C" int gpet_kde_bands_f32(const float* dens, const uint32_t* minmax, const int32_t* bands,
                                  const int32_t* group_cols, int n_groups, int B, int M, int N, float* kde, void* stream) {
    GPET_REQUIRE(dens && minmax && bands && group_cols && kde && B > 0 && M > 0 && N > 0 && n_groups > 0,
                 "gpet_kde_bands_f32: bad argument");
    GPET_SUPPORTED(B <= 65535, "gpet_kde_bands_f32: B too large for one launch");
    kde_bands_kernel<<<dim3(n_groups, B), 256, 0, (cudaStream_t)stream>>>(dens, minmax, bands, group_cols, M, N, kde);
    return check_launch("kde_bands_kernel");
}
