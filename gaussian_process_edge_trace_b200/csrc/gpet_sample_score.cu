// Fused posterior-curve sampling + scoring: the curves of an iteration never travel through HBM.
//
// Reference seams: sklearn_gpr.py:460-464 (sample_y: Z @ A + mean), gpet.py:261 (* y_s), gpet.py:437-440 + 371-410
// (cost_funct of every sampled curve).  The unfused pair gpet_sample_f64 -> gpet_score_f64 writes Y[b][n][S] (4 MB per
// trace and iteration at cfg 1 sizes) and reads it back; here a CTA owns 64 curves of one trace and walks down the n grid
// columns in chunks of 32:
//   * 4 tensor warps form the chunk  Y[32 x 64] = A[:, chunk]^T Z[:, tile]  with mma.sync.m8n8k4.f64 (DMMA; tcgen05 has
//     no FP64 kind).  Their Z fragments (rp x 16 values per warp) stay in REGISTERS for the whole kernel, the A chunk
//     (shared by the 16 CTAs of a trace, L2 resident) is streamed with cp.async into a double buffer; the finished chunk,
//     scaled and shifted (y_s (acc + mean)), goes to a double-buffered shared-memory tile;
//   * 2 scoring warps (one curve per thread) consume that tile with the arithmetic of gpet_score_math.cuh - the same
//     Simpson state machine as the stand-alone scoring kernels - gathering the gradient taps a sub-block of 8 columns
//     ahead of their use.
// The two groups meet only at named barriers (full / empty per buffer).  The kept curves (N_keep of N_samples, 10 %)
// are recomputed afterwards by sample_keep_kernel with the same DMMA chain (bit-identical values) for the density splat.
#include "gpet_common.cuh"
#include "gpet_score_math.cuh"

namespace gpet {

constexpr int FS_TS = 64;             // curves per CTA
constexpr int FS_TJ = 32;             // grid columns per chunk
constexpr int FS_MMA_T = 128;         // 4 tensor warps
constexpr int FS_SC_T = 64;           // 2 scoring warps
constexpr int FS_T = FS_MMA_T + FS_SC_T;
constexpr int FS_LDA = FS_TJ + 4;     // 36 == 4 (mod 16): conflict-free DMMA operand fetches
constexpr int FS_LDY = FS_TS + 8;     // 72 == 8 (mod 16): conflict-free double2 stores of the accumulator fragments
constexpr int FS_KS = 20;             // k-steps of 4 held in registers: rp <= 80

__device__ __forceinline__ void fs_dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void fs_cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void fs_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void fs_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int count) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void fs_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// barrier ids: 1, 2 = chunk buffer 0 / 1 full;  3, 4 = buffer 0 / 1 consumed;  5 = tensor warps among themselves
template <bool SCAN>
__global__ void __launch_bounds__(FS_T, 2)
sample_score_kernel(const double* __restrict__ Zt, const double* __restrict__ A, const double* __restrict__ mean,
                    const double* __restrict__ ys, const float* __restrict__ gradT,
                    const int32_t* __restrict__ img_index, int rp, int n, int S, int M, int N, int x_st,
                    double* __restrict__ cost) {
    extern __shared__ __align__(16) double sm[];
    double* As = sm;                                      // [2][rp][FS_LDA]
    double* Yb = sm + 2 * (size_t)rp * FS_LDA;            // [2][FS_TJ][FS_LDY]
    const int b = blockIdx.y, s0 = blockIdx.x * FS_TS, tid = threadIdx.x;
    const int nch = (n + FS_TJ - 1) / FS_TJ;
    const int Mp = M + 2, Mm1 = M - 1;
    const int img = img_index ? img_index[b] : b;
    const float* gt = gradT + ((size_t)img * N + x_st) * Mp + 1;     // row 0 of the column of sample 0

    if (tid < FS_MMA_T) {
        // ------------------------------------------------ tensor warps ------------------------------------------------
        const int lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
        const int ws = 16 * w, nks = rp >> 2;
        const double* Ab = A + (size_t)b * rp * n;
        auto load_chunk = [&](int c) {
            double* dst = As + (size_t)(c & 1) * rp * FS_LDA;
            const int j0 = c * FS_TJ;
            for (int p = tid; p < rp * (FS_TJ / 2); p += FS_MMA_T) {
                const int k = p / (FS_TJ / 2), q = p - k * (FS_TJ / 2);
                const int j = j0 + 2 * q;                               // n is even: a pair is inside or outside
                fs_cp_async16(dst + k * FS_LDA + 2 * q, Ab + (size_t)k * n + (j < n ? j : 0), j < n ? 16 : 0);
            }
            fs_commit();
        };
        load_chunk(0);
        double bf[FS_KS][2];
#pragma unroll
        for (int ks = 0; ks < FS_KS; ++ks)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int s = s0 + ws + 8 * c + g;
                bf[ks][c] = (ks < nks && s < S) ? __ldg(Zt + (size_t)(4 * ks + t) * S + s) : 0.0;
            }
        const double y_s = ys[b];
        const char* gcol = reinterpret_cast<const char*>(gt - 1);
        const size_t chunk_cols_bytes = (size_t)FS_TJ * Mp * sizeof(float);
#pragma unroll 1
        for (int c = 0; c < nch; ++c) {
            fs_wait_all();
            bar_sync(5, FS_MMA_T);           // chunk c has landed for everybody; everybody is done with chunk c - 1
            if (c + 1 < nch) {
                load_chunk(c + 1);
                // the gradient columns the scoring warps will gather from during the next chunk -> L2
                const size_t lim = min(chunk_cols_bytes, (size_t)(n - (c + 1) * FS_TJ) * Mp * sizeof(float));
                for (size_t o = (size_t)tid * 128; o < lim; o += FS_MMA_T * 128)
                    fs_prefetch_l2(gcol + (size_t)(c + 1) * chunk_cols_bytes + o);
            }
            const double* as = As + (size_t)(c & 1) * rp * FS_LDA;
            const int j0 = c * FS_TJ;
            double mu[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int j = j0 + a * 8 + g;
                mu[a] = (j < n) ? __ldg(mean + (size_t)b * n + j) : 0.0;
            }
            double acc[4][2][2];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) acc[a][cc][0] = acc[a][cc][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < FS_KS; ++ks) {
                if (ks < nks) {
                    const double* ap = as + (4 * ks + t) * FS_LDA + g;
                    double af[4];
#pragma unroll
                    for (int a = 0; a < 4; ++a) af[a] = ap[a * 8];
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int cc = 0; cc < 2; ++cc) fs_dmma(acc[a][cc][0], acc[a][cc][1], af[a], bf[ks][cc]);
                }
            }
            if (c >= 2) bar_sync(3 + (c & 1), FS_T);      // the scoring warps have consumed chunk c - 2 of this buffer
            double* yb = Yb + (size_t)(c & 1) * FS_TJ * FS_LDY;
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const double v0 = y_s * (acc[a][cc][0] + mu[a]), v1 = y_s * (acc[a][cc][1] + mu[a]);
                    *reinterpret_cast<double2*>(yb + (a * 8 + g) * FS_LDY + ws + 8 * cc + 2 * t) = make_double2(v0, v1);
                }
            __threadfence_block();
            bar_arrive(1 + (c & 1), FS_T);                  // chunk c is in shared memory
        }
        return;
    }

    // ---------------------------------------------------- scoring warps ---------------------------------------------
    const int sl = tid - FS_MMA_T;
    const int s = s0 + sl;
    CurveState cs;
    double tfirst = 0.0;
    Taps carry;
    carry.g0 = carry.g1 = 0.0f;
    carry.f = 0.0;
#pragma unroll 1
    for (int c = 0; c < nch; ++c) {
        bar_sync(1 + (c & 1), FS_T);
        const double* yb = Yb + (size_t)(c & 1) * FS_TJ * FS_LDY + sl;
        const int j0 = c * FS_TJ;
        Taps T[2][8];
        auto fetch8 = [&](Taps* dst, int rb) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int j = j0 + rb + i;
                if (j < n) dst[i] = fetch_taps_off(gt, j * Mp, yb[(rb + i) * FS_LDY], Mm1);
            }
        };
        fetch8(T[0], 0);
#pragma unroll
        for (int sb = 0; sb < 4; ++sb) {
            if (sb + 1 < 4) fetch8(T[(sb + 1) & 1], 8 * (sb + 1));
            const Taps* tp = T[sb & 1];
            const int r = 8 * sb, j = j0 + r;
            if (j < n) {
                if (j == 0) {
                    curve_begin<SCAN>(cs, yb[0], yb[FS_LDY], gt, Mm1, tfirst);
                } else {
                    simpson_pair_math<SCAN>(cs, yb[r * FS_LDY], yb[(r + 1) * FS_LDY], carry, tp[0]);
                }
                if (j + 2 < n) simpson_pair_math<SCAN>(cs, yb[(r + 2) * FS_LDY], yb[(r + 3) * FS_LDY], tp[1], tp[2]);
                if (j + 4 < n) simpson_pair_math<SCAN>(cs, yb[(r + 4) * FS_LDY], yb[(r + 5) * FS_LDY], tp[3], tp[4]);
                if (j + 6 < n) simpson_pair_math<SCAN>(cs, yb[(r + 6) * FS_LDY], yb[(r + 7) * FS_LDY], tp[5], tp[6]);
                carry = tp[7];
            }
        }
        if (c + 2 < nch) {
            __threadfence_block();
            bar_arrive(3 + (c & 1), FS_T);
        }
    }
    if (s < S) cost[(size_t)b * S + s] = curve_cost<SCAN>(cs, tfirst);
}

// ---- the kept curves, recomputed ---------------------------------------------------------------------------------------
// Ykeep[b][j][c] = ys[b] (sum_k A[b][k][j] Zt[k][idx[b][c]] + mean[b][j]), c < Kp: the same DMMA chain over k as the fused
// kernel (k-steps of 4 in ascending order into one accumulator), hence the same bits the scoring saw.  idx < 0 (a curve
// owned by another rank of a sample-sharded run) gives a zero column.  One 64 x 64 tile per CTA, K in chunks of 40.
constexpr int SK_TJ = 64, SK_TC = 64, SK_LD = 68, SK_T = 256, SK_KC = 40;

__global__ void __launch_bounds__(SK_T, 4)
sample_keep_kernel(const double* __restrict__ Zt, const double* __restrict__ A, const double* __restrict__ mean,
                   const double* __restrict__ ys, const int32_t* __restrict__ idx, int rp, int n, int S, int Kp,
                   double* __restrict__ Yk) {
    __shared__ double As[SK_KC * SK_LD];
    __shared__ double Zs[SK_KC * SK_LD];
    __shared__ int sidx[SK_TC];
    const int b = blockIdx.z, j0 = blockIdx.y * SK_TJ, c0 = blockIdx.x * SK_TC;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* Ab = A + (size_t)b * rp * n;
    if (tid < SK_TC) sidx[tid] = (c0 + tid < Kp) ? idx[(size_t)b * Kp + c0 + tid] : -1;
    const int wj = (warp >> 2) * 32, wc = (warp & 3) * 16;
    const int g = lane >> 2, t = lane & 3;
    double acc[4][2][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 2; ++c) acc[a][c][0] = acc[a][c][1] = 0.0;
    for (int kc0 = 0; kc0 < rp; kc0 += SK_KC) {
        const int kc = min(SK_KC, rp - kc0);
        __syncthreads();
        for (int p = tid; p < kc * SK_TJ; p += SK_T) {
            const int k = p / SK_TJ, c = p - k * SK_TJ;
            const size_t kg = (size_t)(kc0 + k);
            As[k * SK_LD + c] = (j0 + c < n) ? Ab[kg * n + j0 + c] : 0.0;
            const int si = sidx[c];
            Zs[k * SK_LD + c] = (si >= 0) ? __ldg(Zt + kg * S + si) : 0.0;
        }
        __syncthreads();
        for (int k0 = 0; k0 < kc; k0 += 4) {
            double af[4], bf[2];
            const double* ap = As + (k0 + t) * SK_LD + wj + g;
            const double* zp = Zs + (k0 + t) * SK_LD + wc + g;
#pragma unroll
            for (int a = 0; a < 4; ++a) af[a] = ap[a * 8];
#pragma unroll
            for (int c = 0; c < 2; ++c) bf[c] = zp[c * 8];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = 0; c < 2; ++c) fs_dmma(acc[a][c][0], acc[a][c][1], af[a], bf[c]);
        }
    }
    const double y_s = ys[b];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int j = j0 + wj + a * 8 + g;
        if (j >= n) continue;
        const double mu = mean[(size_t)b * n + j];
        double* row = Yk + ((size_t)b * n + j) * Kp;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int cc = c0 + wc + c * 8 + t * 2;
            if (cc < Kp) row[cc] = y_s * (acc[a][c][0] + mu);
            if (cc + 1 < Kp) row[cc + 1] = y_s * (acc[a][c][1] + mu);
        }
    }
}

}  // namespace gpet

using namespace gpet;

extern "C" int gpet_sample_score_supported(int rp, int n, int S) {
    return (rp >= 4 && (rp % 4) == 0 && rp <= 4 * FS_KS && n >= 4 && (n % 2) == 0 && S > 0) ? 1 : 0;
}

extern "C" int gpet_sample_score_f64(const double* Zt, const double* A, const double* mean, const double* ys,
                                     const float* gradT, const int32_t* img_index, int B, int rp, int n, int S, int M,
                                     int N, int x_st, double* cost, void* stream) {
    GPET_REQUIRE(Zt && A && mean && ys && gradT && cost && B > 0 && S > 0 && M >= 2, "gpet_sample_score_f64: bad argument");
    GPET_REQUIRE(x_st >= 0 && x_st + n <= N, "gpet_sample_score_f64: edge span outside the image");
    GPET_SUPPORTED(gpet_sample_score_supported(rp, n, S),
                   "gpet_sample_score_f64: needs rp %% 4 == 0, rp <= %d and an even edge_length (rp=%d, n=%d); use "
                   "gpet_sample_f64 + gpet_score_f64", 4 * FS_KS, rp, n);
    GPET_SUPPORTED(((uintptr_t)A % 16) == 0, "gpet_sample_score_f64: A must be 16-byte aligned");
    GPET_SUPPORTED(B <= 65535, "gpet_sample_score_f64: B too large for one launch");
    const size_t smem = (2 * (size_t)rp * FS_LDA + 2 * (size_t)FS_TJ * FS_LDY) * sizeof(double);
    const bool scan = g_tune[GPET_TUNE_SCORE_SCAN] != 0;
    cudaError_t e = scan ? cudaFuncSetAttribute(sample_score_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                         : cudaFuncSetAttribute(sample_score_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("sample_score smem attribute: %s", cudaGetErrorString(e));
        return GPET_ERR_CUDA;
    }
    dim3 grid((S + FS_TS - 1) / FS_TS, B);
    if (scan)
        sample_score_kernel<true><<<grid, FS_T, smem, (cudaStream_t)stream>>>(Zt, A, mean, ys, gradT, img_index, rp, n, S, M, N,
                                                                            x_st, cost);
    else
        sample_score_kernel<false><<<grid, FS_T, smem, (cudaStream_t)stream>>>(Zt, A, mean, ys, gradT, img_index, rp, n, S, M, N,
                                                                             x_st, cost);
    return check_launch("sample_score_kernel");
}

extern "C" int gpet_sample_keep_f64(const double* Zt, const double* A, const double* mean, const double* ys,
                                    const int32_t* idx, int B, int rp, int n, int S, int Kp, double* Ykeep, void* stream) {
    GPET_REQUIRE(Zt && A && mean && ys && idx && Ykeep && B > 0 && n > 0 && S > 0 && Kp > 0, "gpet_sample_keep_f64: bad argument");
    GPET_SUPPORTED(rp >= 4 && (rp % 4) == 0, "gpet_sample_keep_f64: rp=%d must be a multiple of 4", rp);
    GPET_SUPPORTED(B <= 65535 && (n + SK_TJ - 1) / SK_TJ <= 65535, "gpet_sample_keep_f64: grid too large");
    dim3 grid((Kp + SK_TC - 1) / SK_TC, (n + SK_TJ - 1) / SK_TJ, B);
    sample_keep_kernel<<<grid, SK_T, 0, (cudaStream_t)stream>>>(Zt, A, mean, ys, idx, rp, n, S, Kp, Ykeep);
    return check_launch("sample_keep_kernel");
}
