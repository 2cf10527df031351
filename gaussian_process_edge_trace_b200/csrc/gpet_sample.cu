// Posterior-curve sampling:  Y[b][j][s] = ys[b] * ( sum_k A[b][k][j] * Zt[k][s] + mean[b][j] ).
//
// Reference seam: sklearn_gpr.py:460-464 (rng.multivariate_normal(mean, cov, S).T = (Z @ A + mean).T) and
// gpet.py:261 (* y_s).  Z is shared by every trace of the batch (same seed), so the batch is one
// (B*n) x rp x S contraction.  fp64 on the tensor pipe: mma.sync.m8n8k4.f64 (DMMA) - tcgen05 has no FP64.
//
// Tiling: CTA = 64 (j) x 64 (s) outputs, K = rp in chunks of <= 96; 8 warps as 2 (j) x 4 (s), warp tile 32 x 16 =
// 4 x 2 DMMA tiles.  Operand tiles are kept k-major in shared memory with leading dimension == 4 (mod 16)
// doubles so that the 16 lanes of a half-warp hit 16 distinct 8-byte banks.
#include "gpet_common.cuh"

namespace gpet {

constexpr int SM_TJ = 64, SM_TS = 64, SM_LD = 68, SM_THREADS = 256, SM_KC = 40;

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(SM_THREADS, 4)
sample_dmma_kernel(const double* __restrict__ Zt, const double* __restrict__ A, const double* __restrict__ mean,
                   const double* __restrict__ ys, int rp, int n, int S, double* __restrict__ Y) {
    extern __shared__ double sm[];
    const int kc_max = rp < SM_KC ? rp : SM_KC;
    double* As = sm;                            // kc x SM_LD   As[k][j]
    double* Zs = sm + (size_t)kc_max * SM_LD;   // kc x SM_LD   Zs[k][s]
    const int b = blockIdx.z, j0 = blockIdx.y * SM_TJ, s0 = blockIdx.x * SM_TS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* Ab = A + (size_t)b * rp * n;
    const int wj = (warp >> 2) * 32, ws = (warp & 3) * 16;
    const int g = lane >> 2, t = lane & 3;
    double acc[4][2][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 2; ++c) acc[a][c][0] = acc[a][c][1] = 0.0;
    for (int kc0 = 0; kc0 < rp; kc0 += SM_KC) {
        const int kc = (rp - kc0) < SM_KC ? (rp - kc0) : SM_KC;
        if (kc0) __syncthreads();
        // cooperative tile loads: rows of 64 doubles, 2 doubles per thread-step
        for (int p = tid; p < kc * (SM_TJ / 2); p += SM_THREADS) {
            const int k = p / (SM_TJ / 2), c = (p - k * (SM_TJ / 2)) * 2;
            const int j = j0 + c, s = s0 + c;
            const size_t kg = (size_t)(kc0 + k);
            double a0 = 0.0, a1 = 0.0, z0 = 0.0, z1 = 0.0;
            if (j < n) a0 = Ab[kg * n + j];
            if (j + 1 < n) a1 = Ab[kg * n + j + 1];
            if (s < S) z0 = Zt[kg * S + s];
            if (s + 1 < S) z1 = Zt[kg * S + s + 1];
            As[k * SM_LD + c] = a0;
            As[k * SM_LD + c + 1] = a1;
            Zs[k * SM_LD + c] = z0;
            Zs[k * SM_LD + c + 1] = z1;
        }
        __syncthreads();
        for (int k0 = 0; k0 < kc; k0 += 4) {
            double af[4], bf[2];
            const double* ap = As + (k0 + t) * SM_LD + wj + g;
            const double* zp = Zs + (k0 + t) * SM_LD + ws + g;
#pragma unroll
            for (int a = 0; a < 4; ++a) af[a] = ap[a * 8];
#pragma unroll
            for (int c = 0; c < 2; ++c) bf[c] = zp[c * 8];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = 0; c < 2; ++c) dmma_m8n8k4(acc[a][c][0], acc[a][c][1], af[a], bf[c]);
        }
    }
    const double y_s = ys[b];
    const bool vec_ok = (S % 2) == 0;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int j = j0 + wj + a * 8 + g;
        if (j >= n) continue;
        const double mu = mean[(size_t)b * n + j];
        double* row = Y + ((size_t)b * n + j) * S;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int s = s0 + ws + c * 8 + t * 2;
            const double v0 = y_s * (acc[a][c][0] + mu), v1 = y_s * (acc[a][c][1] + mu);
            if (vec_ok && s + 1 < S) {
                *reinterpret_cast<double2*>(row + s) = make_double2(v0, v1);
            } else {
                if (s < S) row[s] = v0;
                if (s + 1 < S) row[s + 1] = v1;
            }
        }
    }
}

// ---- persistent-row variant ----------------------------------------------------------------------------------------------
// One CTA per (trace, 64 grid columns j): its A tile [rp][64] is loaded ONCE and stays in shared memory while the CTA
// walks over all sample tiles; the Z tiles (shared by every trace, L2 resident) are streamed in K chunks of 40 rows
// through a two-stage cp.async pipeline, so the loads of the next chunk overlap the DMMAs of the current one.  Against
// the tile-per-CTA kernel above: 45 % less operand traffic and no load phase in front of every 64x64 tile.
constexpr int SP_KC = 40;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(SM_THREADS, 2)
sample_rows_kernel(const double* __restrict__ Zt, const double* __restrict__ A, const double* __restrict__ mean,
                   const double* __restrict__ ys, int rp, int n, int S, double* __restrict__ Y) {
    extern __shared__ double sm[];
    double* As = sm;                                  // rp x SM_LD        As[k][j]
    double* Zs = sm + (size_t)rp * SM_LD;             // 2 x SP_KC x SM_LD Zs[stage][k][s]
    const int b = blockIdx.y, j0 = blockIdx.x * SM_TJ;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* Ab = A + (size_t)b * rp * n;
    const int wj = (warp >> 2) * 32, ws = (warp & 3) * 16;
    const int g = lane >> 2, t = lane & 3;
    // A tile: rows of 64 doubles, 2 doubles per thread-step (zero beyond n)
    for (int p = tid; p < rp * (SM_TJ / 2); p += SM_THREADS) {
        const int k = p / (SM_TJ / 2), c = (p - k * (SM_TJ / 2)) * 2;
        const int j = j0 + c;
        As[k * SM_LD + c] = (j < n) ? Ab[(size_t)k * n + j] : 0.0;
        As[k * SM_LD + c + 1] = (j + 1 < n) ? Ab[(size_t)k * n + j + 1] : 0.0;
    }
    const int n_st = (S + SM_TS - 1) / SM_TS;         // sample tiles
    const int n_ch = (rp + SP_KC - 1) / SP_KC;        // K chunks per tile
    const int total = n_st * n_ch;
    auto prefetch = [&](int it) {                     // stage it & 1 <- (sample tile it / n_ch, chunk it % n_ch)
        const int st = it / n_ch, ch = it - st * n_ch;
        const int kc0 = ch * SP_KC, kc = min(SP_KC, rp - kc0), s0 = st * SM_TS;
        double* dst = Zs + (size_t)(it & 1) * SP_KC * SM_LD;
        for (int p = tid; p < kc * (SM_TS / 2); p += SM_THREADS) {
            const int k = p / (SM_TS / 2), c = (p - k * (SM_TS / 2)) * 2;
            const int s = s0 + c;
            const int bytes = (s + 1 < S) ? 16 : ((s < S) ? 8 : 0);       // zero fill beyond S
            cp_async16(dst + k * SM_LD + c, Zt + (size_t)(kc0 + k) * S + (s < S ? s : 0), bytes);
        }
        cp_async_commit();
    };
    prefetch(0);
    const double y_s = ys[b];
    double mu[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int j = j0 + wj + a * 8 + g;
        mu[a] = (j < n) ? mean[(size_t)b * n + j] : 0.0;
    }
    double acc[4][2][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 2; ++c) acc[a][c][0] = acc[a][c][1] = 0.0;
    for (int it = 0; it < total; ++it) {
        if (it + 1 < total) {
            prefetch(it + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();                              // stage it (and, the first time, the A tile) is visible to all
        const int st = it / n_ch, ch = it - st * n_ch;
        const int kc0 = ch * SP_KC, kc = min(SP_KC, rp - kc0);
        const double* zs = Zs + (size_t)(it & 1) * SP_KC * SM_LD;
        for (int k0 = 0; k0 < kc; k0 += 4) {
            double af[4], bf[2];
            const double* ap = As + (kc0 + k0 + t) * SM_LD + wj + g;
            const double* zp = zs + (k0 + t) * SM_LD + ws + g;
#pragma unroll
            for (int a = 0; a < 4; ++a) af[a] = ap[a * 8];
#pragma unroll
            for (int c = 0; c < 2; ++c) bf[c] = zp[c * 8];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = 0; c < 2; ++c) dmma_m8n8k4(acc[a][c][0], acc[a][c][1], af[a], bf[c]);
        }
        if (ch == n_ch - 1) {                         // the tile is complete: y_s (acc + mu), accumulators reset
            const int s0 = st * SM_TS;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int j = j0 + wj + a * 8 + g;
                double* row = Y + ((size_t)b * n + j) * S;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int s = s0 + ws + c * 8 + t * 2;
                    const double v0 = y_s * (acc[a][c][0] + mu[a]), v1 = y_s * (acc[a][c][1] + mu[a]);
                    acc[a][c][0] = acc[a][c][1] = 0.0;
                    if (j < n) {
                        if (s + 1 < S) {
                            *reinterpret_cast<double2*>(row + s) = make_double2(v0, v1);
                        } else if (s < S) {
                            row[s] = v0;
                        }
                    }
                }
            }
        }
        __syncthreads();                              // everybody is done with stage it before it is refilled
    }
}

}  // namespace gpet

using namespace gpet;

extern "C" int gpet_sample_f64(const double* Zt, const double* A, const double* mean, const double* ys, int B, int rp, int n,
                               int S, double* Y, void* stream) {
    GPET_REQUIRE(Zt && A && mean && ys && Y && B > 0 && n > 0 && S > 0, "gpet_sample_f64: bad argument");
    GPET_SUPPORTED(rp >= 4 && (rp % 4) == 0, "gpet_sample_f64: rp=%d must be a multiple of 4", rp);
    const size_t smem = 2 * (size_t)(rp < SM_KC ? rp : SM_KC) * SM_LD * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(sample_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("sample smem attribute: %s", cudaGetErrorString(e));
        return GPET_ERR_CUDA;
    }
    GPET_SUPPORTED(B <= 65535 && (n + SM_TJ - 1) / SM_TJ <= 65535, "gpet_sample_f64: grid too large");
    // persistent-row kernel: needs 16-byte aligned sample pairs (S even, aligned buffers) and its A tile in shared memory
    const size_t smem_r = ((size_t)rp + 2 * SP_KC) * SM_LD * sizeof(double);
    if (g_tune[GPET_TUNE_SAMPLE_ROWS] && (S % 2) == 0 && ((uintptr_t)Zt % 16) == 0 && ((uintptr_t)Y % 16) == 0 &&
        smem_r <= 110 * 1024) {
        e = cudaFuncSetAttribute(sample_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r);
        if (e != cudaSuccess) {
            set_error("sample smem attribute: %s", cudaGetErrorString(e));
            return GPET_ERR_CUDA;
        }
        dim3 grid_r((n + SM_TJ - 1) / SM_TJ, B);
        sample_rows_kernel<<<grid_r, SM_THREADS, smem_r, (cudaStream_t)stream>>>(Zt, A, mean, ys, rp, n, S, Y);
        return check_launch("sample_rows_kernel");
    }
    dim3 grid((S + SM_TS - 1) / SM_TS, (n + SM_TJ - 1) / SM_TJ, B);
    sample_dmma_kernel<<<grid, SM_THREADS, smem, (cudaStream_t)stream>>>(Zt, A, mean, ys, rp, n, S, Y);
    return check_launch("sample_dmma_kernel");
}
