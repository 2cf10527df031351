// GP posterior of one iteration (non-converged branch), one CTA per trace, everything in shared memory.
//
// Reference seams: gpet.py:209-230 (training set scaling), sklearn_gpr.py:221-227 (centre, keep std),
// :304-320 (K, Cholesky, alpha), :381-385 (mean, multiplied by the kept std), :392-407 (V, covariance),
// :672-677 (WeightedWhiteKernel: noise only on the training diagonal, dropped when m == edge_length).
//
// Low-rank form: with k** = U diag(lam) U^T on the integer grid and the training x on that grid,
// K* = c k**[:, I], so  Sigma = sy^2 U (c lam - c^2 lam G^T G lam) U^T,  G = L^-1 U[I, :].
// The kernel emits the reduced rp x rp matrix; gpet_factor.cu diagonalises it.
#include "gpet_common.cuh"
#include "gpet_npsum.cuh"

namespace gpet {

constexpr int PT = 256;  // threads per CTA

struct PostScalars {
    double c, sy, ybar, ys;
};

// Storage of the lower triangle of K / L in shared memory: full rows with an odd leading dimension (fast index, the
// default) or packed rows (half the memory: training sets up to GPET_MAX_TRAIN points).
struct FullLowerP {
    int ld;
    __device__ __forceinline__ int operator()(int i, int j) const { return i * ld + j; }
};
struct PackedLowerP {
    __device__ __forceinline__ int operator()(int i, int j) const { return ((i * (i + 1)) >> 1) + j; }
};

// Steps shared by both posterior kernels: scaling, K, Cholesky, alpha, mean.
// smem: Ls[m*ldL], yv[mp], al[mp], tmp[mp], xs[mp] (int).  Returns false when the Cholesky fails.
template <class IX>
__device__ bool posterior_core(const int32_t* __restrict__ xi, const double* __restrict__ y, const double* __restrict__ w,
                               int m, int n, double sigma_f, double noise_y, double gp_alpha,
                               const double* __restrict__ kd, double* Ls, const IX ix, double* yv, double* al, double* tmp,
                               int* xs, PostScalars* sc, int* flag, double* __restrict__ mean_out) {
    const int tid = threadIdx.x;
    for (int i = tid; i < m; i += PT) {
        xs[i] = xi[i];
        yv[i] = y[i];
    }
    if (tid == 0) *flag = 0;
    __syncthreads();
    if (tid == 0) {
        // gpet.py:228-230: y_s = np.std(y) + 1; y /= y_s; constant = sigma_f**2 / y_s**2
        double mu = np_pairwise_sum(yv, m) / (double)m;
        for (int i = 0; i < m; ++i) { double d = yv[i] - mu; tmp[i] = d * d; }
        double ys = sqrt(np_pairwise_sum(tmp, m) / (double)m) + 1.0;
        for (int i = 0; i < m; ++i) yv[i] = yv[i] / ys;
        double c = (sigma_f * sigma_f) / (ys * ys);
        // sklearn_gpr.py:221-227: mean removed, std kept (1.0 when ~0)
        double ybar = np_pairwise_sum(yv, m) / (double)m;
        for (int i = 0; i < m; ++i) { double d = yv[i] - ybar; tmp[i] = d * d; }
        double sy = sqrt(np_pairwise_sum(tmp, m) / (double)m);
        if (sy < 10.0 * 2.220446049250313e-16) sy = 1.0;
        for (int i = 0; i < m; ++i) yv[i] = yv[i] - ybar;
        sc->c = c; sc->sy = sy; sc->ybar = ybar; sc->ys = ys;
    }
    __syncthreads();
    const double c = sc->c;
    const bool add_noise = (m != n);  // sklearn_gpr.py:672-677 quirk
    for (int p = tid; p < m * m; p += PT) {
        int i = p / m, j = p - i * m;
        if (j > i) continue;
        int d = xs[i] - xs[j];
        d = d < 0 ? -d : d;
        double v = c * kd[d];
        if (i == j) {
            if (add_noise) v = v + noise_y * w[i];
            v = v + gp_alpha;
        }
        Ls[ix(i, j)] = v;
    }
    __syncthreads();
    // right-looking Cholesky (lower)
    for (int k = 0; k < m; ++k) {
        if (tid == 0) {
            double dkk = Ls[ix(k, k)];
            if (!(dkk > 0.0)) { *flag = 1; dkk = 1.0; }
            Ls[ix(k, k)] = sqrt(dkk);
        }
        __syncthreads();
        const double inv = 1.0 / Ls[ix(k, k)];
        for (int i = k + 1 + tid; i < m; i += PT) Ls[ix(i, k)] *= inv;
        __syncthreads();
        const int rem = m - k - 1;
        for (int p = tid; p < rem * rem; p += PT) {
            int ii = p / rem, jj = p - ii * rem;
            if (jj > ii) continue;
            int i = k + 1 + ii, j = k + 1 + jj;
            Ls[ix(i, j)] = fma(-Ls[ix(i, k)], Ls[ix(j, k)], Ls[ix(i, j)]);
        }
        __syncthreads();
    }
    // alpha = L^-T L^-1 y  (warp 0)
    if (tid < 32) {
        for (int i = 0; i < m; ++i) {
            double s = 0.0;
            for (int k = tid; k < i; k += 32) s = fma(Ls[ix(i, k)], tmp[k], s);
            s = warp_sum(s);
            if (tid == 0) tmp[i] = (yv[i] - s) / Ls[ix(i, i)];
            __syncwarp();
        }
        for (int i = m - 1; i >= 0; --i) {
            double s = 0.0;
            for (int k = i + 1 + tid; k < m; k += 32) s = fma(Ls[ix(k, i)], al[k], s);
            s = warp_sum(s);
            if (tid == 0) al[i] = (tmp[i] - s) / Ls[ix(i, i)];
            __syncwarp();
        }
    }
    __syncthreads();
    // posterior mean on the grid: sy * (K* alpha) + ybar  (sklearn_gpr.py:381-385)
    const double sy = sc->sy, ybar = sc->ybar;
    for (int j = tid; j < n; j += PT) {
        double s = 0.0;
        for (int i = 0; i < m; ++i) {
            int d = j - xs[i];
            d = d < 0 ? -d : d;
            s = fma(c * kd[d], al[i], s);
        }
        mean_out[j] = sy * s + ybar;
    }
    return *flag == 0;
}

__global__ void __launch_bounds__(PT)
posterior_lowrank_kernel(const int32_t* __restrict__ xi, const double* __restrict__ y, const double* __restrict__ w,
                         const int32_t* __restrict__ m_arr, int mmax, int n, const double* __restrict__ sigma_f,
                         double noise_y, double gp_alpha, const double* __restrict__ kd,
                         const double* __restrict__ Ur, const double* __restrict__ lam, int rp,
                         double* __restrict__ mean, double* __restrict__ ys_out, double* __restrict__ Mr,
                         int32_t* __restrict__ status) {
    extern __shared__ double sm[];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int m = m_arr[b];
    const int ldL = m | 1;
    double* Ls = sm;
    double* Gs = Ls + (size_t)mmax * (mmax | 1);
    double* yv = Gs + (size_t)mmax * rp;
    double* al = yv + mmax;
    double* tmp = al + mmax;
    int* xs = (int*)(tmp + mmax);
    __shared__ PostScalars sc;
    __shared__ int flag;
    bool ok = posterior_core(xi + (size_t)b * mmax, y + (size_t)b * mmax, w + (size_t)b * mmax, m, n, sigma_f[b], noise_y,
                             gp_alpha, kd, Ls, FullLowerP{ldL}, yv, al, tmp, xs, &sc, &flag, mean + (size_t)b * n);
    if (tid == 0) {
        ys_out[b] = sc.ys;
        status[b] = ok ? 0 : 1;
    }
    // G = L^-1 U[I, :]   (m x rp), right-looking forward substitution
    for (int p = tid; p < m * rp; p += PT) {
        int i = p / rp, k = p - i * rp;
        Gs[p] = Ur[(size_t)xs[i] * rp + k];
    }
    __syncthreads();
    for (int k = 0; k < m; ++k) {
        const double inv = 1.0 / Ls[k * ldL + k];
        for (int c2 = tid; c2 < rp; c2 += PT) Gs[k * rp + c2] *= inv;
        __syncthreads();
        const int rem = m - k - 1;
        for (int p = tid; p < rem * rp; p += PT) {
            int ii = p / rp, c2 = p - ii * rp;
            int i = k + 1 + ii;
            Gs[i * rp + c2] = fma(-Ls[i * ldL + k], Gs[k * rp + c2], Gs[i * rp + c2]);
        }
        __syncthreads();
    }
    // Mr = sy^2 (c lam - c^2 lam (G^T G) lam)
    const double c = sc.c, sy2 = sc.sy * sc.sy;
    double* out = Mr + (size_t)b * rp * rp;
    for (int p = tid; p < rp * rp; p += PT) {
        int a = p / rp, bb = p - a * rp;
        double t = 0.0;
        for (int i = 0; i < m; ++i) t = fma(Gs[i * rp + a], Gs[i * rp + bb], t);
        double v = -(c * c) * (lam[a] * t * lam[bb]);
        if (a == bb) v += c * lam[a];
        out[p] = sy2 * v;
    }
}

// ---- training sets beyond the all-in-shared-memory kernels (m up to GPET_MAX_TRAIN, e.g. BASELINE config 4) -------------
// One CTA per trace.  K / L live in shared memory as a PACKED lower triangle (m = 224: 197 KB); the right-hand sides of
// the triangular solve do not fit beside it, so the solve runs one COLUMN PER THREAD straight out of global memory:
// row blocks of 16 held in registers, the already solved rows re-read (coalesced across the threads of a warp, L1/L2
// resident), L broadcast from shared memory, no barrier at all.  Per element the operations and their order are those of
// the shared-memory kernels (right-looking there, left-looking here: the same fma chain), so both paths give the same bits.
//   full == 0:  G = L^-1 U_r[I, :]      (m x rp)   -> gram_lowrank_kernel forms M_r
//   full != 0:  V = L^-1 K*^T           (m x n)    -> posterior_full_phase2_kernel forms Sigma
constexpr int RB = 16;
__global__ void __launch_bounds__(PT)
posterior_packed_kernel(const int32_t* __restrict__ xi, const double* __restrict__ y, const double* __restrict__ w,
                        const int32_t* __restrict__ m_arr, int mmax, int n, const double* __restrict__ sigma_f,
                        double noise_y, double gp_alpha, const double* __restrict__ kd, const double* __restrict__ Ur,
                        int rp, int full, double* __restrict__ mean, double* __restrict__ ys_out,
                        double* __restrict__ scal, double* __restrict__ G, int32_t* __restrict__ status) {
    extern __shared__ double sm[];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int m = m_arr[b];
    const PackedLowerP ix{};
    double* Ls = sm;
    double* yv = Ls + ((size_t)mmax * (mmax + 1)) / 2;
    double* al = yv + mmax;
    double* tmp = al + mmax;
    int* xs = (int*)(tmp + mmax);
    __shared__ PostScalars sc;
    __shared__ int flag;
    bool ok = posterior_core(xi + (size_t)b * mmax, y + (size_t)b * mmax, w + (size_t)b * mmax, m, n, sigma_f[b], noise_y,
                             gp_alpha, kd, Ls, ix, yv, al, tmp, xs, &sc, &flag, mean + (size_t)b * n);
    if (tid == 0) {
        ys_out[b] = sc.ys;
        status[b] = ok ? 0 : 1;
        scal[2 * b] = sc.c;
        scal[2 * b + 1] = sc.sy;
    }
    __syncthreads();
    const double c = sc.c;
    const int ncol = full ? n : rp;
    double* Gb = G + (size_t)b * mmax * ncol;
    for (int col = tid; col < ncol; col += PT) {
        for (int i0 = 0; i0 < m; i0 += RB) {
            const int ib = min(RB, m - i0);
            double acc[RB];
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                acc[r] = 0.0;
                if (r < ib) {
                    const int xr = xs[i0 + r];
                    if (full) {
                        int d = col - xr;
                        d = d < 0 ? -d : d;
                        acc[r] = c * kd[d];
                    } else {
                        acc[r] = Ur[(size_t)xr * rp + col];
                    }
                }
            }
            for (int k = 0; k < i0; ++k) {
                const double gk = Gb[(size_t)k * ncol + col];
#pragma unroll
                for (int r = 0; r < RB; ++r)
                    if (r < ib) acc[r] = fma(-Ls[ix(i0 + r, k)], gk, acc[r]);
            }
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                if (r < ib) {
                    acc[r] *= 1.0 / Ls[ix(i0 + r, i0 + r)];
#pragma unroll
                    for (int r2 = r + 1; r2 < RB; ++r2)
                        if (r2 < ib) acc[r2] = fma(-Ls[ix(i0 + r2, i0 + r)], acc[r], acc[r2]);
                }
            }
#pragma unroll
            for (int r = 0; r < RB; ++r)
                if (r < ib) Gb[(size_t)(i0 + r) * ncol + col] = acc[r];
        }
    }
}

// M_r = sy^2 (c lam - c^2 lam (G^T G) lam): 64 x 64 tile per CTA, 4 x 4 per thread, k ascending like the in-kernel form
__global__ void __launch_bounds__(256)
gram_lowrank_kernel(const double* __restrict__ G, const int32_t* __restrict__ m_arr, int mmax, int rp,
                    const double* __restrict__ lam, const double* __restrict__ scal, double* __restrict__ Mr) {
    constexpr int GT = 64, GK = 16;
    __shared__ double As[GK][GT + 1], Bs[GK][GT + 1];
    const int b = blockIdx.z, m = m_arr[b];
    const int i0 = blockIdx.y * GT, j0 = blockIdx.x * GT;
    const double* Gb = G + (size_t)b * mmax * rp;
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    double acc[4][4] = {};
    for (int k0 = 0; k0 < m; k0 += GK) {
        for (int p = threadIdx.x; p < GK * GT; p += 256) {
            const int k = p / GT, cc = p - k * GT;
            double va = 0.0, vb = 0.0;
            if (k0 + k < m) {
                if (i0 + cc < rp) va = Gb[(size_t)(k0 + k) * rp + i0 + cc];
                if (j0 + cc < rp) vb = Gb[(size_t)(k0 + k) * rp + j0 + cc];
            }
            As[k][cc] = va;
            Bs[k][cc] = vb;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GK; ++k) {
            double a[4], bb[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) { a[r] = As[k][ty * 4 + r]; bb[r] = Bs[k][tx * 4 + r]; }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[r][q] = fma(a[r], bb[q], acc[r][q]);
        }
        __syncthreads();
    }
    const double c = scal[2 * b], sy2 = scal[2 * b + 1] * scal[2 * b + 1];
    double* out = Mr + (size_t)b * rp * rp;
    for (int r = 0; r < 4; ++r)
        for (int q = 0; q < 4; ++q) {
            const int a = i0 + ty * 4 + r, bb = j0 + tx * 4 + q;
            if (a < rp && bb < rp) {
                double v = -(c * c) * (lam[a] * acc[r][q] * lam[bb]);
                if (a == bb) v += c * lam[a];
                out[(size_t)a * rp + bb] = sy2 * v;
            }
        }
}

static size_t packed_smem_bytes(int mmax) {
    return (((size_t)mmax * (mmax + 1)) / 2 + 3 * (size_t)mmax) * sizeof(double) + ((size_t)mmax + 2) * sizeof(int);
}

// ---- full covariance path -------------------------------------------------------------------------
constexpr int VC = 64;  // grid columns per V chunk

__global__ void __launch_bounds__(PT)
posterior_full_phase1_kernel(const int32_t* __restrict__ xi, const double* __restrict__ y, const double* __restrict__ w,
                             const int32_t* __restrict__ m_arr, int mmax, int n, const double* __restrict__ sigma_f,
                             double noise_y, double gp_alpha, const double* __restrict__ kd,
                             double* __restrict__ mean, double* __restrict__ ys_out, double* __restrict__ V,
                             double* __restrict__ scal, int32_t* __restrict__ status) {
    extern __shared__ double sm[];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int m = m_arr[b];
    const int ldL = m | 1;
    double* Ls = sm;
    double* Vs = Ls + (size_t)mmax * (mmax | 1);  // m x VC
    double* yv = Vs + (size_t)mmax * VC;
    double* al = yv + mmax;
    double* tmp = al + mmax;
    int* xs = (int*)(tmp + mmax);
    __shared__ PostScalars sc;
    __shared__ int flag;
    bool ok = posterior_core(xi + (size_t)b * mmax, y + (size_t)b * mmax, w + (size_t)b * mmax, m, n, sigma_f[b], noise_y,
                             gp_alpha, kd, Ls, FullLowerP{ldL}, yv, al, tmp, xs, &sc, &flag, mean + (size_t)b * n);
    if (tid == 0) {
        ys_out[b] = sc.ys;
        status[b] = ok ? 0 : 1;
        scal[2 * b] = sc.c;
        scal[2 * b + 1] = sc.sy;
    }
    const double c = sc.c;
    double* Vb = V + (size_t)b * mmax * n;
    for (int j0 = 0; j0 < n; j0 += VC) {
        const int nc = min(VC, n - j0);
        __syncthreads();
        for (int p = tid; p < m * VC; p += PT) {
            int i = p / VC, jj = p - i * VC;
            double v = 0.0;
            if (jj < nc) {
                int d = (j0 + jj) - xs[i];
                d = d < 0 ? -d : d;
                v = c * kd[d];
            }
            Vs[p] = v;
        }
        __syncthreads();
        for (int k = 0; k < m; ++k) {
            const double inv = 1.0 / Ls[k * ldL + k];
            for (int c2 = tid; c2 < VC; c2 += PT) Vs[k * VC + c2] *= inv;
            __syncthreads();
            const int rem = m - k - 1;
            for (int p = tid; p < rem * VC; p += PT) {
                int ii = p / VC, c2 = p - ii * VC;
                int i = k + 1 + ii;
                Vs[i * VC + c2] = fma(-Ls[i * ldL + k], Vs[k * VC + c2], Vs[i * VC + c2]);
            }
            __syncthreads();
        }
        for (int p = tid; p < m * VC; p += PT) {
            int i = p / VC, jj = p - i * VC;
            if (jj < nc) Vb[(size_t)i * n + j0 + jj] = Vs[p];
        }
    }
}

// cov tile = sy^2 (c kd[|i-j|] - V_i . V_j), 64x64 tile per CTA, 4x4 per thread
constexpr int CT = 64, CK = 16;
__global__ void __launch_bounds__(256)
posterior_full_phase2_kernel(const double* __restrict__ V, const int32_t* __restrict__ m_arr, int mmax, int n,
                             const double* __restrict__ kd, const double* __restrict__ scal, double* __restrict__ cov) {
    __shared__ double As[CK][CT + 1], Bs[CK][CT + 1];
    const int b = blockIdx.z, m = m_arr[b];
    const int i0 = blockIdx.y * CT, j0 = blockIdx.x * CT;
    const double* Vb = V + (size_t)b * mmax * n;
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    double acc[4][4] = {};
    for (int k0 = 0; k0 < m; k0 += CK) {
        for (int p = threadIdx.x; p < CK * CT; p += 256) {
            int k = p / CT, cc = p - k * CT;
            double va = 0.0, vb = 0.0;
            if (k0 + k < m) {
                if (i0 + cc < n) va = Vb[(size_t)(k0 + k) * n + i0 + cc];
                if (j0 + cc < n) vb = Vb[(size_t)(k0 + k) * n + j0 + cc];
            }
            As[k][cc] = va;
            Bs[k][cc] = vb;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < CK; ++k) {
            double a[4], bb[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) { a[r] = As[k][ty * 4 + r]; bb[r] = Bs[k][tx * 4 + r]; }
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int s = 0; s < 4; ++s) acc[r][s] = fma(a[r], bb[s], acc[r][s]);
        }
        __syncthreads();
    }
    const double c = scal[2 * b], sy2 = scal[2 * b + 1] * scal[2 * b + 1];
    double* out = cov + (size_t)b * n * n;
    for (int r = 0; r < 4; ++r)
        for (int s = 0; s < 4; ++s) {
            int i = i0 + ty * 4 + r, j = j0 + tx * 4 + s;
            if (i < n && j < n) {
                int d = i - j;
                d = d < 0 ? -d : d;
                out[(size_t)i * n + j] = (c * kd[d] - acc[r][s]) * sy2;
            }
        }
}

static size_t core_smem_doubles(int mmax) { return (size_t)mmax * (mmax | 1) + 3 * (size_t)mmax + (size_t)(mmax + 1) / 2 + 2; }

}  // namespace gpet

using namespace gpet;

// the all-in-shared-memory kernels take a batch when its sizes fit them (and GPET_TUNE_POSTERIOR_PACKED is off)
static bool small_path(int mmax, int extra_cols) {
    if (g_tune[GPET_TUNE_POSTERIOR_PACKED]) return false;
    return (core_smem_doubles(mmax) + (size_t)mmax * extra_cols) * sizeof(double) <= 227 * 1024;
}

static int launch_packed(const int32_t* xi, const double* y, const double* w, const int32_t* m, int mmax, int B, int n,
                         const double* sigma_f, double noise_y, double gp_alpha, const double* kd, const double* Ur, int rp,
                         int full, double* mean, double* ys, double* scal, double* G, int32_t* status, cudaStream_t st) {
    const size_t smem = packed_smem_bytes(mmax);
    GPET_SUPPORTED(smem <= 227 * 1024, "posterior: %d training points need %zu B shared memory", mmax, smem);
    cudaError_t e = cudaFuncSetAttribute(posterior_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("posterior_packed smem attribute: %s", cudaGetErrorString(e));
        return GPET_ERR_CUDA;
    }
    posterior_packed_kernel<<<B, PT, smem, st>>>(xi, y, w, m, mmax, n, sigma_f, noise_y, gp_alpha, kd, Ur, rp, full, mean, ys,
                                                scal, G, status);
    return check_launch("posterior_packed_kernel");
}

extern "C" int64_t gpet_posterior_lowrank_workspace_bytes(int B, int mmax, int rp) {
    if (small_path(mmax, rp)) return 0;
    return (int64_t)B * mmax * rp * 8 + (int64_t)B * 2 * 8 + 256;
}

extern "C" int gpet_posterior_lowrank_f64(const int32_t* xi, const double* y, const double* w, const int32_t* m, int mmax,
                                          int B, int n, const double* sigma_f, double noise_y, double gp_alpha,
                                          const double* kd, const double* Ur, const double* lam, int rp, double* mean,
                                          double* ys, double* Mr, int32_t* status, void* work, void* stream) {
    GPET_REQUIRE(xi && y && w && m && sigma_f && kd && Ur && lam && mean && ys && Mr && status,
                 "gpet_posterior_lowrank_f64: null pointer");
    GPET_REQUIRE(B > 0 && n > 1 && mmax >= 2 && rp > 0, "gpet_posterior_lowrank_f64: bad shape");
    GPET_SUPPORTED(mmax <= GPET_MAX_TRAIN && rp <= GPET_MAX_RANK,
                   "gpet_posterior_lowrank_f64: mmax=%d (max %d) rp=%d (max %d)", mmax, GPET_MAX_TRAIN, rp, GPET_MAX_RANK);
    if (!small_path(mmax, rp)) {
        GPET_REQUIRE(work != nullptr, "gpet_posterior_lowrank_f64: workspace required (gpet_posterior_lowrank_workspace_bytes)");
        cudaStream_t st = (cudaStream_t)stream;
        double* G = (double*)work;
        double* scal = G + (size_t)B * mmax * rp;
        int rc = launch_packed(xi, y, w, m, mmax, B, n, sigma_f, noise_y, gp_alpha, kd, Ur, rp, 0, mean, ys, scal, G, status, st);
        if (rc) return rc;
        dim3 grid((rp + 63) / 64, (rp + 63) / 64, B);
        gram_lowrank_kernel<<<grid, 256, 0, st>>>(G, m, mmax, rp, lam, scal, Mr);
        return check_launch("gram_lowrank_kernel");
    }
    const size_t smem = (core_smem_doubles(mmax) + (size_t)mmax * rp) * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(posterior_lowrank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("posterior_lowrank smem attribute: %s", cudaGetErrorString(e));
        return GPET_ERR_CUDA;
    }
    posterior_lowrank_kernel<<<B, PT, smem, (cudaStream_t)stream>>>(xi, y, w, m, mmax, n, sigma_f, noise_y, gp_alpha, kd, Ur,
                                                                   lam, rp, mean, ys, Mr, status);
    return check_launch("posterior_lowrank_kernel");
}

extern "C" int64_t gpet_posterior_full_workspace_bytes(int B, int mmax, int n) {
    return (int64_t)B * mmax * n * 8 + (int64_t)B * 2 * 8 + 256;
}

extern "C" int gpet_posterior_full_f64(const int32_t* xi, const double* y, const double* w, const int32_t* m, int mmax, int B,
                                       int n, const double* sigma_f, double noise_y, double gp_alpha, const double* kd,
                                       double* mean, double* ys, double* cov, int32_t* status, void* work, void* stream) {
    GPET_REQUIRE(xi && y && w && m && sigma_f && kd && mean && ys && cov && status && work,
                 "gpet_posterior_full_f64: null pointer");
    GPET_REQUIRE(B > 0 && n > 1 && mmax >= 2, "gpet_posterior_full_f64: bad shape");
    GPET_SUPPORTED(mmax <= GPET_MAX_TRAIN, "gpet_posterior_full_f64: mmax=%d (max %d)", mmax, GPET_MAX_TRAIN);
    cudaStream_t st = (cudaStream_t)stream;
    double* V = (double*)work;
    double* scal = V + (size_t)B * mmax * n;
    int rc;
    if (small_path(mmax, VC)) {
        const size_t smem = (core_smem_doubles(mmax) + (size_t)mmax * VC) * sizeof(double);
        cudaError_t e = cudaFuncSetAttribute(posterior_full_phase1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("posterior_full smem attribute: %s", cudaGetErrorString(e));
            return GPET_ERR_CUDA;
        }
        posterior_full_phase1_kernel<<<B, PT, smem, st>>>(xi, y, w, m, mmax, n, sigma_f, noise_y, gp_alpha, kd, mean, ys, V,
                                                         scal, status);
        rc = check_launch("posterior_full_phase1_kernel");
    } else {
        rc = launch_packed(xi, y, w, m, mmax, B, n, sigma_f, noise_y, gp_alpha, kd, nullptr, 0, 1, mean, ys, scal, V, status, st);
    }
    if (rc) return rc;
    dim3 grid((n + CT - 1) / CT, (n + CT - 1) / CT, B);
    posterior_full_phase2_kernel<<<grid, 256, 0, st>>>(V, m, mmax, n, kd, scal, cov);
    return check_launch("posterior_full_phase2_kernel");
}
