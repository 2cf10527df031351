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

// Steps shared by both posterior kernels: scaling, K, Cholesky, alpha, mean.
// smem: Ls[m*ldL], yv[mp], al[mp], tmp[mp], xs[mp] (int).  Returns false when the Cholesky fails.
__device__ bool posterior_core(const int32_t* __restrict__ xi, const double* __restrict__ y, const double* __restrict__ w,
                               int m, int n, double sigma_f, double noise_y, double gp_alpha,
                               const double* __restrict__ kd, double* Ls, int ldL, double* yv, double* al, double* tmp,
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
        Ls[i * ldL + j] = v;
    }
    __syncthreads();
    // right-looking Cholesky (lower)
    for (int k = 0; k < m; ++k) {
        if (tid == 0) {
            double dkk = Ls[k * ldL + k];
            if (!(dkk > 0.0)) { *flag = 1; dkk = 1.0; }
            Ls[k * ldL + k] = sqrt(dkk);
        }
        __syncthreads();
        const double inv = 1.0 / Ls[k * ldL + k];
        for (int i = k + 1 + tid; i < m; i += PT) Ls[i * ldL + k] *= inv;
        __syncthreads();
        const int rem = m - k - 1;
        for (int p = tid; p < rem * rem; p += PT) {
            int ii = p / rem, jj = p - ii * rem;
            if (jj > ii) continue;
            int i = k + 1 + ii, j = k + 1 + jj;
            Ls[i * ldL + j] = fma(-Ls[i * ldL + k], Ls[j * ldL + k], Ls[i * ldL + j]);
        }
        __syncthreads();
    }
    // alpha = L^-T L^-1 y  (warp 0)
    if (tid < 32) {
        for (int i = 0; i < m; ++i) {
            double s = 0.0;
            for (int k = tid; k < i; k += 32) s = fma(Ls[i * ldL + k], tmp[k], s);
            s = warp_sum(s);
            if (tid == 0) tmp[i] = (yv[i] - s) / Ls[i * ldL + i];
            __syncwarp();
        }
        for (int i = m - 1; i >= 0; --i) {
            double s = 0.0;
            for (int k = i + 1 + tid; k < m; k += 32) s = fma(Ls[k * ldL + i], al[k], s);
            s = warp_sum(s);
            if (tid == 0) al[i] = (tmp[i] - s) / Ls[i * ldL + i];
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
                             gp_alpha, kd, Ls, ldL, yv, al, tmp, xs, &sc, &flag, mean + (size_t)b * n);
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
                             gp_alpha, kd, Ls, ldL, yv, al, tmp, xs, &sc, &flag, mean + (size_t)b * n);
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

extern "C" int gpet_posterior_lowrank_f64(const int32_t* xi, const double* y, const double* w, const int32_t* m, int mmax,
                                          int B, int n, const double* sigma_f, double noise_y, double gp_alpha,
                                          const double* kd, const double* Ur, const double* lam, int rp, double* mean,
                                          double* ys, double* Mr, int32_t* status, void* stream) {
    GPET_REQUIRE(xi && y && w && m && sigma_f && kd && Ur && lam && mean && ys && Mr && status,
                 "gpet_posterior_lowrank_f64: null pointer");
    GPET_REQUIRE(B > 0 && n > 1 && mmax >= 2 && rp > 0, "gpet_posterior_lowrank_f64: bad shape");
    GPET_SUPPORTED(mmax <= GPET_MAX_TRAIN && rp <= GPET_MAX_RANK,
                   "gpet_posterior_lowrank_f64: mmax=%d (max %d) rp=%d (max %d)", mmax, GPET_MAX_TRAIN, rp, GPET_MAX_RANK);
    const size_t smem = (core_smem_doubles(mmax) + (size_t)mmax * rp) * sizeof(double);
    GPET_SUPPORTED(smem <= 227 * 1024, "gpet_posterior_lowrank_f64: needs %zu B shared memory (mmax=%d, rp=%d)", smem, mmax, rp);
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
    const size_t smem = (core_smem_doubles(mmax) + (size_t)mmax * VC) * sizeof(double);
    GPET_SUPPORTED(smem <= 227 * 1024, "gpet_posterior_full_f64: needs %zu B shared memory", smem);
    cudaError_t e = cudaFuncSetAttribute(posterior_full_phase1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("posterior_full smem attribute: %s", cudaGetErrorString(e));
        return GPET_ERR_CUDA;
    }
    cudaStream_t st = (cudaStream_t)stream;
    double* V = (double*)work;
    double* scal = V + (size_t)B * mmax * n;
    posterior_full_phase1_kernel<<<B, PT, smem, st>>>(xi, y, w, m, mmax, n, sigma_f, noise_y, gp_alpha, kd, mean, ys, V, scal,
                                                     status);
    int rc = check_launch("posterior_full_phase1_kernel");
    if (rc) return rc;
    dim3 grid((n + CT - 1) / CT, (n + CT - 1) / CT, B);
    posterior_full_phase2_kernel<<<grid, 256, 0, st>>>(V, m, mmax, n, kd, scal, cov);
    return check_launch("posterior_full_phase2_kernel");
}
