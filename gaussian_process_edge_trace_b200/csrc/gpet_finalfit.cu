// Final hyper-parameter fit (converged branch of fit_predict_GP): batched objective for the host L-BFGS-B
// drivers, and the final predictive mean / std.
//
// Reference seams: sklearn_gpr.py:475-585 (log_marginal_likelihood + analytic gradient), :257-262 (obj_func =
// -LML, -grad), sklearn kernels Constant*RBF|Matern + WeightedWhiteKernel (sklearn_gpr.py:647-694) with
// theta = log[constant, length_scale, noise_level]; sklearn_gpr.py:379-436 (predict with return_std) and
// gpet.py:263-266.
//
// One CTA per evaluation.  Shared memory holds ONE m x (m+1) matrix: lower triangle = K -> L -> L^-1 in place,
// strict upper triangle + the extra column = K^-1 (symmetric), so m = 160 still fits in 227 KB.
#include "gpet_common.cuh"

namespace gpet {

#define FF_THREADS ((int)blockDim.x)
#define FF_WARPS ((int)(blockDim.x >> 5))
constexpr int FF_MAX_THREADS = 1024;

// kind: 0 RBF, 1 Matern nu=0.5, 2 Matern nu=1.5, 3 Matern nu=2.5.  D = squared scaled distance.
__device__ __forceinline__ double kern_val(int kind, double D) {
    if (kind == 0) return exp(-0.5 * D);
    const double d = sqrt(D);
    if (kind == 1) return exp(-d);
    if (kind == 2) { const double t = d * 1.7320508075688772; return (1.0 + t) * exp(-t); }
    const double t = d * 2.23606797749979;
    return (1.0 + t + t * t / 3.0) * exp(-t);
}
// d k / d log(length_scale) (sklearn kernels.py RBF/Matern eval_gradient)
__device__ __forceinline__ double kern_dlogl(int kind, double D) {
    if (kind == 0) return exp(-0.5 * D) * D;
    if (kind == 1) { const double d = sqrt(D); return d > 0.0 ? exp(-d) * d : 0.0; }
    if (kind == 2) return 3.0 * D * exp(-sqrt(3.0 * D));
    const double t = sqrt(5.0 * D);
    return 5.0 / 3.0 * D * (t + 1.0) * exp(-t);
}

// value and d/dlog(length_scale) with one exponential
__device__ __forceinline__ void kern_both(int kind, double D, double& k, double& dk) {
    if (kind == 0) { k = exp(-0.5 * D); dk = k * D; return; }
    const double d = sqrt(D);
    if (kind == 1) { k = exp(-d); dk = k * d; return; }
    if (kind == 2) { const double t = d * 1.7320508075688772; const double e = exp(-t); k = (1.0 + t) * e; dk = 3.0 * D * e; return; }
    const double t = d * 2.23606797749979;
    const double e = exp(-t);
    k = (1.0 + t + t * t / 3.0) * e;
    dk = 5.0 / 3.0 * D * (t + 1.0) * e;
}

// K (lower triangle of Ms) = c k(X/l) + diag(noise w + alpha)
__device__ void build_kernel_matrix(int kind, int m, int ld, double c, double ls, double noise, double gp_alpha,
                                    const double* __restrict__ X, const double* __restrict__ y,
                                    const double* __restrict__ w, double* Ms, double* xs, double* yv) {
    const int tid = threadIdx.x;
    for (int i = tid; i < m; i += FF_THREADS) {
        xs[i] = X[i] / ls;
        yv[i] = y[i];
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    for (int i = warp; i < m; i += FF_WARPS) {       // one warp per row, lanes across the columns j <= i
        const double xi = xs[i];
        for (int j = lane; j <= i; j += 32) {
            double v;
            if (i == j) {
                v = (c + noise * w[i]) + gp_alpha;
            } else {
                const double d = xi - xs[j];
                v = c * kern_val(kind, d * d);
            }
            Ms[i * ld + j] = v;
        }
    }
    __syncthreads();
}

// In-place lower Cholesky, two barriers per column, triangular work mapping.  Returns false (uniformly) on a
// non-positive pivot.
__device__ bool cholesky_inplace(int m, int ld, double* Ms) {
    const int tid = threadIdx.x;
    bool ok = true;
    for (int k = 0; k < m; ++k) {
        double dkk = Ms[k * ld + k];
        if (!(dkk > 0.0)) { ok = false; dkk = 1.0; }
        const double sq = sqrt(dkk), inv = 1.0 / sq;
        for (int i = k + 1 + tid; i < m; i += FF_THREADS) Ms[i * ld + k] *= inv;
        __syncthreads();
        if (tid == 0) Ms[k * ld + k] = sq;      // nobody reads the pivot during the trailing update
        for (int i = k + 1 + (tid >> 5); i < m; i += FF_WARPS) {   // rank-1 update, one warp per row
            const double lik = -Ms[i * ld + k];
            for (int j = k + 1 + (tid & 31); j <= i; j += 32) Ms[i * ld + j] = fma(lik, Ms[j * ld + k], Ms[i * ld + j]);
        }
        __syncthreads();
    }
    return ok;
}

// In-place inverse of the lower-triangular factor (LAPACK dtrti2 column order); every row's dot product is split
// over 4 lanes and reduced with shuffles.
__device__ void tri_inverse_inplace(int m, int ld, double* Ms, double* tmp) {
    const int tid = threadIdx.x, quad = tid >> 2, l = tid & 3;
    for (int j = m - 1; j >= 0; --j) {
        const double ajj = 1.0 / Ms[j * ld + j];
        for (int i = j + 1 + tid; i < m; i += FF_THREADS) tmp[i] = Ms[i * ld + j];
        __syncthreads();
        const int rem = m - j - 1;
        for (int base = 0; base < rem; base += FF_THREADS / 4) {
            const int i = j + 1 + base + quad;
            double s = 0.0;
            if (i < m)
                for (int k = j + 1 + l; k <= i; k += 4) s = fma(Ms[i * ld + k], tmp[k], s);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (i < m && l == 0) Ms[i * ld + j] = -ajj * s;
        }
        if (tid == 0) Ms[j * ld + j] = ajj;
        __syncthreads();
    }
}

// alpha = K^-1 y = T^T (T y) with T = L^-1 (lower) stored in Ms
__device__ void alpha_from_inverse(int m, int ld, const double* Ms, const double* yv, double* tmp, double* al) {
    const int tid = threadIdx.x;
    for (int i = tid; i < m; i += FF_THREADS) {
        double s = 0.0;
        for (int k = 0; k <= i; ++k) s = fma(Ms[i * ld + k], yv[k], s);
        tmp[i] = s;
    }
    __syncthreads();
    for (int i = tid; i < m; i += FF_THREADS) {
        double s = 0.0;
        for (int k = i; k < m; ++k) s = fma(Ms[k * ld + i], tmp[k], s);
        al[i] = s;
    }
    __syncthreads();
}

__device__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < FF_WARPS; ++i) t += red[i];
    return t;
}

__global__ void __launch_bounds__(FF_MAX_THREADS)
lml_kernel(const double* __restrict__ X, const double* __restrict__ y, const double* __restrict__ w,
           const int32_t* __restrict__ m_arr, int mmax, const int32_t* __restrict__ trace_of,
           const double* __restrict__ theta, int kind, double gp_alpha, double* __restrict__ f_out,
           double* __restrict__ g_out) {
    extern __shared__ double sm[];
    __shared__ double red[FF_MAX_THREADS / 32];
    const int e = blockIdx.x, tid = threadIdx.x;
    const int tr = trace_of[e];
    const int m = m_arr[tr];
    const int ld = (m + 1) | 1;       // odd: conflict-free row-strided walks; column m holds diag(K^-1)
    double* Ms = sm;
    double* xs = Ms + (size_t)mmax * ((mmax + 1) | 1);
    double* yv = xs + mmax;
    double* al = yv + mmax;
    double* tmp = al + mmax;
    const double c = exp(theta[3 * e]), ls = exp(theta[3 * e + 1]), noise = exp(theta[3 * e + 2]);
    const double* wt = w + (size_t)tr * mmax;
    build_kernel_matrix(kind, m, ld, c, ls, noise, gp_alpha, X + (size_t)tr * mmax, y + (size_t)tr * mmax, wt, Ms, xs, yv);
    if (!cholesky_inplace(m, ld, Ms)) {  // sklearn_gpr.py:521-522: LML = -inf, gradient = 0
        if (tid == 0) {
            f_out[e] = __longlong_as_double(0x7ff0000000000000LL);
            g_out[3 * e] = g_out[3 * e + 1] = g_out[3 * e + 2] = 0.0;
        }
        return;
    }
    tri_inverse_inplace(m, ld, Ms, tmp);
    alpha_from_inverse(m, ld, Ms, yv, tmp, al);
    // -LML = 0.5 y^T alpha + sum log diag(L) + m/2 log(2 pi);  diag(L) = 1 / diag(L^-1)
    double part = 0.0;
    for (int i = tid; i < m; i += FF_THREADS) part += 0.5 * yv[i] * al[i] - log(Ms[i * ld + i]);
    const double nlml = block_sum(part, red) + 0.5 * (double)m * 1.8378770664093453;
    // K^-1 = T^T T: strict lower part -> strict upper triangle (transposed slot), diagonal -> column m
    const int warp = tid >> 5, lane = tid & 31;
    for (int i = warp; i < m; i += FF_WARPS) {
        for (int j = lane; j <= i; j += 32) {
            double s = 0.0;
            for (int k = i; k < m; ++k) s = fma(Ms[k * ld + i], Ms[k * ld + j], s);
            // the lower triangle (L^-1) is still being read: results go to the unused upper storage
            if (i == j) Ms[i * ld + m] = s; else Ms[j * ld + i] = s;
        }
    }
    __syncthreads();
    // gradient: 0.5 sum_ij (alpha_i alpha_j - Kinv_ij) dK_ij   (sklearn_gpr.py:558-578)
    double g0 = 0.0, g1 = 0.0, g2 = 0.0;
    for (int i = warp; i < m; i += FF_WARPS) {
        const double ai = al[i], xi = xs[i];
        for (int j = lane; j <= i; j += 32) {
            if (i == j) {
                const double q = ai * ai - Ms[i * ld + m];
                g0 += q * c;                    // dK/dlog c = c k, k_ii = 1
                g2 += q * (noise * wt[i]);      // dK/dlog noise = noise diag(w)
            } else {
                const double q = 2.0 * (ai * al[j] - Ms[j * ld + i]);
                const double d = xi - xs[j];
                double kv, dk;
                kern_both(kind, d * d, kv, dk);
                g0 += q * (c * kv);
                g1 += q * (c * dk);
            }
        }
    }
    g0 = block_sum(g0, red);
    g1 = block_sum(g1, red);
    g2 = block_sum(g2, red);
    if (tid == 0) {
        f_out[e] = nlml;
        g_out[3 * e] = -0.5 * g0;
        g_out[3 * e + 1] = -0.5 * g1;
        g_out[3 * e + 2] = -0.5 * g2;
    }
}

// alpha by substitution (used by the final prediction, which keeps L)
__device__ void alpha_by_substitution(int m, int ld, const double* Ms, const double* yv, double* tmp, double* al) {
    const int tid = threadIdx.x;
    if (tid < 32) {
        for (int i = 0; i < m; ++i) {
            double s = 0.0;
            for (int k = tid; k < i; k += 32) s = fma(Ms[i * ld + k], tmp[k], s);
            s = warp_sum(s);
            if (tid == 0) tmp[i] = (yv[i] - s) / Ms[i * ld + i];
            __syncwarp();
        }
        for (int i = m - 1; i >= 0; --i) {
            double s = 0.0;
            for (int k = i + 1 + tid; k < m; k += 32) s = fma(Ms[k * ld + i], al[k], s);
            s = warp_sum(s);
            if (tid == 0) al[i] = (tmp[i] - s) / Ms[i * ld + i];
            __syncwarp();
        }
    }
    __syncthreads();
}

// Final prediction on the standardised grid: mean = ts (K* alpha) + tm, var = c - diag(V^T V) clipped at 0,
// std = sqrt(var ts^2)   (sklearn_gpr.py:381-385, 392, 414-436; noise term 0 on the grid, :714-715)
__global__ void __launch_bounds__(FF_MAX_THREADS)
final_predict_kernel(const double* __restrict__ X, const double* __restrict__ y, const double* __restrict__ w,
                     const int32_t* __restrict__ m_arr, int mmax, const double* __restrict__ theta, int kind,
                     double gp_alpha, const double* __restrict__ xq, int n, const double* __restrict__ tm_ts,
                     double* __restrict__ mean, double* __restrict__ sd, int32_t* __restrict__ status, int FP_COLS) {
    extern __shared__ double sm[];
    const int tr = blockIdx.x, tid = threadIdx.x;
    const int m = m_arr[tr];
    const int ld = (m + 1) | 1;
    double* Ms = sm;
    double* xs = Ms + (size_t)mmax * ((mmax + 1) | 1);
    double* yv = xs + mmax;
    double* al = yv + mmax;
    double* tmp = al + mmax;
    double* Vs = tmp + mmax;  // m x FP_COLS
    const double c = exp(theta[3 * tr]), ls = exp(theta[3 * tr + 1]), noise = exp(theta[3 * tr + 2]);
    build_kernel_matrix(kind, m, ld, c, ls, noise, gp_alpha, X + (size_t)tr * mmax, y + (size_t)tr * mmax,
                        w + (size_t)tr * mmax, Ms, xs, yv);
    const bool ok = cholesky_inplace(m, ld, Ms);
    alpha_by_substitution(m, ld, Ms, yv, tmp, al);
    if (tid == 0) status[tr] = ok ? 0 : 1;
    const double tm = tm_ts[2 * tr], ts = tm_ts[2 * tr + 1];
    const double* xg = xq + (size_t)tr * n;
    for (int j0 = 0; j0 < n; j0 += FP_COLS) {
        const int nc = min(FP_COLS, n - j0);
        __syncthreads();
        for (int p = tid; p < m * FP_COLS; p += FF_THREADS) {
            const int i = p / FP_COLS, jj = p - i * FP_COLS;
            double v = 0.0;
            if (jj < nc) {
                const double d = xg[j0 + jj] / ls - xs[i];
                v = c * kern_val(kind, d * d);
            }
            Vs[p] = v;
        }
        __syncthreads();
        if (tid < nc) {  // mean before the in-place solve destroys K*
            double s = 0.0;
            for (int i = 0; i < m; ++i) s = fma(Vs[i * FP_COLS + tid], al[i], s);
            mean[(size_t)tr * n + j0 + tid] = ts * s + tm;
        }
        __syncthreads();
        for (int k = 0; k < m; ++k) {
            const double inv = 1.0 / Ms[k * ld + k];
            for (int c2 = tid; c2 < FP_COLS; c2 += FF_THREADS) Vs[k * FP_COLS + c2] *= inv;
            __syncthreads();
            const int rem = m - k - 1;
            for (int p = tid; p < rem * FP_COLS; p += FF_THREADS) {
                const int ii = p / FP_COLS, c2 = p - ii * FP_COLS;
                const int i = k + 1 + ii;
                Vs[i * FP_COLS + c2] = fma(-Ms[i * ld + k], Vs[k * FP_COLS + c2], Vs[i * FP_COLS + c2]);
            }
            __syncthreads();
        }
        if (tid < nc) {
            double s = 0.0;
            for (int i = 0; i < m; ++i) s = fma(Vs[i * FP_COLS + tid], Vs[i * FP_COLS + tid], s);
            double var = c - s;
            if (var < 0.0) var = 0.0;
            sd[(size_t)tr * n + j0 + tid] = sqrt(var * (ts * ts));
        }
    }
}

}  // namespace gpet

using namespace gpet;

extern "C" int gpet_lml_f64(const double* X, const double* y, const double* w, const int32_t* m, int mmax,
                            const int32_t* trace_of, const double* theta, int E, int kind, double gp_alpha, double* f,
                            double* g, void* stream) {
    GPET_REQUIRE(X && y && w && m && trace_of && theta && f && g, "gpet_lml_f64: null pointer");
    GPET_REQUIRE(E > 0 && mmax >= 2 && kind >= 0 && kind <= 3, "gpet_lml_f64: bad argument");
    GPET_SUPPORTED(mmax <= GPET_MAX_TRAIN, "gpet_lml_f64: mmax=%d (max %d)", mmax, GPET_MAX_TRAIN);
    const size_t smem = ((size_t)mmax * ((mmax + 1) | 1) + 4 * (size_t)mmax) * sizeof(double);
    GPET_SUPPORTED(smem <= 227 * 1024, "gpet_lml_f64: needs %zu B shared memory", smem);
    cudaError_t e = cudaFuncSetAttribute(lml_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("lml smem attribute: %s", cudaGetErrorString(e));
        return GPET_ERR_CUDA;
    }
    int nt = g_tune[GPET_TUNE_LML_THREADS];
    nt = nt < 64 ? 64 : (nt > 1024 ? 1024 : (nt / 32) * 32);
    lml_kernel<<<E, nt, smem, (cudaStream_t)stream>>>(X, y, w, m, mmax, trace_of, theta, kind, gp_alpha, f, g);
    return check_launch("lml_kernel");
}

extern "C" int gpet_final_predict_f64(const double* X, const double* y, const double* w, const int32_t* m, int mmax, int T,
                                      const double* theta, int kind, double gp_alpha, const double* xq, int n,
                                      const double* tm_ts, double* mean, double* sd, int32_t* status, void* stream) {
    GPET_REQUIRE(X && y && w && m && theta && xq && tm_ts && mean && sd && status, "gpet_final_predict_f64: null pointer");
    GPET_REQUIRE(T > 0 && mmax >= 2 && n > 0 && kind >= 0 && kind <= 3, "gpet_final_predict_f64: bad argument");
    GPET_SUPPORTED(mmax <= GPET_MAX_TRAIN, "gpet_final_predict_f64: mmax=%d (max %d)", mmax, GPET_MAX_TRAIN);
    int cols = 64;
    size_t smem = 0;
    for (; cols >= 8; cols >>= 1) {
        smem = ((size_t)mmax * ((mmax + 1) | 1) + 4 * (size_t)mmax + (size_t)mmax * cols) * sizeof(double);
        if (smem <= 227 * 1024) break;
    }
    GPET_SUPPORTED(cols >= 8, "gpet_final_predict_f64: needs %zu B shared memory (mmax=%d)", smem, mmax);
    cudaError_t e = cudaFuncSetAttribute(final_predict_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("final_predict smem attribute: %s", cudaGetErrorString(e));
        return GPET_ERR_CUDA;
    }
    final_predict_kernel<<<T, 512, smem, (cudaStream_t)stream>>>(X, y, w, m, mmax, theta, kind, gp_alpha, xq, n, tm_ts,
                                                                       mean, sd, status, cols);
    return check_launch("final_predict_kernel");
}
