// Final hyper-parameter fit (converged branch of fit_predict_GP): batched objective for the host L-BFGS-B
// drivers, and the final predictive mean / std.
//
// Reference seams: sklearn_gpr.py:475-585 (log_marginal_likelihood + analytic gradient), :257-262 (obj_func =
// -LML, -grad), sklearn kernels Constant*RBF|Matern + WeightedWhiteKernel (sklearn_gpr.py:647-694) with
// theta = log[constant, length_scale, noise_level]; sklearn_gpr.py:379-436 (predict with return_std) and
// gpet.py:263-266.
//
// One CTA per evaluation.  Shared memory holds ONE packed lower triangle: K -> L -> L^-1 in place; K^-1 = L^-T L^-1
// is never stored, each entry goes straight into the gradient sums (41 KB at m = 101 => 5 evaluations per SM).
#include "gpet_common.cuh"
#include "gpet_gpkernels.cuh"
#include <type_traits>

namespace gpet {

#define FF_THREADS ((int)blockDim.x)
#define FF_WARPS ((int)(blockDim.x >> 5))
constexpr int FF_MAX_THREADS = 1024;

// Storage of the lower triangle: full rows with an odd leading dimension (final prediction) or packed rows
// (objective kernel: half the shared memory => more evaluations resident per SM).
struct FullLower {
    int ld;
    __device__ __forceinline__ int operator()(int i, int j) const { return i * ld + j; }
};
struct PackedLower {
    __device__ __forceinline__ int operator()(int i, int j) const { return ((i * (i + 1)) >> 1) + j; }
};

// K (lower triangle of Ms) = c k(X/l) + diag(noise w + alpha)
template <class IX>
__device__ void build_kernel_matrix(int kind, int m, IX ix, double c, double ls, double noise, double gp_alpha,
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
            Ms[ix(i, j)] = v;
        }
    }
    __syncthreads();
}

// In-place lower Cholesky, two barriers per column, triangular work mapping.  Returns false (uniformly) on a
// non-positive pivot.
template <class IX>
__device__ bool cholesky_inplace(int m, IX ix, double* Ms) {
    const int tid = threadIdx.x;
    bool ok = true;
    for (int k = 0; k < m; ++k) {
        double dkk = Ms[ix(k, k)];
        if (!(dkk > 0.0)) { ok = false; dkk = 1.0; }
        const double sq = sqrt(dkk), inv = 1.0 / sq;
        for (int i = k + 1 + tid; i < m; i += FF_THREADS) Ms[ix(i, k)] *= inv;
        __syncthreads();
        if (tid == 0) Ms[ix(k, k)] = sq;      // nobody reads the pivot during the trailing update
        for (int i = k + 1 + (tid >> 5); i < m; i += FF_WARPS) {   // rank-1 update, one warp per row
            const double lik = -Ms[ix(i, k)];
            for (int j = k + 1 + (tid & 31); j <= i; j += 32) Ms[ix(i, j)] = fma(lik, Ms[ix(j, k)], Ms[ix(i, j)]);
        }
        __syncthreads();
    }
    return ok;
}

// In-place inverse of the lower-triangular factor (LAPACK dtrti2 column order); every row's dot product is split
// over 4 lanes and reduced with shuffles.
template <class IX>
__device__ void tri_inverse_inplace(int m, IX ix, double* Ms, double* tmp) {
    const int tid = threadIdx.x, quad = tid >> 2, l = tid & 3;
    for (int j = m - 1; j >= 0; --j) {
        const double ajj = 1.0 / Ms[ix(j, j)];
        for (int i = j + 1 + tid; i < m; i += FF_THREADS) tmp[i] = Ms[ix(i, j)];
        __syncthreads();
        const int rem = m - j - 1;
        for (int base = 0; base < rem; base += FF_THREADS / 4) {
            const int i = j + 1 + base + quad;
            double s = 0.0;
            if (i < m)
                for (int k = j + 1 + l; k <= i; k += 4) s = fma(Ms[ix(i, k)], tmp[k], s);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (i < m && l == 0) Ms[ix(i, j)] = -ajj * s;
        }
        if (tid == 0) Ms[ix(j, j)] = ajj;
        __syncthreads();
    }
}

// alpha = K^-1 y = T^T (T y) with T = L^-1 (lower) stored in Ms
template <class IX>
__device__ void alpha_from_inverse(int m, IX ix, const double* Ms, const double* yv, double* tmp, double* al) {
    const int tid = threadIdx.x;
    for (int i = tid; i < m; i += FF_THREADS) {
        double s = 0.0;
        for (int k = 0; k <= i; ++k) s = fma(Ms[ix(i, k)], yv[k], s);
        tmp[i] = s;
    }
    __syncthreads();
    for (int i = tid; i < m; i += FF_THREADS) {
        double s = 0.0;
        for (int k = i; k < m; ++k) s = fma(Ms[ix(k, i)], tmp[k], s);
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
    if (tr < 0) return;     // evaluation slot not in use this round (device-driven fit: the run has ended)
    const int m = m_arr[tr];
    const PackedLower ix;
    double* Ms = sm;                                        // packed lower triangle: K -> L -> L^-1 in place
    double* xs = Ms + ((size_t)mmax * (mmax + 1)) / 2;
    double* yv = xs + mmax;
    double* al = yv + mmax;
    double* tmp = al + mmax;
    const double c = exp(theta[3 * e]), ls = exp(theta[3 * e + 1]), noise = exp(theta[3 * e + 2]);
    const double* wt = w + (size_t)tr * mmax;
    build_kernel_matrix(kind, m, ix, c, ls, noise, gp_alpha, X + (size_t)tr * mmax, y + (size_t)tr * mmax, wt, Ms, xs, yv);
    if (!cholesky_inplace(m, ix, Ms)) {  // sklearn_gpr.py:521-522: LML = -inf, gradient = 0
        if (tid == 0) {
            f_out[e] = __longlong_as_double(0x7ff0000000000000LL);
            g_out[3 * e] = g_out[3 * e + 1] = g_out[3 * e + 2] = 0.0;
        }
        return;
    }
    tri_inverse_inplace(m, ix, Ms, tmp);
    alpha_from_inverse(m, ix, Ms, yv, tmp, al);
    // -LML = 0.5 y^T alpha + sum log diag(L) + m/2 log(2 pi);  diag(L) = 1 / diag(L^-1)
    double part = 0.0;
    for (int i = tid; i < m; i += FF_THREADS) part += 0.5 * yv[i] * al[i] - log(Ms[ix(i, i)]);
    const double nlml = block_sum(part, red) + 0.5 * (double)m * 1.8378770664093453;
    // gradient: 0.5 sum_ij (alpha_i alpha_j - Kinv_ij) dK_ij   (sklearn_gpr.py:558-578) with Kinv = T^T T, T = L^-1:
    // every Kinv_ij = sum_{k >= i} T_ki T_kj is consumed as soon as it is formed, it is never stored.
    const int warp = tid >> 5, lane = tid & 31;
    double g0 = 0.0, g1 = 0.0, g2 = 0.0;
    for (int i = warp; i < m; i += FF_WARPS) {
        const double ai = al[i], xi = xs[i];
        for (int j = lane; j <= i; j += 32) {
            double s0 = 0.0, s1 = 0.0;
            int k = i;
            for (; k + 1 < m; k += 2) {
                s0 = fma(Ms[ix(k, i)], Ms[ix(k, j)], s0);
                s1 = fma(Ms[ix(k + 1, i)], Ms[ix(k + 1, j)], s1);
            }
            if (k < m) s0 = fma(Ms[ix(k, i)], Ms[ix(k, j)], s0);
            const double kinv = s0 + s1;
            if (i == j) {
                const double q = ai * ai - kinv;
                g0 += q * c;                    // dK/dlog c = c k, k_ii = 1
                g2 += q * (noise * wt[i]);      // dK/dlog noise = noise diag(w)
            } else {
                const double q = 2.0 * (ai * al[j] - kinv);
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

// ---------------------------------------------------------------------------------------------------------------------
// Blocked objective kernel: 128 threads per evaluation, packed lower triangle, panels of 8 columns, register tiles.
// The column-at-a-time kernel above spends its time in ~4 m barrier phases of a few hundred cycles each; here a
// panel costs two barriers and the O(m^3) parts (trailing update, L^-1 panel product, K^-1 contraction) run as
// 4x4 / k-split register tiles.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int LB_T = 128;      // threads per evaluation
constexpr int LB_NB = 8;       // panel width

__device__ __forceinline__ int tri_start(int i) { return (i * (i + 1)) >> 1; }

// (row, col) of a linear index into a packed lower triangle (idx < 2^23: float sqrt + fix-up)
__device__ __forceinline__ void tri_decode(int idx, int& i, int& j) {
    i = (int)((sqrtf(8.0f * (float)idx + 1.0f) - 1.0f) * 0.5f);
    if (tri_start(i + 1) <= idx) ++i;
    if (tri_start(i) > idx) --i;
    j = idx - tri_start(i);
}

__device__ __forceinline__ void dmma_884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(LB_T, 4)
lml_blocked_kernel(const double* __restrict__ X, const double* __restrict__ y, const double* __restrict__ w,
                   const int32_t* __restrict__ xcol, const int32_t* __restrict__ m_arr, int mmax,
                   const int32_t* __restrict__ trace_of,
                   const double* __restrict__ theta, int kind, double gp_alpha, double* __restrict__ f_out,
                   double* __restrict__ g_out) {
    extern __shared__ double sm[];
    __shared__ double red[LB_T / 32];
    __shared__ double Dblk[LB_NB * LB_NB + LB_NB];   // factored diagonal block + reciprocal pivots (Cholesky)
    __shared__ int fail;
    const int e = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tr = trace_of[e];
    if (tr < 0) return;     // evaluation slot not in use this round (device-driven fit: the run has ended)
    const int m = m_arr[tr];
    const int npan = (m + LB_NB - 1) / LB_NB;
    double* P = sm;                                        // packed lower triangle: K -> L -> T = L^-1 in place
    double* xs = P + ((size_t)mmax * (mmax + 1)) / 2;
    double* yv = xs + mmax;
    double* al = yv + mmax;
    double* tmp = al + mmax;
    double* dinv = tmp + mmax;                             // 1 / diag(L)
    double* Bt = dinv + mmax;                              // mmax x 8 panel copy used by the inverse; outside the
                                                           // inverse phase it holds the kernel-value table
    int* xc = (int*)(Bt + (size_t)mmax * LB_NB);           // integer pixel columns of the training points (optional)
    const double c = exp(theta[3 * e]), ls = exp(theta[3 * e + 1]), noise = exp(theta[3 * e + 2]);
    const double* wt = w + (size_t)tr * mmax;
    const double* Xt = X + (size_t)tr * mmax;
    const double* yt = y + (size_t)tr * mmax;
    for (int i = tid; i < m; i += LB_T) {
        xs[i] = Xt[i] / ls;
        yv[i] = yt[i];
        if (xcol) xc[i] = xcol[(size_t)tr * mmax + i];
    }
    if (tid == 0) fail = 0;
    __syncthreads();
    // The training inputs are pixel columns (integers) before standardisation, so the RBF kernel takes at most
    // span + 1 distinct values: with xcol given they are tabulated (one exp per distinct distance instead of one per
    // matrix entry, here and again for the gradient), D_ij = ((xcol_i - xcol_j) u)^2 with u = the scaled pixel pitch.
    const int span = xcol ? xc[m - 1] - xc[0] : 0;
    const bool tabled = xcol && kind == 0 && span > 0 && span < mmax * LB_NB;
    const double pitch = tabled ? (xs[m - 1] - xs[0]) / (double)span : 0.0;
    double* tab = Bt;
    auto build_table = [&]() {
        for (int dlt = tid; dlt <= span; dlt += LB_T) {
            const double dd = (double)dlt * pitch;
            tab[dlt] = exp(-0.5 * (dd * dd));
        }
        __syncthreads();
    };
    if (tabled) build_table();
    // ---- K = c k(X/l) + diag(noise w + alpha) ------------------------------------------------------------------------
    {
        const int ntri = tri_start(m);
        int i, j;
        tri_decode(tid, i, j);
        for (int idx = tid; idx < ntri; idx += LB_T) {
            double v;
            if (i == j) {
                v = (c + noise * wt[i]) + gp_alpha;
            } else if (tabled) {
                v = c * tab[xc[i] - xc[j]];
            } else {
                const double d = xs[i] - xs[j];
                v = c * kern_val(kind, d * d);
            }
            P[idx] = v;
            j += LB_T;                      // advance (i, j) by LB_T packed positions
            while (j > i) { j -= i + 1; ++i; }
        }
    }
    __syncthreads();
    // ---- Cholesky, right-looking, panels of 8 ------------------------------------------------------------------------
    for (int k0 = 0; k0 < m; k0 += LB_NB) {
        const int nb = min(LB_NB, m - k0), k1 = k0 + nb;
        // (a) warp 0 factors the diagonal block: every lane runs the same 8x8 register Cholesky (broadcast loads)
        if (warp == 0) {
            double D[LB_NB][LB_NB], iv[LB_NB];
#pragma unroll
            for (int r = 0; r < LB_NB; ++r)
#pragma unroll
                for (int cc = 0; cc <= r; ++cc)
                    D[r][cc] = (r < nb) ? P[tri_start(k0 + r) + k0 + cc] : (r == cc ? 1.0 : 0.0);
            bool bad = false;
#pragma unroll
            for (int cc = 0; cc < LB_NB; ++cc) {
                double d = D[cc][cc];
#pragma unroll
                for (int k = 0; k < cc; ++k) d = fma(-D[cc][k], D[cc][k], d);
                if (!(d > 0.0)) { bad = true; d = 1.0; }
                iv[cc] = rsqrt(d);
                D[cc][cc] = d * iv[cc];
#pragma unroll
                for (int r = cc + 1; r < LB_NB; ++r) {
                    double v = D[r][cc];
#pragma unroll
                    for (int k = 0; k < cc; ++k) v = fma(-D[r][k], D[cc][k], v);
                    D[r][cc] = v * iv[cc];
                }
            }
            __syncwarp();
            if (lane == 0) {
                if (bad) fail = 1;
#pragma unroll
                for (int r = 0; r < LB_NB; ++r) {
#pragma unroll
                    for (int cc = 0; cc <= r; ++cc) {
                        Dblk[r * LB_NB + cc] = D[r][cc];
                        if (r < nb) P[tri_start(k0 + r) + k0 + cc] = D[r][cc];
                    }
                    Dblk[LB_NB * LB_NB + r] = iv[r];
                    if (r < nb) dinv[k0 + r] = iv[r];
                }
            }
        }
        __syncthreads();
        // (b) rows below the block: forward substitution against the block, one row per thread
        for (int r = k1 + tid; r < m; r += LB_T) {
            double* row = P + tri_start(r) + k0;
            double a[LB_NB];
#pragma unroll
            for (int cc = 0; cc < LB_NB; ++cc) a[cc] = row[cc];     // nb == 8 whenever rows exist below
#pragma unroll
            for (int cc = 0; cc < LB_NB; ++cc) {
                double v = a[cc];
#pragma unroll
                for (int k = 0; k < cc; ++k) v = fma(-a[k], Dblk[cc * LB_NB + k], v);
                a[cc] = v * Dblk[LB_NB * LB_NB + cc];
                row[cc] = a[cc];
            }
        }
        __syncthreads();
        // (c) trailing update A22 -= L21 L21^T: one 8x8 tile of the lower triangle per warp iteration, two DMMAs
        const int t = m - k1;
        if (t > 0) {
            // one ROW of 8x8 tiles per warp iteration (longest rows first): the negated A operand of the row is loaded
            // once, every tile costs two B loads, two C loads, two DMMAs and two stores
            const int nt = (t + 7) >> 3;
            const int gr = lane >> 2, gk = lane & 3;
            for (int I = nt - 1 - warp; I >= 0; I -= LB_T / 32) {
                const int i0 = k1 + 8 * I;
                const int ia = i0 + gr;                                  // operand / accumulator row of this lane
                const bool rok = ia < m;
                double* crow = P + tri_start(min(ia, m - 1));
                const double a0 = rok ? -crow[k0 + gk] : 0.0, a1 = rok ? -crow[k0 + gk + 4] : 0.0;
                // tiles left of the diagonal one: every entry and every operand row exists, so nothing is predicated but
                // the store of a row beyond the matrix (its operands come from the clamped row m - 1 and are discarded);
                // the B operand pointer walks down 8 rows per tile: tri_start(j + 8) - tri_start(j) = 8 j + 36
                {
                    int jb = k1 + gr;
                    const double* lb = P + tri_start(jb) + k0 + gk;
                    double* cp = crow + k1 + 2 * gk;
                    for (int J = 0; J < I; ++J) {
                        double c0 = cp[0], c1 = cp[1];
                        dmma_884(c0, c1, a0, lb[0]);
                        dmma_884(c0, c1, a1, lb[4]);
                        if (rok) {
                            cp[0] = c0;
                            cp[1] = c1;
                        }
                        lb += 8 * jb + 36;
                        jb += 8;
                        cp += 8;
                    }
                }
                {   // the diagonal tile: entries above the diagonal and rows beyond the matrix do not exist
                    const int j0 = k1 + 8 * I;
                    const int jb = j0 + gr;
                    const double* lb = P + tri_start(min(jb, m - 1)) + k0 + gk;
                    const double b0 = (jb < m) ? lb[0] : 0.0, b1 = (jb < m) ? lb[4] : 0.0;
                    const int cj = j0 + 2 * gk;                          // accumulator columns cj, cj + 1
                    const bool v0 = rok && (cj <= ia), v1 = rok && (cj + 1 <= ia);
                    double c0 = v0 ? crow[cj] : 0.0, c1 = v1 ? crow[cj + 1] : 0.0;
                    dmma_884(c0, c1, a0, b0);
                    dmma_884(c0, c1, a1, b1);
                    if (v0) crow[cj] = c0;
                    if (v1) crow[cj + 1] = c1;
                }
            }
        }
        __syncthreads();
    }
    if (fail) {  // sklearn_gpr.py:521-522: LML = -inf, gradient = 0
        if (tid == 0) {
            f_out[e] = __longlong_as_double(0x7ff0000000000000LL);
            g_out[3 * e] = g_out[3 * e + 1] = g_out[3 * e + 2] = 0.0;
        }
        return;
    }
    // ---- T = L^-1 in place (LAPACK dtrtri, lower):  T_kk = inv(L_kk),  T[k1:, k0:k1] = -T22 * L[k1:, k0:k1] * T_kk ----
    // all diagonal blocks first: thread (panel, column) solves one column of one block in registers, then stores it
    for (int base = 0; base < npan * LB_NB; base += LB_T) {
        const int task = base + tid;
        const int pan = task >> 3, cc = task & 7, k0 = pan * LB_NB, nb = min(LB_NB, m - k0);
        double x[LB_NB];
#pragma unroll
        for (int r = 0; r < LB_NB; ++r) {
            double v = 0.0;
            if (pan < npan && r < nb && cc < nb && r >= cc) {
                if (r == cc) {
                    v = dinv[k0 + r];
                } else {
                    double sacc = 0.0;
#pragma unroll
                    for (int k = 0; k < LB_NB; ++k)
                        if (k >= cc && k < r) sacc = fma(P[tri_start(k0 + r) + k0 + k], x[k], sacc);
                    v = -sacc * dinv[k0 + r];
                }
            }
            x[r] = v;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < LB_NB; ++r)
            if (pan < npan && r < nb && cc < nb && r >= cc) P[tri_start(k0 + r) + k0 + cc] = x[r];
        __syncthreads();
    }
    for (int k0 = (npan - 2) * LB_NB; k0 >= 0; k0 -= LB_NB) {              // the last panel has no rows below it
        const int k1 = k0 + LB_NB, t = m - k1;
        const double* Dk = P + k0;        // inverse diagonal block: Dk[tri_start(k0 + k) + cc], k >= cc
        // Bt = -(L21 * inv(L_kk)) : t x 8, one output per thread-iteration
        for (int idx = tid; idx < t * LB_NB; idx += LB_T) {
            const int r = idx >> 3, cc = idx & 7;
            const double* row = P + tri_start(k1 + r) + k0;
            double v = 0.0;
#pragma unroll
            for (int k = 0; k < LB_NB; ++k)
                if (k >= cc) v = fma(row[k], Dk[tri_start(k0 + k) + cc], v);
            Bt[idx] = -v;
        }
        __syncthreads();
        // panel <- T22 * Bt: one block of 8 rows per warp iteration, DMMA m8n8k4 along the (triangular) k range
        {
            const int gr = lane >> 2, gk = lane & 3;
            const int nrb = (t + 7) >> 3;
            for (int rb = nrb - 1 - warp; rb >= 0; rb -= LB_T / 32) {     // longest blocks first
                const int r = 8 * rb + gr;                                 // row (relative to k1) of this lane's A operand
                const double* trow = P + tri_start(k1 + min(r, t - 1)) + k1;
                double c0 = 0.0, c1 = 0.0;
                // columns left of the diagonal block of this row block: every T entry exists (a row beyond the matrix
                // reads the clamped last row and is not stored), so the body is two loads and one DMMA
                {
                    const double* pa = trow + gk;
                    const double* pb = Bt + gk * LB_NB + gr;
                    for (int kb = 0; kb < 8 * rb; kb += 4) {
                        dmma_884(c0, c1, *pa, *pb);
                        pa += 4;
                        pb += 4 * LB_NB;
                    }
                }
                const int kend = min(8 * rb + 8, t);                       // columns 0 .. kend-1 can be non-zero
                for (int kb = 8 * rb; kb < kend; kb += 4) {
                    const int k = kb + gk;
                    const double a = (r < t && k <= r) ? trow[k] : 0.0;
                    const double bv = (k < t) ? Bt[k * LB_NB + gr] : 0.0;
                    dmma_884(c0, c1, a, bv);
                }
                if (r < t) {
                    double* orow = P + tri_start(k1 + r) + k0 + 2 * gk;
                    orow[0] = c0;
                    orow[1] = c1;
                }
            }
        }
        __syncthreads();
    }
    // ---- alpha = T^T (T y) ---------------------------------------------------------------------------------------------
    for (int i = tid; i < m; i += LB_T) {
        const double* row = P + tri_start(i);
        double s0 = 0.0, s1 = 0.0;
        int k = 0;
        for (; k + 1 <= i; k += 2) { s0 = fma(row[k], yv[k], s0); s1 = fma(row[k + 1], yv[k + 1], s1); }
        if (k <= i) s0 = fma(row[k], yv[k], s0);
        tmp[i] = s0 + s1;
    }
    __syncthreads();
    for (int i = tid; i < m; i += LB_T) {
        double s = 0.0;
        for (int k = i; k < m; ++k) s = fma(P[tri_start(k) + i], tmp[k], s);
        al[i] = s;
    }
    __syncthreads();
    // -LML = 0.5 y^T alpha + sum log diag(L) + m/2 log(2 pi);  diag(L) = 1 / diag(T)
    double part = 0.0;
    for (int i = tid; i < m; i += LB_T) part += 0.5 * yv[i] * al[i] - log(P[tri_start(i) + i]);
    // ---- gradient: 0.5 sum_ij (alpha_i alpha_j - Kinv_ij) dK_ij (sklearn_gpr.py:558-578), Kinv = T^T T formed in 4x4
    //      tiles and consumed at once
    double g0 = 0.0, g1 = 0.0, g2 = 0.0;
    if (tabled) build_table();        // Bt was the inverse's panel buffer in between
    {
        // Kinv = T^T T tile by tile on the fp64 tensor instruction: one 8x8 tile (I, J <= I) per warp trip,
        // D[i][j] += sum_k T[k][8I + i] T[k][8J + j] over k >= 8I in steps of 4 (two shared-memory loads per DMMA; the
        // 4x4 register-tile form it replaces issued 8 loads per 16 scalar FMAs and its loads of different tiles hit the
        // same banks: 35 % of the kernel's shared-memory wavefronts, profiles/r02_ncu_lml.txt).  A lane then owns the
        // entries (8I + gr, 8J + 2 gk) and (.., + 1) and adds their terms to the three gradient sums.
        const int nt8 = (m + 7) >> 3, ntiles = tri_start(nt8);
        const int gr = lane >> 2, gk = lane & 3;
        for (int tile = warp; tile < ntiles; tile += LB_T / 32) {
            int I, J;
            tri_decode(tile, I, J);
            const int ia = 8 * I + gr, jb = 8 * J + gr;           // column of T this lane loads for the A / B operand
            double c0 = 0.0, c1 = 0.0;
            int k0 = 8 * I;
            // the two k-blocks that cross the diagonal of tile row I (T[k][i] exists for i <= k only), and a ragged end
            auto masked_block = [&](int kk) {
                const int k = kk + gk;
                const double* row = P + tri_start(min(k, m - 1));
                const double a = (k < m && ia <= k) ? row[ia] : 0.0;
                const double bv = (k < m && jb <= k) ? row[jb] : 0.0;
                dmma_884(c0, c1, a, bv);
            };
            for (int h = 0; h < 2 && k0 < m; ++h, k0 += 4) masked_block(k0);
            if (k0 + 4 <= m) {      // full blocks below the diagonal: two loads and one DMMA each, row pointer advanced in place
                int rl = k0 + gk;
                const double* rowp = P + tri_start(rl);
                for (; k0 + 4 <= m; k0 += 4) {
                    dmma_884(c0, c1, rowp[ia], rowp[jb]);
                    rowp += 4 * rl + 10;                        // tri_start(k + 4) - tri_start(k)
                    rl += 4;
                }
            }
            if (k0 < m) masked_block(k0);
            const int i = ia;
            if (i < m && J < I && tabled) {
                // off-diagonal tile, tabulated kernel: every entry is below the diagonal - no branch per entry
                const double ai = al[i];
                const int xi_c = xc[i];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int j = 8 * J + 2 * gk + h;
                    const double q = 2.0 * (ai * al[j] - (h ? c1 : c0));
                    const int dlt = xi_c - xc[j];
                    const double dd = (double)dlt * pitch;
                    const double kv = tab[dlt];
                    g0 += q * (c * kv);
                    g1 += q * (c * (kv * (dd * dd)));
                }
            } else if (i < m) {
                const double ai = al[i], xi = xs[i];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int j = 8 * J + 2 * gk + h;
                    if (j > i) continue;
                    const double kinv = h ? c1 : c0;
                    if (i == j) {
                        const double q = ai * ai - kinv;
                        g0 += q * c;                    // dK/dlog c = c k, k_ii = 1
                        g2 += q * (noise * wt[i]);      // dK/dlog noise = noise diag(w)
                    } else {
                        const double q = 2.0 * (ai * al[j] - kinv);
                        double kv, dk;
                        if (tabled) {
                            const int dlt = xc[i] - xc[j];
                            const double dd = (double)dlt * pitch;
                            kv = tab[dlt];
                            dk = kv * (dd * dd);
                        } else {
                            const double d = xi - xs[j];
                            kern_both(kind, d * d, kv, dk);
                        }
                        g0 += q * (c * kv);
                        g1 += q * (c * dk);
                    }
                }
            }
        }
    }
    const double nlml = block_sum(part, red) + 0.5 * (double)m * 1.8378770664093453;
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
template <class IX>
__device__ void alpha_by_substitution(int m, IX ix, const double* Ms, const double* yv, double* tmp, double* al) {
    const int tid = threadIdx.x;
    if (tid < 32) {
        for (int i = 0; i < m; ++i) {
            double s = 0.0;
            for (int k = tid; k < i; k += 32) s = fma(Ms[ix(i, k)], tmp[k], s);
            s = warp_sum(s);
            if (tid == 0) tmp[i] = (yv[i] - s) / Ms[ix(i, i)];
            __syncwarp();
        }
        for (int i = m - 1; i >= 0; --i) {
            double s = 0.0;
            for (int k = i + 1 + tid; k < m; k += 32) s = fma(Ms[ix(k, i)], al[k], s);
            s = warp_sum(s);
            if (tid == 0) al[i] = (tmp[i] - s) / Ms[ix(i, i)];
            __syncwarp();
        }
    }
    __syncthreads();
}

// Final prediction on the standardised grid: mean = ts (K* alpha) + tm, var = c - diag(V^T V) clipped at 0,
// std = sqrt(var ts^2)   (sklearn_gpr.py:381-385, 392, 414-436; noise term 0 on the grid, :714-715)
constexpr int FP_RB = 16;      // rows of a query column solved together (register block)
template <class IX>      // FullLower (fast index) or PackedLower (training sets whose square does not fit shared memory)
__global__ void __launch_bounds__(512)
final_predict_kernel(const double* __restrict__ X, const double* __restrict__ y, const double* __restrict__ w,
                     const int32_t* __restrict__ m_arr, int mmax, const double* __restrict__ theta, int kind,
                     double gp_alpha, const double* __restrict__ xq, int n, const double* __restrict__ tm_ts,
                     double* __restrict__ mean, double* __restrict__ sd, int32_t* __restrict__ status, int FP_COLS) {
    extern __shared__ double sm[];
    const int tr = blockIdx.x, tid = threadIdx.x;
    const int m = m_arr[tr];
    constexpr bool kPacked = std::is_same<IX, PackedLower>::value;
    IX ix;
    if constexpr (!kPacked) ix.ld = (m + 1) | 1;
    double* Ms = sm;
    double* xs = Ms + (kPacked ? ((size_t)mmax * (mmax + 1)) / 2 : (size_t)mmax * ((mmax + 1) | 1));
    double* yv = xs + mmax;
    double* al = yv + mmax;
    double* tmp = al + mmax;
    double* Vs = tmp + mmax;  // m x FP_COLS
    const double c = exp(theta[3 * tr]), ls = exp(theta[3 * tr + 1]), noise = exp(theta[3 * tr + 2]);
    build_kernel_matrix(kind, m, ix, c, ls, noise, gp_alpha, X + (size_t)tr * mmax, y + (size_t)tr * mmax,
                        w + (size_t)tr * mmax, Ms, xs, yv);
    const bool ok = cholesky_inplace(m, ix, Ms);
    alpha_by_substitution(m, ix, Ms, yv, tmp, al);
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
        // One thread per query column from here on, no barrier: the mean (before the in-place solve destroys K*), then
        // V[:, j] = L^-1 K*[:, j] by forward substitution down the thread's own column - blocks of FP_RB rows in
        // registers (FP_RB independent fma chains per step), L broadcast from shared memory - and the variance.  Per
        // element this is the fma sequence of the barrier-per-column form it replaces (k ascending, then the scaling by
        // 1 / L_kk), so the results are bit-identical; it was 8.6 ms per 1250 traces with 2 m barriers per tile.
        if (tid < nc) {
            double* vc = Vs + tid;
            double s = 0.0;
            for (int i = 0; i < m; ++i) s = fma(vc[i * FP_COLS], al[i], s);
            mean[(size_t)tr * n + j0 + tid] = ts * s + tm;
            for (int i0 = 0; i0 < m; i0 += FP_RB) {
                const int ib = min(FP_RB, m - i0);
                double acc[FP_RB];
#pragma unroll
                for (int r = 0; r < FP_RB; ++r) acc[r] = (r < ib) ? vc[(i0 + r) * FP_COLS] : 0.0;
                if (ib == FP_RB) {
                    for (int k = 0; k < i0; ++k) {
                        const double gk = vc[k * FP_COLS];
#pragma unroll
                        for (int r = 0; r < FP_RB; ++r) acc[r] = fma(-Ms[ix(i0 + r, k)], gk, acc[r]);
                    }
                } else {
                    for (int k = 0; k < i0; ++k) {
                        const double gk = vc[k * FP_COLS];
#pragma unroll
                        for (int r = 0; r < FP_RB; ++r)
                            if (r < ib) acc[r] = fma(-Ms[ix(i0 + r, k)], gk, acc[r]);
                    }
                }
#pragma unroll
                for (int r = 0; r < FP_RB; ++r) {
                    if (r < ib) {
                        acc[r] *= 1.0 / Ms[ix(i0 + r, i0 + r)];
#pragma unroll
                        for (int r2 = r + 1; r2 < FP_RB; ++r2)
                            if (r2 < ib) acc[r2] = fma(-Ms[ix(i0 + r2, i0 + r)], acc[r], acc[r2]);
                    }
                }
#pragma unroll
                for (int r = 0; r < FP_RB; ++r)
                    if (r < ib) vc[(i0 + r) * FP_COLS] = acc[r];
            }
            s = 0.0;
            for (int i = 0; i < m; ++i) s = fma(vc[i * FP_COLS], vc[i * FP_COLS], s);
            double var = c - s;
            if (var < 0.0) var = 0.0;
            sd[(size_t)tr * n + j0 + tid] = sqrt(var * (ts * ts));
        }
    }
}

}  // namespace gpet

using namespace gpet;

extern "C" int gpet_lml_f64(const double* X, const double* y, const double* w, const int32_t* xcol, const int32_t* m,
                            int mmax,
                            const int32_t* trace_of, const double* theta, int E, int kind, double gp_alpha, double* f,
                            double* g, void* stream) {
    GPET_REQUIRE(X && y && w && m && trace_of && theta && f && g, "gpet_lml_f64: null pointer");   // xcol is optional
    GPET_REQUIRE(E > 0 && mmax >= 2 && kind >= 0 && kind <= 3, "gpet_lml_f64: bad argument");
    GPET_SUPPORTED(mmax <= GPET_MAX_TRAIN, "gpet_lml_f64: mmax=%d (max %d)", mmax, GPET_MAX_TRAIN);
    const size_t smem = (((size_t)mmax * (mmax + 1)) / 2 + 4 * (size_t)mmax) * sizeof(double);
    GPET_SUPPORTED(smem <= 227 * 1024, "gpet_lml_f64: needs %zu B shared memory", smem);
    cudaError_t e = cudaFuncSetAttribute(lml_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("lml smem attribute: %s", cudaGetErrorString(e));
        return GPET_ERR_CUDA;
    }
    int nt = g_tune[GPET_TUNE_LML_THREADS];
    if (nt == 0) {     // blocked kernel (default)
        const size_t smem_b = (((size_t)mmax * (mmax + 1)) / 2 + 5 * (size_t)mmax + (size_t)mmax * LB_NB) * sizeof(double) +
                              (size_t)mmax * sizeof(int);
        GPET_SUPPORTED(smem_b <= 227 * 1024, "gpet_lml_f64: needs %zu B shared memory", smem_b);
        e = cudaFuncSetAttribute(lml_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b);
        if (e != cudaSuccess) {
            set_error("lml smem attribute: %s", cudaGetErrorString(e));
            return GPET_ERR_CUDA;
        }
        lml_blocked_kernel<<<E, LB_T, smem_b, (cudaStream_t)stream>>>(X, y, w, xcol, m, mmax, trace_of, theta, kind, gp_alpha,
                                                                     f, g);
        return check_launch("lml_blocked_kernel");
    }
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
    int cols = 128;
    size_t smem = 0;
    bool packed = false;
    for (; cols >= 8; cols >>= 1) {
        smem = ((size_t)mmax * ((mmax + 1) | 1) + 4 * (size_t)mmax + (size_t)mmax * cols) * sizeof(double);
        if (smem <= 227 * 1024) break;
    }
    if (cols < 8) {       // the square of L does not fit: packed triangle
        packed = true;
        for (cols = 128; cols >= 8; cols >>= 1) {
            smem = (((size_t)mmax * (mmax + 1)) / 2 + 4 * (size_t)mmax + (size_t)mmax * cols) * sizeof(double);
            if (smem <= 227 * 1024) break;
        }
    }
    GPET_SUPPORTED(cols >= 8, "gpet_final_predict_f64: needs %zu B shared memory (mmax=%d)", smem, mmax);
    cudaError_t e = packed ? cudaFuncSetAttribute(final_predict_kernel<PackedLower>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                           : cudaFuncSetAttribute(final_predict_kernel<FullLower>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("final_predict smem attribute: %s", cudaGetErrorString(e));
        return GPET_ERR_CUDA;
    }
    if (packed)
        final_predict_kernel<PackedLower><<<T, 512, smem, (cudaStream_t)stream>>>(X, y, w, m, mmax, theta, kind, gp_alpha, xq, n,
                                                                                 tm_ts, mean, sd, status, cols);
    else
        final_predict_kernel<FullLower><<<T, 512, smem, (cudaStream_t)stream>>>(X, y, w, m, mmax, theta, kind, gp_alpha, xq, n,
                                                                               tm_ts, mean, sd, status, cols);
    return check_launch("final_predict_kernel");
}
