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
#include "gpet_chol_panels.cuh"
#include "gpet_dense.cuh"

namespace gpet {

constexpr int GRB = 16;  // rows of a column of G solved together (register block)

struct PostScalars {
    double c, sy, ybar, ys;
};

// Steps shared by the posterior kernels: scaling, K, Cholesky (posterior_setup), alpha (one warp), mean.
// smem: Ls[m*ldL], yv[mp], al[mp], tmp[mp], xs[mp] (int), blk[PNB*PNB + PNB].
template <class IX>
__device__ void posterior_setup(const int32_t* __restrict__ xi, const double* __restrict__ y, const double* __restrict__ w,
                                int m, int n, double sigma_f, double noise_y, double gp_alpha,
                                const double* __restrict__ kd, double* Ls, const IX ix, double* yv, double* tmp, int* xs,
                                double* blk, PostScalars* sc, int* flag) {
    const int tid = threadIdx.x;
    for (int i = tid; i < m; i += PT) {
        xs[i] = xi[i];
        yv[i] = y[i];
    }
    if (tid == 0) *flag = 0;
    __syncthreads();
    // gpet.py:228-230: y_s = np.std(y) + 1; y /= y_s; constant = sigma_f**2 / y_s**2; then sklearn_gpr.py:221-227: mean
    // removed, std kept (1.0 when ~0).  The four sums are numpy's pairwise sums (one thread); the element-wise steps in
    // between run on all threads.
    if (tid == 0) sc->ybar = np_pairwise_sum(yv, m) / (double)m;                 // mean of the raw rows
    __syncthreads();
    for (int i = tid; i < m; i += PT) { const double d = yv[i] - sc->ybar; tmp[i] = d * d; }
    __syncthreads();
    if (tid == 0) sc->ys = sqrt(np_pairwise_sum(tmp, m) / (double)m) + 1.0;
    __syncthreads();
    for (int i = tid; i < m; i += PT) yv[i] = yv[i] / sc->ys;
    __syncthreads();
    if (tid == 0) {
        sc->c = (sigma_f * sigma_f) / (sc->ys * sc->ys);
        sc->ybar = np_pairwise_sum(yv, m) / (double)m;
    }
    __syncthreads();
    for (int i = tid; i < m; i += PT) { const double d = yv[i] - sc->ybar; tmp[i] = d * d; }
    __syncthreads();
    if (tid == 0) {
        double sy = sqrt(np_pairwise_sum(tmp, m) / (double)m);
        if (sy < 10.0 * 2.220446049250313e-16) sy = 1.0;
        sc->sy = sy;
    }
    for (int i = tid; i < m; i += PT) yv[i] = yv[i] - sc->ybar;
    __syncthreads();
    const double c = sc->c;
    const bool add_noise = (m != n);  // sklearn_gpr.py:672-677 quirk
    for (int i = tid >> 5; i < m; i += PT / 32) {      // one warp per row, lanes across the columns j <= i
        const int xi_ = xs[i];
        for (int j = tid & 31; j <= i; j += 32) {
            int d = xi_ - xs[j];
            d = d < 0 ? -d : d;
            double v = c * kd[d];
            if (i == j) {
                if (add_noise) v = v + noise_y * w[i];
                v = v + gp_alpha;
            }
            Ls[ix(i, j)] = v;
        }
    }
    __syncthreads();
    cholesky_panels(m, ix, Ls, blk, flag);
}

// alpha = L^-T L^-1 y, run by ONE warp (no block-wide barrier inside)
template <class IX>
__device__ void posterior_alpha_warp(int m, const IX ix, const double* Ls, const double* yv, double* tmp, double* al) {
    const int lane = threadIdx.x & 31;
    for (int i = 0; i < m; ++i) {
        double s = 0.0;
        for (int k = lane; k < i; k += 32) s = fma(Ls[ix(i, k)], tmp[k], s);
        s = warp_sum(s);
        if (lane == 0) tmp[i] = (yv[i] - s) / Ls[ix(i, i)];
        __syncwarp();
    }
    for (int i = m - 1; i >= 0; --i) {
        double s = 0.0;
        for (int k = i + 1 + lane; k < m; k += 32) s = fma(Ls[ix(k, i)], al[k], s);
        s = warp_sum(s);
        if (lane == 0) al[i] = (tmp[i] - s) / Ls[ix(i, i)];
        __syncwarp();
    }
}

// posterior mean on the grid: sy * (K* alpha) + ybar  (sklearn_gpr.py:381-385)
__device__ void posterior_mean(int m, int n, const PostScalars* sc, const double* __restrict__ kd, const int* xs,
                               const double* al, double* __restrict__ mean_out) {
    const double c = sc->c, sy = sc->sy, ybar = sc->ybar;
    for (int j = threadIdx.x; j < n; j += PT) {
        double s = 0.0;
        for (int i = 0; i < m; ++i) {
            int d = j - xs[i];
            d = d < 0 ? -d : d;
            s = fma(c * kd[d], al[i], s);
        }
        mean_out[j] = sy * s + ybar;
    }
}

template <class IX>
__device__ bool posterior_core(const int32_t* __restrict__ xi, const double* __restrict__ y, const double* __restrict__ w,
                               int m, int n, double sigma_f, double noise_y, double gp_alpha,
                               const double* __restrict__ kd, double* Ls, const IX ix, double* yv, double* al, double* tmp,
                               int* xs, double* blk, PostScalars* sc, int* flag, double* __restrict__ mean_out) {
    posterior_setup(xi, y, w, m, n, sigma_f, noise_y, gp_alpha, kd, Ls, ix, yv, tmp, xs, blk, sc, flag);
    if (threadIdx.x < 32) posterior_alpha_warp(m, ix, Ls, yv, tmp, al);
    __syncthreads();
    posterior_mean(m, n, sc, kd, xs, al, mean_out);
    return *flag == 0;
}

__global__ void __launch_bounds__(PT)
posterior_lowrank_kernel(const int32_t* __restrict__ xi, const double* __restrict__ y, const double* __restrict__ w,
                         const int32_t* __restrict__ m_arr, int mmax, int m_cap, int n, const double* __restrict__ sigma_f,
                         double noise_y, double gp_alpha, const double* __restrict__ kd,
                         const double* __restrict__ Ur, const double* __restrict__ lam, int rp,
                         double* __restrict__ mean, double* __restrict__ ys_out, double* __restrict__ Mr,
                         int32_t* __restrict__ status) {
    extern __shared__ double sm[];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int m = m_arr[b];
    if (m > m_cap) {            // the caller's bound on this call's training-set sizes (it sizes the shared memory) is wrong
        if (tid == 0) status[b] = 2;
        return;
    }
    const int ldL = m | 1;
    double* Ls = sm;
    double* Gs = Ls + (size_t)m_cap * (m_cap | 1);
    double* yv = Gs + (size_t)m_cap * rp;
    double* al = yv + m_cap;
    double* tmp = al + m_cap;
    int* xs = (int*)(tmp + m_cap);
    __shared__ PostScalars sc;
    __shared__ int flag;
    __shared__ double blk[PNB * PNB + PNB];
    const FullLowerP ix{ldL};
    posterior_setup(xi + (size_t)b * mmax, y + (size_t)b * mmax, w + (size_t)b * mmax, m, n, sigma_f[b], noise_y, gp_alpha,
                    kd, Ls, ix, yv, tmp, xs, blk, &sc, &flag);
    if (tid == 0) {
        ys_out[b] = sc.ys;
        status[b] = flag ? 1 : 0;
    }
    // warp 0: alpha.  The other warps: G = L^-1 U[I, :] (m x rp), one THREAD per column of G, no barrier: blocks of GRB
    // rows in registers, L broadcast from shared memory - per element the fma sequence of the right-looking form
    // (k ascending, then the scaling by 1 / L_kk), the same bits.
    if (tid < 32) {
        posterior_alpha_warp(m, ix, Ls, yv, tmp, al);
    } else {
        for (int col = tid - 32; col < rp; col += PT - 32) {
            double* gc = Gs + col;
            for (int i0 = 0; i0 < m; i0 += GRB) {
                const int ib = min(GRB, m - i0);
                double acc[GRB];
#pragma unroll
                for (int r = 0; r < GRB; ++r) acc[r] = (r < ib) ? Ur[(size_t)xs[i0 + r] * rp + col] : 0.0;
                if (ib == GRB) {
                    for (int k = 0; k < i0; ++k) {
                        const double gk = gc[k * rp];
#pragma unroll
                        for (int r = 0; r < GRB; ++r) acc[r] = fma(-Ls[(i0 + r) * ldL + k], gk, acc[r]);
                    }
                } else {
                    for (int k = 0; k < i0; ++k) {
                        const double gk = gc[k * rp];
#pragma unroll
                        for (int r = 0; r < GRB; ++r)
                            if (r < ib) acc[r] = fma(-Ls[(i0 + r) * ldL + k], gk, acc[r]);
                    }
                }
#pragma unroll
                for (int r = 0; r < GRB; ++r) {
                    if (r < ib) {
                        acc[r] *= 1.0 / Ls[(i0 + r) * ldL + i0 + r];
#pragma unroll
                        for (int r2 = r + 1; r2 < GRB; ++r2)
                            if (r2 < ib) acc[r2] = fma(-Ls[(i0 + r2) * ldL + i0 + r], acc[r], acc[r2]);
                    }
                }
#pragma unroll
                for (int r = 0; r < GRB; ++r)
                    if (r < ib) gc[(i0 + r) * rp] = acc[r];
            }
        }
    }
    __syncthreads();
    posterior_mean(m, n, &sc, kd, xs, al, mean + (size_t)b * n);
    // Mr = sy^2 (c lam - c^2 lam (G^T G) lam): 4 x 4 register tiles over the lower triangle of tiles; t_ab = t_ba bit for
    // bit (fma(x, y, t) is symmetric in x, y), the two scalings are formed separately (the products are not)
    const double c = sc.c, sy2 = sc.sy * sc.sy;
    double* out = Mr + (size_t)b * rp * rp;
    const int nt = rp >> 2, ntiles = (nt * (nt + 1)) >> 1;          // rp is a multiple of 4
    for (int tile = tid; tile < ntiles; tile += PT) {
        int I = (int)((sqrtf(8.0f * (float)tile + 1.0f) - 1.0f) * 0.5f);
        if (((I + 1) * (I + 2)) >> 1 <= tile) ++I;
        if ((I * (I + 1)) >> 1 > tile) --I;
        const int J = tile - ((I * (I + 1)) >> 1);
        const double* ga = Gs + 4 * I;
        const double* gb = Gs + 4 * J;
        double t[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int q = 0; q < 4; ++q) t[a][q] = 0.0;
        for (int i = 0; i < m; ++i) {
            double va[4], vb[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) { va[a] = ga[i * rp + a]; vb[a] = gb[i * rp + a]; }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int q = 0; q < 4; ++q) t[a][q] = fma(va[a], vb[q], t[a][q]);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int ra = 4 * I + a, rb = 4 * J + q;
                double v = -(c * c) * (lam[ra] * t[a][q] * lam[rb]);
                if (ra == rb) v += c * lam[ra];
                out[(size_t)ra * rp + rb] = sy2 * v;
                if (I != J) out[(size_t)rb * rp + ra] = sy2 * (-(c * c) * (lam[rb] * t[a][q] * lam[ra]));
            }
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
    __shared__ double blk[PNB * PNB + PNB];
    bool ok = posterior_core(xi + (size_t)b * mmax, y + (size_t)b * mmax, w + (size_t)b * mmax, m, n, sigma_f[b], noise_y,
                             gp_alpha, kd, Ls, ix, yv, al, tmp, xs, blk, &sc, &flag, mean + (size_t)b * n);
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
    __shared__ double blk[PNB * PNB + PNB];
    bool ok = posterior_core(xi + (size_t)b * mmax, y + (size_t)b * mmax, w + (size_t)b * mmax, m, n, sigma_f[b], noise_y,
                             gp_alpha, kd, Ls, FullLowerP{ldL}, yv, al, tmp, xs, blk, &sc, &flag, mean + (size_t)b * n);
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
    if (mmax > GPET_MAX_TRAIN) return posterior_big_workspace_bytes(B, mmax, rp);
    if (small_path(mmax, rp)) return 0;
    return (int64_t)B * mmax * rp * 8 + (int64_t)B * 2 * 8 + 256;
}

extern "C" int gpet_posterior_lowrank_f64(const int32_t* xi, const double* y, const double* w, const int32_t* m, int mmax,
                                          int m_cap, int B, int n, const double* sigma_f, double noise_y, double gp_alpha,
                                          const double* kd, const double* Ur, const double* lam, int rp, double* mean,
                                          double* ys, double* Mr, int32_t* status, void* work, void* stream) {
    GPET_REQUIRE(xi && y && w && m && sigma_f && kd && Ur && lam && mean && ys && Mr && status,
                 "gpet_posterior_lowrank_f64: null pointer");
    GPET_REQUIRE(B > 0 && n > 1 && mmax >= 2 && rp > 0, "gpet_posterior_lowrank_f64: bad shape");
    GPET_SUPPORTED(rp <= GPET_MAX_RANK, "gpet_posterior_lowrank_f64: rp=%d (max %d)", rp, GPET_MAX_RANK);
    GPET_REQUIRE((rp & 3) == 0, "gpet_posterior_lowrank_f64: rp must be a multiple of 4");
    if (m_cap <= 0 || m_cap > mmax) m_cap = mmax;
    if (m_cap < 2) m_cap = 2;
    if (mmax > GPET_MAX_TRAIN) {      // training matrices in HBM, blocked factorisation (gpet_dense.cu)
        GPET_REQUIRE(work != nullptr, "gpet_posterior_lowrank_f64: workspace required (gpet_posterior_lowrank_workspace_bytes)");
        return posterior_big_lowrank(xi, y, w, m, mmax, m_cap, B, n, sigma_f, noise_y, gp_alpha, kd, Ur, lam, rp, mean, ys, Mr,
                                     status, work, (cudaStream_t)stream);
    }
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
    // the working set is laid out for m_cap training points: early iterations (a handful of observations) then need a few
    // KB per trace instead of the 150 KB of the largest training set, and several CTAs share an SM
    const size_t smem_cap = (core_smem_doubles(mmax) + (size_t)mmax * rp) * sizeof(double);
    const size_t smem = (core_smem_doubles(m_cap) + (size_t)m_cap * rp) * sizeof(double);
    static size_t smem_set[64] = {};      // per device: the attribute only ever grows
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || smem_set[dev] < smem_cap) {
        cudaError_t e = cudaFuncSetAttribute(posterior_lowrank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap);
        if (e != cudaSuccess) {
            set_error("posterior_lowrank smem attribute: %s", cudaGetErrorString(e));
            return GPET_ERR_CUDA;
        }
        if (dev >= 0 && dev < 64) smem_set[dev] = smem_cap;
    }
    posterior_lowrank_kernel<<<B, PT, smem, (cudaStream_t)stream>>>(xi, y, w, m, mmax, m_cap, n, sigma_f, noise_y, gp_alpha, kd,
                                                                   Ur, lam, rp, mean, ys, Mr, status);
    return check_launch("posterior_lowrank_kernel");
}

extern "C" int64_t gpet_posterior_full_workspace_bytes(int B, int mmax, int n) {
    if (mmax > GPET_MAX_TRAIN) return posterior_big_workspace_bytes(B, mmax, n);
    return (int64_t)B * mmax * n * 8 + (int64_t)B * 2 * 8 + 256;
}

extern "C" int gpet_posterior_full_f64(const int32_t* xi, const double* y, const double* w, const int32_t* m, int mmax, int B,
                                       int n, const double* sigma_f, double noise_y, double gp_alpha, const double* kd,
                                       double* mean, double* ys, double* cov, int32_t* status, void* work, void* stream) {
    GPET_REQUIRE(xi && y && w && m && sigma_f && kd && mean && ys && cov && status && work,
                 "gpet_posterior_full_f64: null pointer");
    GPET_REQUIRE(B > 0 && n > 1 && mmax >= 2, "gpet_posterior_full_f64: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    if (mmax > GPET_MAX_TRAIN)        // training matrices in HBM, blocked factorisation (gpet_dense.cu)
        return posterior_big_full(xi, y, w, m, mmax, mmax, B, n, sigma_f, noise_y, gp_alpha, kd, mean, ys, cov, status, work, st);
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
