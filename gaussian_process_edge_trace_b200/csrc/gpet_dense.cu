// Training sets beyond the shared-memory kernels (mmax > GPET_MAX_TRAIN; BASELINE config 3: delta_x = 2 on 4096 columns
// -> up to 2046 training points): the kernel matrices live in HBM and are factored in 64 x 64 blocks.
//
// Reference seams (same arithmetic as gpet_posterior.cu / gpet_finalfit.cu):
//   gpet.py:209-230, sklearn_gpr.py:221-227, 304-320, 379-407   posterior: K, Cholesky, alpha, mean, V = L^-1 K*^T, Sigma
//   sklearn_gpr.py:475-585, 257-262                               -(log marginal likelihood) and its gradient
//   sklearn_gpr.py:379-436, gpet.py:263-266                        final predictive mean / std
//
// Building blocks, all batched over matrices of different sizes (matrix b has m[b] rows; every matrix is stored with the
// leading dimension ld = mmax rounded up to 64 and is treated as padded by an identity block up to the next multiple of
// 64, so every tile is full and no kernel has a ragged edge):
//   potrf   right-looking blocked Cholesky: per block column  diag (panel Cholesky of gpet_chol_panels.cuh in shared
//           memory) -> panel (one row per thread against the 64 x 64 block) -> trailing update (DMMA tiles)
//   trsm    left-looking blocked forward substitution  X = L^-1 R  (R in place; "ident": R = I, only the lower tiles are
//           touched, giving T = L^-1): per block row one launch, tile = R_k - sum_j L_kj X_j as DMMA, then one column
//           per thread against L_kk
//   solve   alpha = L^-T L^-1 y, one CTA per matrix
//   gram    C = V^T V as DMMA tiles with fused epilogues (posterior covariance / reduced covariance / gradient sums)
// fp64 on the tensor pipe: mma.sync.m8n8k4.f64 (tcgen05 has no FP64 form), operands staged k-major in shared memory.
#include "gpet_dense.cuh"

#include "gpet_chol_panels.cuh"
#include "gpet_dmma_tiles.cuh"
#include "gpet_gpkernels.cuh"
#include "gpet_npsum.cuh"

namespace gpet {

// Solves Lkk x = v in place (v in registers), right-looking: as soon as x_q is known every later entry takes its update,
// so the 63 + 62 + ... fma are independent chains; every entry still sees its updates in ascending q, then the
// multiplication by 1 / L_qq (the order of the shared-memory kernels).  Lkk: 64 x 65 in shared memory (lower), inv[64].
__device__ __forceinline__ void solve64(double (&v)[DB], const double* Lkk, const double* inv) {
#pragma unroll
    for (int q = 0; q < DB; ++q) {
        const double x = v[q] * inv[q];
        v[q] = x;
#pragma unroll
        for (int r = q + 1; r < DB; ++r) v[r] = fma(-Lkk[r * DLK + q], x, v[r]);
    }
}

// lower triangle of the 64 x 64 block at G (leading dimension ldg) -> Lkk[64][65], inv[r] = 1 / L_rr
__device__ __forceinline__ void load_diag_block(double* Lkk, double* inv, const double* G, size_t ldg) {
    for (int p = threadIdx.x; p < DB * DB; p += blockDim.x) {
        const int r = p >> 6, c = p & 63;
        const double v = (c <= r) ? G[(size_t)r * ldg + c] : 0.0;
        Lkk[r * DLK + c] = v;
        if (r == c) inv[r] = 1.0 / v;
    }
}

// ---- blocked Cholesky ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PT)
potrf_diag_kernel(double* __restrict__ A, int ld, const int32_t* __restrict__ m, const int32_t* __restrict__ sel, int k0,
                  int32_t* __restrict__ status) {
    const int b = blockIdx.x, tid = threadIdx.x;
    if (k0 >= dense_rows(m, sel, b)) return;
    __shared__ double Ls[DB * DLK];
    __shared__ double blk[PNB * PNB + PNB];
    __shared__ int flag;
    double* Ab = A + (size_t)b * ld * ld + (size_t)k0 * ld + k0;
    for (int p = tid; p < DB * DB; p += PT) {
        const int r = p >> 6, c = p & 63;
        if (c <= r) Ls[r * DLK + c] = Ab[(size_t)r * ld + c];
    }
    if (tid == 0) flag = 0;
    __syncthreads();
    cholesky_panels(DB, FullLowerP{DLK}, Ls, blk, &flag);
    for (int p = tid; p < DB * DB; p += PT) {
        const int r = p >> 6, c = p & 63;
        if (c <= r) Ab[(size_t)r * ld + c] = Ls[r * DLK + c];
    }
    if (tid == 0 && flag) status[b] = 1;
}

// rows below the diagonal block: L_ik = A_ik L_kk^-T, one row per thread
__global__ void __launch_bounds__(PT)
potrf_panel_kernel(double* __restrict__ A, int ld, const int32_t* __restrict__ m, const int32_t* __restrict__ sel, int k0) {
    const int b = blockIdx.y, tid = threadIdx.x;
    const int mp = round_up64(dense_rows(m, sel, b));
    if (k0 + DB >= mp) return;
    __shared__ double Lkk[DB * DLK];
    __shared__ double inv[DB];
    double* Ab = A + (size_t)b * ld * ld;
    load_diag_block(Lkk, inv, Ab + (size_t)k0 * ld + k0, ld);
    __syncthreads();
    const int r = k0 + DB + blockIdx.x * PT + tid;
    if (r >= mp) return;
    double* row = Ab + (size_t)r * ld + k0;
    double v[DB];
#pragma unroll
    for (int c = 0; c < DB; c += 2) {
        const double2 t = *reinterpret_cast<const double2*>(row + c);
        v[c] = t.x;
        v[c + 1] = t.y;
    }
    solve64(v, Lkk, inv);
#pragma unroll
    for (int c = 0; c < DB; c += 2) *reinterpret_cast<double2*>(row + c) = make_double2(v[c], v[c + 1]);
}

// trailing update: A_ij -= L_ik L_jk^T for the lower tiles below / right of block column kb
__global__ void __launch_bounds__(DT)
potrf_update_kernel(double* __restrict__ A, int ld, const int32_t* __restrict__ m, const int32_t* __restrict__ sel, int kb) {
    __shared__ __align__(16) double As[DKC * DLD];
    __shared__ __align__(16) double Bs[DKC * DLD];
    const int b = blockIdx.y;
    const int nt = round_up64(dense_rows(m, sel, b)) / DB;
    const int rem = nt - kb - 1;
    int ti, tj;
    pair_decode(blockIdx.x, ti, tj);
    if (ti >= rem) return;
    const int gi = (kb + 1 + ti) * DB, gj = (kb + 1 + tj) * DB, k0 = kb * DB;
    double* Ab = A + (size_t)b * ld * ld;
    const TilePos tp;
    double acc[4][2][2];
    zero_acc(acc);
    tile_product<true, true>(acc, Ab + (size_t)gi * ld + k0, ld, Ab + (size_t)gj * ld + k0, ld, DB, As, Bs, tp);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            double2* p = reinterpret_cast<double2*>(Ab + (size_t)(gi + tp.row(a)) * ld + gj + tp.col(c, 0));
            double2 v = *p;
            v.x -= acc[a][c][0];
            v.y -= acc[a][c][1];
            *p = v;
        }
}

// ---- blocked forward substitution X = L^-1 R ---------------------------------------------------------------------------------
// One launch per block row kb; CTA = (column tile jt, matrix b).  ident != 0: R = I (never read), tiles right of the
// diagonal are zero and skipped, the sum starts at block row jt.
constexpr int TRSM_SMEM = (2 * DB * DLK + DB) * (int)sizeof(double);
__global__ void __launch_bounds__(DT)
trsm_row_kernel(const double* __restrict__ L, int ld, const int32_t* __restrict__ m, const int32_t* __restrict__ sel,
                double* R, size_t rstride, int ldr, int kb, int ident) {
    extern __shared__ __align__(16) double dsm[];
    double* Ts = dsm;                    // 64 x 65: right-hand side tile minus the accumulated product
    double* Lkk = dsm + DB * DLK;        // 64 x 65
    double* inv = Lkk + DB * DLK;        // 64
    double* As = dsm;                    // operand stages alias Ts / Lkk while the product runs
    double* Bs = dsm + DKC * DLD;
    const int b = blockIdx.y, jt = blockIdx.x, k0 = kb * DB;
    if (k0 >= dense_rows(m, sel, b)) return;
    if (ident && jt > kb) return;
    const int jstart = ident ? jt : 0;
    const double* Lb = L + (size_t)b * ld * ld;
    double* Rb = R + (size_t)b * rstride;
    const TilePos tp;
    double acc[4][2][2];
    zero_acc(acc);
    if (kb > jstart)
        tile_product<true, false>(acc, Lb + (size_t)k0 * ld + jstart * DB, ld, Rb + (size_t)jstart * DB * ldr + jt * DB, ldr,
                                  (kb - jstart) * DB, As, Bs, tp);
    __syncthreads();                     // the stages are dead, Ts / Lkk take their place
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = tp.row(a), cc = tp.col(c, h);
                const double rhs = ident ? ((kb == jt && r == cc) ? 1.0 : 0.0) : Rb[(size_t)(k0 + r) * ldr + jt * DB + cc];
                Ts[r * DLK + cc] = rhs - acc[a][c][h];
            }
    load_diag_block(Lkk, inv, Lb + (size_t)k0 * ld + k0, ld);
    __syncthreads();
    if (threadIdx.x < DB) {
        const int cc = threadIdx.x;
        double v[DB];
#pragma unroll
        for (int r = 0; r < DB; ++r) v[r] = Ts[r * DLK + cc];
        solve64(v, Lkk, inv);
#pragma unroll
        for (int r = 0; r < DB; ++r) Rb[(size_t)(k0 + r) * ldr + jt * DB + cc] = v[r];
    }
}

// ---- alpha = L^-T L^-1 rhs, one CTA per matrix -----------------------------------------------------------------------------------
// rhs row of matrix b: rhs + t * rhs_stride with t = sel ? sel[b] : b; x[b][ld] receives alpha (zero in the padding).
// dynamic smem: tv[ld], av[ld], red[8][64], rr[64], Dk[64][65]
__global__ void __launch_bounds__(DT)
chol_solve_vec_kernel(const double* __restrict__ L, int ld, const int32_t* __restrict__ m, const int32_t* __restrict__ sel,
                      const double* __restrict__ rhs, size_t rhs_stride, double* __restrict__ x) {
    extern __shared__ __align__(16) double dsm[];
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int mr = dense_rows(m, sel, b);
    if (mr == 0) return;
    const int mp = round_up64(mr);
    double* tv = dsm;
    double* av = tv + ld;
    double* red = av + ld;          // 8 x 64
    double* rr = red + 8 * DB;      // 64
    double* Dk = rr + DB;           // 64 x 65
    const double* Lb = L + (size_t)b * ld * ld;
    const double* yb = rhs + (size_t)(sel ? sel[b] : b) * rhs_stride;
    // forward: L t = y
    for (int k0 = 0; k0 < mp; k0 += DB) {
        for (int q = 0; q < 8; ++q) {
            const int i = k0 + warp * 8 + q;
            const double* Li = Lb + (size_t)i * ld;
            double s = 0.0;
            for (int j = lane; j < k0; j += 32) s = fma(Li[j], tv[j], s);
            s = warp_sum(s);
            if (lane == 0) rr[warp * 8 + q] = (i < mr ? yb[i] : 0.0) - s;
        }
        for (int p = tid; p < DB * DB; p += DT) {
            const int r = p >> 6, c = p & 63;
            Dk[r * DLK + c] = (c <= r) ? Lb[(size_t)(k0 + r) * ld + k0 + c] : 0.0;
        }
        __syncthreads();
        if (warp == 0) {
            for (int i = 0; i < DB; ++i) {
                double s = 0.0;
                if (lane < i) s = Dk[i * DLK + lane] * tv[k0 + lane];
                if (lane + 32 < i) s = fma(Dk[i * DLK + lane + 32], tv[k0 + lane + 32], s);
                s = warp_sum(s);
                if (lane == 0) tv[k0 + i] = (rr[i] - s) / Dk[i * DLK + i];
                __syncwarp();
            }
        }
        __syncthreads();
    }
    // backward: L^T a = t
    for (int k0 = mp - DB; k0 >= 0; k0 -= DB) {
        double a0 = 0.0, a1 = 0.0;
        for (int i = k0 + DB + warp; i < mp; i += 8) {
            const double xi = av[i];
            const double* Li = Lb + (size_t)i * ld + k0;
            a0 = fma(Li[lane], xi, a0);
            a1 = fma(Li[lane + 32], xi, a1);
        }
        red[warp * DB + lane] = a0;
        red[warp * DB + lane + 32] = a1;
        for (int p = tid; p < DB * DB; p += DT) {
            const int r = p >> 6, c = p & 63;
            Dk[r * DLK + c] = (c <= r) ? Lb[(size_t)(k0 + r) * ld + k0 + c] : 0.0;
        }
        __syncthreads();
        if (tid < DB) {
            double s = 0.0;
#pragma unroll
            for (int w8 = 0; w8 < 8; ++w8) s += red[w8 * DB + tid];
            rr[tid] = tv[k0 + tid] - s;
        }
        __syncthreads();
        if (warp == 0) {
            for (int i = DB - 1; i >= 0; --i) {
                double s = 0.0;
                if (lane > i) s = Dk[lane * DLK + i] * av[k0 + lane];
                if (lane + 32 > i) s = fma(Dk[(lane + 32) * DLK + i], av[k0 + lane + 32], s);
                s = warp_sum(s);
                if (lane == 0) av[k0 + i] = (rr[i] - s) / Dk[i * DLK + i];
                __syncwarp();
            }
        }
        __syncthreads();
    }
    double* xb = x + (size_t)b * ld;
    for (int i = tid; i < ld; i += DT) xb[i] = (i < mr) ? av[i] : 0.0;
}

// ---- posterior (non-converged branch) ------------------------------------------------------------------------------------------
struct BigScal {
    double c, sy, ybar, ys;
};

// gpet.py:226-230 and sklearn_gpr.py:221-227 with numpy's own pairwise sums (as posterior_setup in gpet_posterior.cu):
// scal[b], centred targets yn[b][ld] (zero padded), ys[b], status[b] = 0.   dynamic smem: 2 * mmax doubles
__global__ void __launch_bounds__(PT)
big_scalars_kernel(const double* __restrict__ y, const int32_t* __restrict__ m_arr, int mmax, int m_cap, int ld,
                   const double* __restrict__ sigma_f, BigScal* __restrict__ scal, double* __restrict__ yn,
                   double* __restrict__ ys_out, int32_t* __restrict__ status) {
    extern __shared__ __align__(16) double dsm[];
    double* yv = dsm;
    double* tmp = dsm + mmax;
    __shared__ BigScal sc;
    const int b = blockIdx.x, tid = threadIdx.x, m = m_arr[b];
    for (int i = tid; i < m; i += PT) yv[i] = y[(size_t)b * mmax + i];
    __syncthreads();
    if (tid == 0) sc.ybar = np_pairwise_sum(yv, m) / (double)m;
    __syncthreads();
    for (int i = tid; i < m; i += PT) { const double d = yv[i] - sc.ybar; tmp[i] = d * d; }
    __syncthreads();
    if (tid == 0) sc.ys = sqrt(np_pairwise_sum(tmp, m) / (double)m) + 1.0;
    __syncthreads();
    for (int i = tid; i < m; i += PT) yv[i] = yv[i] / sc.ys;
    __syncthreads();
    if (tid == 0) {
        const double sf = sigma_f[b];
        sc.c = (sf * sf) / (sc.ys * sc.ys);
        sc.ybar = np_pairwise_sum(yv, m) / (double)m;
    }
    __syncthreads();
    for (int i = tid; i < m; i += PT) { const double d = yv[i] - sc.ybar; tmp[i] = d * d; }
    __syncthreads();
    if (tid == 0) {
        double sy = sqrt(np_pairwise_sum(tmp, m) / (double)m);
        if (sy < 10.0 * 2.220446049250313e-16) sy = 1.0;
        sc.sy = sy;
    }
    __syncthreads();
    for (int i = tid; i < ld; i += PT) yn[(size_t)b * ld + i] = (i < m) ? yv[i] - sc.ybar : 0.0;
    if (tid == 0) {
        scal[b] = sc;
        ys_out[b] = sc.ys;
        status[b] = (m > m_cap) ? 2 : 0;      // the launches below cover m_cap rows
    }
}

// K = c kd[|x_i - x_j|] + diag(noise_y w + alpha) on the lower tiles, identity in the padding (sklearn_gpr.py:304-306,
// :672-677: no noise when the training set has exactly edge_length rows).  grid (tile column, tile row, matrix)
__global__ void __launch_bounds__(DT)
big_kbuild_post_kernel(const int32_t* __restrict__ xi, const double* __restrict__ w, const int32_t* __restrict__ m_arr,
                       int mmax, int ld, int n, double noise_y, double gp_alpha, const double* __restrict__ kd,
                       const BigScal* __restrict__ scal, double* __restrict__ K) {
    const int b = blockIdx.z, ti = blockIdx.y, tj = blockIdx.x;
    const int mr = m_arr[b];
    if (tj > ti || ti * DB >= mr) return;
    const double c = scal[b].c;
    const bool add_noise = (mr != n);
    const int32_t* xs = xi + (size_t)b * mmax;
    double* Kb = K + (size_t)b * ld * ld;
    for (int p = threadIdx.x; p < DB * DB; p += DT) {
        const int i = ti * DB + (p >> 6), j = tj * DB + (p & 63);
        double v;
        if (i < mr && j < mr) {
            int d = xs[i] - xs[j];
            d = d < 0 ? -d : d;
            v = c * kd[d];
            if (i == j) {
                if (add_noise) v = v + noise_y * w[(size_t)b * mmax + i];
                v = v + gp_alpha;
            }
        } else {
            v = (i == j) ? 1.0 : 0.0;
        }
        Kb[(size_t)i * ld + j] = v;
    }
}

// posterior mean on the grid: sy (K* alpha) + ybar (sklearn_gpr.py:381-385), the fma order of posterior_mean
__global__ void __launch_bounds__(DT)
big_mean_kernel(const int32_t* __restrict__ xi, const int32_t* __restrict__ m_arr, int mmax, int ld, int n,
                const double* __restrict__ kd, const BigScal* __restrict__ scal, const double* __restrict__ al,
                double* __restrict__ mean) {
    const int b = blockIdx.y, j = blockIdx.x * DT + threadIdx.x;
    if (j >= n) return;
    const int mr = m_arr[b];
    const BigScal sc = scal[b];
    const int32_t* xs = xi + (size_t)b * mmax;
    const double* ab = al + (size_t)b * ld;
    double s = 0.0;
    for (int i = 0; i < mr; ++i) {
        int d = j - xs[i];
        d = d < 0 ? -d : d;
        s = fma(sc.c * kd[d], ab[i], s);
    }
    mean[(size_t)b * n + j] = sc.sy * s + sc.ybar;
}

// right-hand sides of V = L^-1 K*^T: R[i][j] = c kd[|j - x_i|] (Ur == nullptr) or U_r[x_i][j] (low-rank form); zero in the
// padding.  grid (column tile, row tile, matrix)
__global__ void __launch_bounds__(DT)
big_rhs_post_kernel(const int32_t* __restrict__ xi, const int32_t* __restrict__ m_arr, int mmax, int ncols, int ldr,
                    size_t rstride, const double* __restrict__ kd, const BigScal* __restrict__ scal,
                    const double* __restrict__ Ur, double* __restrict__ R) {
    const int b = blockIdx.z, ti = blockIdx.y, tj = blockIdx.x;
    const int mr = m_arr[b];
    if (ti * DB >= mr) return;
    const double c = scal[b].c;
    const int32_t* xs = xi + (size_t)b * mmax;
    double* Rb = R + (size_t)b * rstride;
    for (int p = threadIdx.x; p < DB * DB; p += DT) {
        const int i = ti * DB + (p >> 6), j = tj * DB + (p & 63);
        double v = 0.0;
        if (i < mr && j < ncols) {
            if (Ur) {
                v = Ur[(size_t)xs[i] * ncols + j];
            } else {
                int d = j - xs[i];
                d = d < 0 ? -d : d;
                v = c * kd[d];
            }
        }
        Rb[(size_t)i * ldr + j] = v;
    }
}

// V^T V tiles with the posterior epilogues.  MODE 0: Sigma = sy^2 (c kd[|i-j|] - V^T V) (sklearn_gpr.py:392-407), lower
// tiles computed, both halves written (exactly symmetric).  MODE 1: M_r = sy^2 (c lam - c^2 lam (G^T G) lam).
template <int MODE>
__global__ void __launch_bounds__(DT)
big_gram_kernel(const double* __restrict__ V, size_t rstride, int ldr, const int32_t* __restrict__ m_arr, int nout,
                const double* __restrict__ kd, const double* __restrict__ lam, const BigScal* __restrict__ scal,
                double* __restrict__ out) {
    __shared__ __align__(16) double As[DKC * DLD];
    __shared__ __align__(16) double Bs[DKC * DLD];
    const int b = blockIdx.y;
    int ti, tj;
    pair_decode(blockIdx.x, ti, tj);
    const int mp = round_up64(m_arr[b]);
    const double* Vb = V + (size_t)b * rstride;
    const TilePos tp;
    double acc[4][2][2];
    zero_acc(acc);
    tile_product<false, false>(acc, Vb + ti * DB, ldr, Vb + tj * DB, ldr, mp, As, Bs, tp);
    const BigScal sc = scal[b];
    const double sy2 = sc.sy * sc.sy;
    double* ob = out + (size_t)b * nout * nout;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = ti * DB + tp.row(a), j = tj * DB + tp.col(c, h);
                if (i >= nout || j > i) continue;
                double v;
                if (MODE == 0) {
                    v = (sc.c * kd[i - j] - acc[a][c][h]) * sy2;
                } else {
                    v = -(sc.c * sc.c) * (lam[i] * acc[a][c][h] * lam[j]);
                    if (i == j) v += sc.c * lam[i];
                    v = sy2 * v;
                }
                ob[(size_t)i * nout + j] = v;
                ob[(size_t)j * nout + i] = v;
            }
}

// ---- final fit: objective and prediction ------------------------------------------------------------------------------------------
// K(theta) = c k(X / l) + diag(noise w + alpha) (build_kernel_matrix of gpet_finalfit.cu), identity in the padding.
// Matrix e belongs to trace t = sel ? sel[e] : e and uses theta[e][3].  grid (tile column, tile row, matrix)
__global__ void __launch_bounds__(DT)
big_kbuild_theta_kernel(const double* __restrict__ X, const double* __restrict__ w, const int32_t* __restrict__ m_arr,
                        const int32_t* __restrict__ sel, int mmax, int ld, const double* __restrict__ theta, int kind,
                        double gp_alpha, double* __restrict__ K, int32_t* __restrict__ status) {
    const int e = blockIdx.z, ti = blockIdx.y, tj = blockIdx.x;
    const int t = sel ? sel[e] : e;
    if (t < 0) return;
    const int mr = m_arr[t];
    if (tj > ti || ti * DB >= mr) return;
    if (ti == 0 && threadIdx.x == 0) status[e] = 0;
    const double c = exp(theta[3 * e]), ls = exp(theta[3 * e + 1]), noise = exp(theta[3 * e + 2]);
    const double* Xt = X + (size_t)t * mmax;
    double* Kb = K + (size_t)e * ld * ld;
    for (int p = threadIdx.x; p < DB * DB; p += DT) {
        const int i = ti * DB + (p >> 6), j = tj * DB + (p & 63);
        double v;
        if (i < mr && j < mr) {
            if (i == j) {
                v = (c + noise * w[(size_t)t * mmax + i]) + gp_alpha;
            } else {
                const double d = Xt[i] / ls - Xt[j] / ls;
                v = c * kern_val(kind, d * d);
            }
        } else {
            v = (i == j) ? 1.0 : 0.0;
        }
        Kb[(size_t)i * ld + j] = v;
    }
}

__device__ __forceinline__ double block_sum256(double v, double* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < DT / 32; ++i) t += red[i];
    return t;
}

// gradient sums over one lower 64 x 64 tile of K^-1 = T^T T, T = L^-1 (sklearn_gpr.py:558-578): every
// Kinv_ij = sum_{k >= i} T_ki T_kj is consumed where it is formed.  partial[e][pair][3]
__global__ void __launch_bounds__(DT)
big_lml_grad_kernel(const double* __restrict__ T, int ld, const double* __restrict__ X, const double* __restrict__ w,
                    const int32_t* __restrict__ m_arr, const int32_t* __restrict__ sel, int mmax,
                    const double* __restrict__ theta, int kind, const double* __restrict__ al, int npairs,
                    double* __restrict__ partial) {
    __shared__ __align__(16) double As[DKC * DLD];
    __shared__ __align__(16) double Bs[DKC * DLD];
    __shared__ double red[DT / 32];
    const int e = blockIdx.y;
    const int t = sel ? sel[e] : e;
    if (t < 0) return;
    const int mr = m_arr[t], mp = round_up64(mr);
    int ta, tb;
    pair_decode(blockIdx.x, ta, tb);
    if (ta * DB >= mp) return;
    const double* Tb = T + (size_t)e * ld * ld + (size_t)ta * DB * ld;
    const TilePos tp;
    double acc[4][2][2];
    zero_acc(acc);
    tile_product<false, false>(acc, Tb + ta * DB, ld, Tb + tb * DB, ld, mp - ta * DB, As, Bs, tp);
    const double c = exp(theta[3 * e]), ls = exp(theta[3 * e + 1]), noise = exp(theta[3 * e + 2]);
    const double* Xt = X + (size_t)t * mmax;
    const double* wt = w + (size_t)t * mmax;
    const double* ae = al + (size_t)e * ld;
    double g0 = 0.0, g1 = 0.0, g2 = 0.0;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int cc = 0; cc < 2; ++cc)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = ta * DB + tp.row(a), j = tb * DB + tp.col(cc, h);
                if (i >= mr || j > i) continue;
                const double kinv = acc[a][cc][h], ai = ae[i];
                if (i == j) {
                    const double q = ai * ai - kinv;
                    g0 += q * c;
                    g2 += q * (noise * wt[i]);
                } else {
                    const double q = 2.0 * (ai * ae[j] - kinv);
                    const double d = Xt[i] / ls - Xt[j] / ls;
                    double kv, dk;
                    kern_both(kind, d * d, kv, dk);
                    g0 += q * (c * kv);
                    g1 += q * (c * dk);
                }
            }
    g0 = block_sum256(g0, red);
    g1 = block_sum256(g1, red);
    g2 = block_sum256(g2, red);
    if (threadIdx.x == 0) {
        double* pp = partial + ((size_t)e * npairs + blockIdx.x) * 3;
        pp[0] = g0;
        pp[1] = g1;
        pp[2] = g2;
    }
}

// f = 0.5 y^T alpha + sum log L_ii + m/2 log(2 pi), g = -0.5 (sums of the tile partials in a fixed order);
// f = +inf, g = 0 when the Cholesky failed (sklearn_gpr.py:521-522)
__global__ void __launch_bounds__(DT)
big_lml_finish_kernel(const double* __restrict__ L, int ld, const double* __restrict__ y, const int32_t* __restrict__ m_arr,
                      const int32_t* __restrict__ sel, int mmax, const double* __restrict__ al,
                      const double* __restrict__ partial, int npairs, const int32_t* __restrict__ status,
                      double* __restrict__ f, double* __restrict__ g) {
    __shared__ double red[DT / 32];
    const int e = blockIdx.x, tid = threadIdx.x;
    const int t = sel ? sel[e] : e;
    if (t < 0) return;
    if (status[e] != 0) {
        if (tid == 0) {
            f[e] = __longlong_as_double(0x7ff0000000000000LL);
            g[3 * e] = g[3 * e + 1] = g[3 * e + 2] = 0.0;
        }
        return;
    }
    const int mr = m_arr[t], nt = round_up64(mr) / DB;
    const double* Lb = L + (size_t)e * ld * ld;
    const double* yt = y + (size_t)t * mmax;
    const double* ae = al + (size_t)e * ld;
    double part = 0.0;
    for (int i = tid; i < mr; i += DT) part += 0.5 * yt[i] * ae[i] + log(Lb[(size_t)i * ld + i]);
    const double nlml = block_sum256(part, red) + 0.5 * (double)mr * 1.8378770664093453;
    const int valid = nt * (nt + 1) / 2;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int p = tid; p < valid; p += DT) {
        const double* pp = partial + ((size_t)e * npairs + p) * 3;
        s0 += pp[0];
        s1 += pp[1];
        s2 += pp[2];
    }
    s0 = block_sum256(s0, red);
    s1 = block_sum256(s1, red);
    s2 = block_sum256(s2, red);
    if (tid == 0) {
        f[e] = nlml;
        g[3 * e] = -0.5 * s0;
        g[3 * e + 1] = -0.5 * s1;
        g[3 * e + 2] = -0.5 * s2;
    }
}

// K*^T on the standardised grid: R[i][j] = c k((xq_j / l - x_i / l)^2), zero in the padding.  grid (column tile, row tile, trace)
__global__ void __launch_bounds__(DT)
big_rhs_predict_kernel(const double* __restrict__ X, const int32_t* __restrict__ m_arr, int mmax,
                       const double* __restrict__ theta, int kind, const double* __restrict__ xq, int n, int ldr,
                       size_t rstride, double* __restrict__ R) {
    const int b = blockIdx.z, ti = blockIdx.y, tj = blockIdx.x;
    const int mr = m_arr[b];
    if (ti * DB >= mr) return;
    const double c = exp(theta[3 * b]), ls = exp(theta[3 * b + 1]);
    const double* Xt = X + (size_t)b * mmax;
    double* Rb = R + (size_t)b * rstride;
    for (int p = threadIdx.x; p < DB * DB; p += DT) {
        const int i = ti * DB + (p >> 6), j = tj * DB + (p & 63);
        double v = 0.0;
        if (i < mr && j < n) {
            const double d = xq[(size_t)b * n + j] / ls - Xt[i] / ls;
            v = c * kern_val(kind, d * d);
        }
        Rb[(size_t)i * ldr + j] = v;
    }
}

// mean = ts (K* alpha) + tm (before the solve overwrites K*^T)
__global__ void __launch_bounds__(DT)
big_predict_mean_kernel(const double* __restrict__ R, size_t rstride, int ldr, const int32_t* __restrict__ m_arr, int ld,
                        const double* __restrict__ al, int n, const double* __restrict__ tm_ts, double* __restrict__ mean) {
    const int b = blockIdx.y, j = blockIdx.x * DT + threadIdx.x;
    if (j >= n) return;
    const int mr = m_arr[b];
    const double* Rb = R + (size_t)b * rstride + j;
    const double* ab = al + (size_t)b * ld;
    double s = 0.0;
    for (int i = 0; i < mr; ++i) s = fma(Rb[(size_t)i * ldr], ab[i], s);
    mean[(size_t)b * n + j] = tm_ts[2 * b + 1] * s + tm_ts[2 * b];
}

// sd = sqrt(max(c - diag(V^T V), 0) ts^2)   (sklearn_gpr.py:414-436)
__global__ void __launch_bounds__(DT)
big_predict_sd_kernel(const double* __restrict__ V, size_t rstride, int ldr, const int32_t* __restrict__ m_arr,
                      const double* __restrict__ theta, int n, const double* __restrict__ tm_ts, double* __restrict__ sd) {
    const int b = blockIdx.y, j = blockIdx.x * DT + threadIdx.x;
    if (j >= n) return;
    const int mr = m_arr[b];
    const double* Vb = V + (size_t)b * rstride + j;
    double s = 0.0;
    for (int i = 0; i < mr; ++i) {
        const double v = Vb[(size_t)i * ldr];
        s = fma(v, v, s);
    }
    const double ts = tm_ts[2 * b + 1];
    double var = exp(theta[3 * b]) - s;
    if (var < 0.0) var = 0.0;
    sd[(size_t)b * n + j] = sqrt(var * (ts * ts));
}

// ---- host side -----------------------------------------------------------------------------------------------------------------
static int set_smem(const void* fn, int bytes, const char* what) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) {
        set_error("%s shared-memory attribute (%d B): %s", what, bytes, cudaGetErrorString(e));
        return GPET_ERR_CUDA;
    }
    return GPET_OK;
}

// in-place lower Cholesky of nb matrices (rows <= mcap)
static int potrf_batched(double* A, int ld, const int32_t* m, const int32_t* sel, int nb, int mcap, int32_t* status,
                         cudaStream_t st) {
    const int nt = (mcap + DB - 1) / DB;
    for (int kb = 0; kb < nt; ++kb) {
        potrf_diag_kernel<<<nb, PT, 0, st>>>(A, ld, m, sel, kb * DB, status);
        const int rem = nt - kb - 1;
        if (rem == 0) break;
        potrf_panel_kernel<<<dim3((rem * DB + PT - 1) / PT, nb), PT, 0, st>>>(A, ld, m, sel, kb * DB);
        potrf_update_kernel<<<dim3(rem * (rem + 1) / 2, nb), DT, 0, st>>>(A, ld, m, sel, kb);
    }
    return check_launch("dense potrf kernels");
}

// R <- L^-1 R (ncol_tiles column tiles of 64) or, ident != 0, the lower tiles of R <- L^-1
static int trsm_batched(const double* L, int ld, const int32_t* m, const int32_t* sel, int nb, int mcap, double* R,
                        size_t rstride, int ldr, int ncol_tiles, int ident, cudaStream_t st) {
    int rc = set_smem((const void*)trsm_row_kernel, TRSM_SMEM, "trsm_row_kernel");
    if (rc) return rc;
    const int nt = (mcap + DB - 1) / DB;
    for (int kb = 0; kb < nt; ++kb)
        trsm_row_kernel<<<dim3(ident ? kb + 1 : ncol_tiles, nb), DT, TRSM_SMEM, st>>>(L, ld, m, sel, R, rstride, ldr, kb, ident);
    return check_launch("trsm_row_kernel");
}

static int solve_vec_batched(const double* L, int ld, const int32_t* m, const int32_t* sel, int nb, const double* rhs,
                             size_t rhs_stride, double* x, cudaStream_t st) {
    const int smem = (2 * ld + 8 * DB + DB + DB * DLK) * (int)sizeof(double);
    GPET_SUPPORTED(smem <= 227 * 1024, "dense solve: %d training points need %d B shared memory", ld, smem);
    int rc = set_smem((const void*)chol_solve_vec_kernel, smem, "chol_solve_vec_kernel");
    if (rc) return rc;
    chol_solve_vec_kernel<<<nb, DT, smem, st>>>(L, ld, m, sel, rhs, rhs_stride, x);
    return check_launch("chol_solve_vec_kernel");
}

struct PostWork {
    double *K, *yn, *al, *R;
    BigScal* scal;
    int ld, ldr;
    size_t rstride;
};
static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
static int64_t carve_post(void* work, int B, int mmax, int ncols, PostWork* pw) {
    const int ld = h_round_up64(mmax), ldr = h_round_up64(ncols);
    size_t off = 0;
    char* base = work ? (char*)(((uintptr_t)work + 255) & ~(uintptr_t)255) : nullptr;     // the tiles are moved 16 bytes at a time
    auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += align256(bytes); return p; };
    double* K = (double*)take((size_t)B * ld * ld * 8);
    double* yn = (double*)take((size_t)B * ld * 8);
    double* al = (double*)take((size_t)B * ld * 8);
    BigScal* scal = (BigScal*)take((size_t)B * sizeof(BigScal));
    double* R = (double*)take((size_t)B * ld * ldr * 8);
    if (pw) *pw = PostWork{K, yn, al, R, scal, ld, ldr, (size_t)ld * ldr};
    return (int64_t)off + 256;
}

int64_t posterior_big_workspace_bytes(int B, int mmax, int ncols) { return carve_post(nullptr, B, mmax, ncols, nullptr); }

// scalars, K, Cholesky, alpha, mean, V = L^-1 R for the B traces; Ur == nullptr: R = K*^T (n columns), else U_r[I, :]
static int posterior_big_common(const int32_t* xi, const double* y, const double* w, const int32_t* m, int mmax, int m_cap,
                                int B, int n, const double* sigma_f, double noise_y, double gp_alpha, const double* kd,
                                const double* Ur, int ncols, double* mean, double* ys, int32_t* status, const PostWork& pw,
                                cudaStream_t st) {
    GPET_SUPPORTED(B <= 65535, "posterior (large training sets): at most 65535 traces per call");
    if (m_cap <= 0 || m_cap > mmax) m_cap = mmax;
    const int ld = pw.ld, nt = (m_cap + DB - 1) / DB;
    const int sm_sc = 2 * mmax * (int)sizeof(double);
    GPET_SUPPORTED(sm_sc <= 200 * 1024, "posterior: %d training points", mmax);
    int rc = set_smem((const void*)big_scalars_kernel, sm_sc, "big_scalars_kernel");
    if (rc) return rc;
    big_scalars_kernel<<<B, PT, sm_sc, st>>>(y, m, mmax, m_cap, ld, sigma_f, pw.scal, pw.yn, ys, status);
    big_kbuild_post_kernel<<<dim3(nt, nt, B), DT, 0, st>>>(xi, w, m, mmax, ld, n, noise_y, gp_alpha, kd, pw.scal, pw.K);
    rc = check_launch("posterior (large) set-up kernels");
    if (rc) return rc;
    rc = potrf_batched(pw.K, ld, m, nullptr, B, m_cap, status, st);
    if (rc) return rc;
    rc = solve_vec_batched(pw.K, ld, m, nullptr, B, pw.yn, (size_t)ld, pw.al, st);
    if (rc) return rc;
    big_mean_kernel<<<dim3((n + DT - 1) / DT, B), DT, 0, st>>>(xi, m, mmax, ld, n, kd, pw.scal, pw.al, mean);
    const int nct = pw.ldr / DB;
    big_rhs_post_kernel<<<dim3(nct, nt, B), DT, 0, st>>>(xi, m, mmax, ncols, pw.ldr, pw.rstride, kd, pw.scal, Ur, pw.R);
    rc = check_launch("posterior (large) mean / right-hand sides");
    if (rc) return rc;
    return trsm_batched(pw.K, ld, m, nullptr, B, m_cap, pw.R, pw.rstride, pw.ldr, nct, 0, st);
}

int posterior_big_full(const int32_t* xi, const double* y, const double* w, const int32_t* m, int mmax, int m_cap, int B, int n,
                       const double* sigma_f, double noise_y, double gp_alpha, const double* kd, double* mean, double* ys,
                       double* cov, int32_t* status, void* work, cudaStream_t st) {
    PostWork pw;
    carve_post(work, B, mmax, n, &pw);
    int rc = posterior_big_common(xi, y, w, m, mmax, m_cap, B, n, sigma_f, noise_y, gp_alpha, kd, nullptr, n, mean, ys, status,
                                  pw, st);
    if (rc) return rc;
    const int nct = pw.ldr / DB;
    big_gram_kernel<0><<<dim3(nct * (nct + 1) / 2, B), DT, 0, st>>>(pw.R, pw.rstride, pw.ldr, m, n, kd, nullptr, pw.scal, cov);
    return check_launch("big_gram_kernel<0>");
}

int posterior_big_lowrank(const int32_t* xi, const double* y, const double* w, const int32_t* m, int mmax, int m_cap, int B,
                          int n, const double* sigma_f, double noise_y, double gp_alpha, const double* kd, const double* Ur,
                          const double* lam, int rp, double* mean, double* ys, double* Mr, int32_t* status, void* work,
                          cudaStream_t st) {
    PostWork pw;
    carve_post(work, B, mmax, rp, &pw);
    int rc = posterior_big_common(xi, y, w, m, mmax, m_cap, B, n, sigma_f, noise_y, gp_alpha, kd, Ur, rp, mean, ys, status, pw,
                                  st);
    if (rc) return rc;
    const int nct = pw.ldr / DB;
    big_gram_kernel<1><<<dim3(nct * (nct + 1) / 2, B), DT, 0, st>>>(pw.R, pw.rstride, pw.ldr, m, rp, kd, lam, pw.scal, Mr);
    return check_launch("big_gram_kernel<1>");
}

}  // namespace gpet

using namespace gpet;

// ---- C ABI -----------------------------------------------------------------------------------------------------------------------
extern "C" int gpet_dense_potrf_f64(double* A, int ld, const int32_t* m, int B, int m_cap, int32_t* status, void* stream) {
    GPET_REQUIRE(A && m && status && B > 0 && B <= 65535 && ld > 0 && (ld % DB) == 0 && m_cap > 0 && m_cap <= ld,
                 "gpet_dense_potrf_f64: bad argument");
    GPET_REQUIRE(((uintptr_t)A & 15) == 0, "gpet_dense_potrf_f64: A must be 16-byte aligned");
    cudaError_t e = cudaMemsetAsync(status, 0, (size_t)B * sizeof(int32_t), (cudaStream_t)stream);
    if (e != cudaSuccess) {
        set_error("gpet_dense_potrf_f64 memset: %s", cudaGetErrorString(e));
        return GPET_ERR_CUDA;
    }
    return potrf_batched(A, ld, m, nullptr, B, m_cap, status, (cudaStream_t)stream);
}

extern "C" int gpet_dense_trsm_f64(const double* L, int ld, const int32_t* m, int B, int m_cap, double* R, int ldr, int ident,
                                   void* stream) {
    GPET_REQUIRE(L && m && R && B > 0 && B <= 65535 && ld > 0 && (ld % DB) == 0 && m_cap > 0 && m_cap <= ld && ldr > 0 &&
                     (ldr % DB) == 0,
                 "gpet_dense_trsm_f64: bad argument");
    GPET_REQUIRE(!ident || ldr == ld, "gpet_dense_trsm_f64: the inverse needs ldr == ld");
    GPET_REQUIRE((((uintptr_t)L | (uintptr_t)R) & 15) == 0, "gpet_dense_trsm_f64: L and R must be 16-byte aligned");
    return trsm_batched(L, ld, m, nullptr, B, m_cap, R, (size_t)ld * ldr, ldr, ldr / DB, ident, (cudaStream_t)stream);
}

static int64_t lml_big_slot_bytes(int mmax) {
    const int ld = h_round_up64(mmax), nt = ld / DB;
    return (int64_t)(2 * (size_t)ld * ld * 8 + align256((size_t)ld * 8) + align256((size_t)nt * (nt + 1) / 2 * 3 * 8) + 256);
}
extern "C" int64_t gpet_lml_big_workspace_bytes(int E, int mmax) {
    if (E <= 0 || mmax < 2) return 0;
    return (int64_t)E * lml_big_slot_bytes(mmax) + 1024;
}

extern "C" int gpet_lml_big_f64(const double* X, const double* y, const double* w, const int32_t* m, int mmax,
                                const int32_t* trace_of, const double* theta, int E, int kind, double gp_alpha, double* f,
                                double* g, void* work, int64_t work_bytes, void* stream) {
    GPET_REQUIRE(X && y && w && m && trace_of && theta && f && g && work, "gpet_lml_big_f64: null pointer");
    GPET_REQUIRE(E > 0 && mmax >= 2 && kind >= 0 && kind <= 3, "gpet_lml_big_f64: bad argument");
    const int64_t slot = lml_big_slot_bytes(mmax);
    int chunk = (int)((work_bytes - 1024) / slot);
    GPET_REQUIRE(chunk >= 1, "gpet_lml_big_f64: workspace smaller than one evaluation (gpet_lml_big_workspace_bytes)");
    if (chunk > E) chunk = E;
    if (chunk > 65535) chunk = 65535;
    cudaStream_t st = (cudaStream_t)stream;
    const int ld = h_round_up64(mmax), nt = ld / DB, npairs = nt * (nt + 1) / 2;
    char* base = (char*)(((uintptr_t)work + 255) & ~(uintptr_t)255);
    double* Lm = (double*)base;
    double* Tm = Lm + (size_t)chunk * ld * ld;
    double* al = Tm + (size_t)chunk * ld * ld;
    double* partial = al + (size_t)chunk * ld;
    int32_t* status = (int32_t*)(partial + (size_t)chunk * npairs * 3);
    for (int e0 = 0; e0 < E; e0 += chunk) {
        const int ne = (E - e0) < chunk ? (E - e0) : chunk;
        const int32_t* sel = trace_of + e0;
        const double* th = theta + 3 * (size_t)e0;
        big_kbuild_theta_kernel<<<dim3(nt, nt, ne), DT, 0, st>>>(X, w, m, sel, mmax, ld, th, kind, gp_alpha, Lm, status);
        int rc = check_launch("big_kbuild_theta_kernel");
        if (rc) return rc;
        rc = potrf_batched(Lm, ld, m, sel, ne, mmax, status, st);
        if (rc) return rc;
        rc = solve_vec_batched(Lm, ld, m, sel, ne, y, (size_t)mmax, al, st);
        if (rc) return rc;
        rc = trsm_batched(Lm, ld, m, sel, ne, mmax, Tm, (size_t)ld * ld, ld, nt, 1, st);
        if (rc) return rc;
        big_lml_grad_kernel<<<dim3(npairs, ne), DT, 0, st>>>(Tm, ld, X, w, m, sel, mmax, th, kind, al, npairs, partial);
        big_lml_finish_kernel<<<ne, DT, 0, st>>>(Lm, ld, y, m, sel, mmax, al, partial, npairs, status, f + e0, g + 3 * (size_t)e0);
        rc = check_launch("big_lml_grad / finish kernels");
        if (rc) return rc;
    }
    return GPET_OK;
}

// n_rounds x [gpet_lbfgsb_advance_f64 -> gpet_lml_big_f64(trace_of = trace_eval)], then the counters to pinned host memory
// (gpet_fit_rounds_f64 for training sets beyond GPET_MAX_TRAIN)
extern "C" int gpet_fit_rounds_big_f64(const double* X, const double* y, const double* w, const int32_t* m, int mmax, int kind,
                                       double gp_alpha, double* dstate, int32_t* istate, int E, int first, int n_rounds,
                                       const int32_t* trace_of, double* f, double* g, double* theta, int32_t* trace_eval,
                                       int32_t* counters, int32_t* counters_host, void* work, int64_t work_bytes,
                                       void* stream) {
    GPET_REQUIRE(n_rounds > 0 && counters && counters_host, "gpet_fit_rounds_big_f64: bad argument");
    for (int r = 0; r < n_rounds; ++r) {
        int rc = gpet_lbfgsb_advance_f64(dstate, istate, E, (first && r == 0) ? 1 : 0, trace_of, f, g, theta, trace_eval,
                                         counters, stream);
        if (rc != GPET_OK) return rc;
        rc = gpet_lml_big_f64(X, y, w, m, mmax, trace_eval, theta, E, kind, gp_alpha, f, g, work, work_bytes, stream);
        if (rc != GPET_OK) return rc;
    }
    cudaError_t err = cudaMemcpyAsync(counters_host, counters, 3 * sizeof(int32_t), cudaMemcpyDeviceToHost,
                                      (cudaStream_t)stream);
    if (err != cudaSuccess) {
        set_error("gpet_fit_rounds_big_f64 copy: %s", cudaGetErrorString(err));
        return GPET_ERR_CUDA;
    }
    return GPET_OK;
}

extern "C" int64_t gpet_final_predict_big_workspace_bytes(int T, int mmax, int n) {
    if (T <= 0 || mmax < 2 || n <= 0) return 0;
    const int ld = h_round_up64(mmax), ldr = h_round_up64(n);
    return (int64_t)((size_t)T * ld * ld * 8 + (size_t)T * ld * ldr * 8 + align256((size_t)T * ld * 8) + 1024);
}

extern "C" int gpet_final_predict_big_f64(const double* X, const double* y, const double* w, const int32_t* m, int mmax, int T,
                                          const double* theta, int kind, double gp_alpha, const double* xq, int n,
                                          const double* tm_ts, double* mean, double* sd, int32_t* status, void* work,
                                          void* stream) {
    GPET_REQUIRE(X && y && w && m && theta && xq && tm_ts && mean && sd && status && work,
                 "gpet_final_predict_big_f64: null pointer");
    GPET_REQUIRE(T > 0 && T <= 65535 && mmax >= 2 && n > 0 && kind >= 0 && kind <= 3, "gpet_final_predict_big_f64: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int ld = h_round_up64(mmax), ldr = h_round_up64(n), nt = ld / DB, nct = ldr / DB;
    char* base = (char*)(((uintptr_t)work + 255) & ~(uintptr_t)255);
    double* Lm = (double*)base;
    double* R = Lm + (size_t)T * ld * ld;
    double* al = R + (size_t)T * ld * ldr;
    const size_t rstride = (size_t)ld * ldr;
    big_kbuild_theta_kernel<<<dim3(nt, nt, T), DT, 0, st>>>(X, w, m, nullptr, mmax, ld, theta, kind, gp_alpha, Lm, status);
    int rc = check_launch("big_kbuild_theta_kernel");
    if (rc) return rc;
    rc = potrf_batched(Lm, ld, m, nullptr, T, mmax, status, st);
    if (rc) return rc;
    rc = solve_vec_batched(Lm, ld, m, nullptr, T, y, (size_t)mmax, al, st);
    if (rc) return rc;
    big_rhs_predict_kernel<<<dim3(nct, nt, T), DT, 0, st>>>(X, m, mmax, theta, kind, xq, n, ldr, rstride, R);
    big_predict_mean_kernel<<<dim3((n + DT - 1) / DT, T), DT, 0, st>>>(R, rstride, ldr, m, ld, al, n, tm_ts, mean);
    rc = check_launch("final prediction (large) right-hand sides / mean");
    if (rc) return rc;
    rc = trsm_batched(Lm, ld, m, nullptr, T, mmax, R, rstride, ldr, nct, 0, st);
    if (rc) return rc;
    big_predict_sd_kernel<<<dim3((n + DT - 1) / DT, T), DT, 0, st>>>(R, rstride, ldr, m, theta, n, tm_ts, sd);
    return check_launch("big_predict_sd_kernel");
}
