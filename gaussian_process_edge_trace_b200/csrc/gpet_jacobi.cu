// Symmetric eigen-decomposition of the full n x n posterior covariance (full-rank kernels: Matern; numpy's
// multivariate_normal factors it by SVD, sklearn_gpr.py:460-464) for matrices beyond the shared-memory eigensolver of
// gpet_factor.cu: two-sided BLOCK Jacobi in HBM.
//
// The matrix is cut into blocks of 64 columns.  A sweep visits every pair of blocks once (round-robin order: nb / 2
// disjoint pairs per step, nb - 1 steps); per step
//   gather   the 128 x 128 pivot sub-matrices [[A_pp, A_pq], [A_qp, A_qq]] of all pairs of all matrices,
//   solve    them with the batched shared-memory eigensolver (gpet_sym_eig_f64: Householder + QL), J = its eigenvectors,
//   apply    A <- A J on the two block columns and V <- V J (DMMA tiles, the 64 x 128 input slab resident in shared
//            memory so the update is in place), then A <- J^T A on the two block rows; the pivot becomes diag(d).
// Every step removes the pivot's off-diagonal mass from off(A); a handful of sweeps reach off(A) <= 1e-13 ||A||.  The
// arithmetic is GEMM shaped (12 n^3 flops per sweep on the fp64 tensor instruction) - the price of not tridiagonalising
// an HBM-resident matrix column by column.
#include "gpet_common.cuh"
#include "gpet_dmma_tiles.cuh"

namespace gpet {

constexpr int JB = DB;          // block of columns
constexpr int JP = 2 * JB;      // pivot size

// round-robin pairing (circle method): nb blocks (even), step s in [0, nb - 1), pair k in [0, nb / 2)
__host__ __device__ __forceinline__ void rr_pair(int nb, int s, int k, int& p, int& q) {
    const int N = nb - 1;
    int a, b;
    if (k == 0) {
        a = N;
        b = s % N;
    } else {
        a = (s + k) % N;
        b = (s - k + N) % N;
    }
    p = a < b ? a : b;
    q = a < b ? b : a;
}

// A[b] = cov[b] padded to np x np (negative diagonal in the padding: those eigenpairs stay e_i, well separated from the
// positive semi-definite spectrum, and sort last), V[b] = I
__global__ void __launch_bounds__(256)
bj_init_kernel(const double* __restrict__ cov, int n, int np, double* __restrict__ A, double* __restrict__ V) {
    const int b = blockIdx.y, i = blockIdx.x;
    const double* cb = cov + (size_t)b * n * n;
    const double pad = -1.0 - fabs(cb[0]);
    double* Ab = A + ((size_t)b * np + i) * np;
    double* Vb = V + ((size_t)b * np + i) * np;
    for (int j = threadIdx.x; j < np; j += 256) {
        double v;
        if (i < n && j < n) v = (j <= i) ? cb[(size_t)i * n + j] : cb[(size_t)j * n + i];    // lower triangle mirrored
        else v = (i == j) ? pad : 0.0;
        Ab[j] = v;
        Vb[j] = (i == j) ? 1.0 : 0.0;
    }
}

// pivot sub-matrices of step s: P[(b * npairs + k)][128][128], exactly symmetric (lower triangle mirrored)
__global__ void __launch_bounds__(256)
bj_gather_kernel(const double* __restrict__ A, int np, int nb, int s, double* __restrict__ P) {
    const int k = blockIdx.x, b = blockIdx.y, npairs = nb / 2;
    int p, q;
    rr_pair(nb, s, k, p, q);
    const double* Ab = A + (size_t)b * np * np;
    double* Pb = P + ((size_t)b * npairs + k) * JP * JP;
    for (int e = threadIdx.x; e < JP * JP; e += 256) {
        const int r = e / JP, c = e - r * JP;
        const int hi = r >= c ? r : c, lo = r >= c ? c : r;
        const int gr = (hi < JB ? p * JB + hi : q * JB + hi - JB), gc = (lo < JB ? p * JB + lo : q * JB + lo - JB);
        Pb[e] = Ab[(size_t)gr * np + gc];
    }
}

constexpr int BJ_SMEM = (JP + DKC) * DLD * (int)sizeof(double);

// T[rows, (p | q)] <- T[rows, (p | q)] J for T = A (blockIdx.z < B) and T = V (blockIdx.z >= B); CTA = (64 rows, pair)
__global__ void __launch_bounds__(DT)
bj_apply_cols_kernel(double* __restrict__ A, double* __restrict__ V, int B, int np, int nb, int s, const double* __restrict__ Q) {
    extern __shared__ __align__(16) double dsm[];
    double* Xs = dsm;                     // 128 x DLD: Xs[k][i] = T[r0 + i][col(k)]
    double* Qs = dsm + JP * DLD;          // DKC x DLD chunk of J
    const int k = blockIdx.y, npairs = nb / 2, r0 = blockIdx.x * DB;
    const int b = blockIdx.z % B;
    double* T = (blockIdx.z < B ? A : V) + (size_t)b * np * np;
    int p, q;
    rr_pair(nb, s, k, p, q);
    const double* Jb = Q + ((size_t)b * npairs + k) * JP * JP;
    for (int half = 0; half < 2; ++half)
        for (int kc = 0; kc < JB; kc += DKC)
            load_transposed(Xs + (half * JB + kc) * DLD, T + (size_t)r0 * np + (half ? q : p) * JB + kc, np);
    const TilePos tp;
    for (int h = 0; h < 2; ++h) {
        double acc[4][2][2];
        zero_acc(acc);
        for (int k0 = 0; k0 < JP; k0 += DKC) {
            __syncthreads();
            load_kmajor(Qs, Jb + (size_t)k0 * JP + h * JB, JP);
            __syncthreads();
            mma_chunk(acc, Xs + k0 * DLD, Qs, tp);
        }
        double* out = T + (size_t)r0 * np + (h ? q : p) * JB;
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 2; ++c)
                *reinterpret_cast<double2*>(out + (size_t)tp.row(a) * np + tp.col(c, 0)) = make_double2(acc[a][c][0], acc[a][c][1]);
    }
}

// A[(p | q), cols] <- J^T A[(p | q), cols]; CTA = (64 columns, pair, matrix)
__global__ void __launch_bounds__(DT)
bj_apply_rows_kernel(double* __restrict__ A, int np, int nb, int s, const double* __restrict__ Q) {
    extern __shared__ __align__(16) double dsm[];
    double* Xs = dsm;                     // 128 x DLD: Xs[k][j] = A[row(k)][c0 + j]
    double* Qs = dsm + JP * DLD;
    const int k = blockIdx.y, npairs = nb / 2, c0 = blockIdx.x * DB, b = blockIdx.z;
    double* T = A + (size_t)b * np * np;
    int p, q;
    rr_pair(nb, s, k, p, q);
    const double* Jb = Q + ((size_t)b * npairs + k) * JP * JP;
    for (int half = 0; half < 2; ++half)
        for (int kc = 0; kc < JB; kc += DKC)
            load_kmajor(Xs + (half * JB + kc) * DLD, T + (size_t)((half ? q : p) * JB + kc) * np + c0, np);
    const TilePos tp;
    for (int h = 0; h < 2; ++h) {
        double acc[4][2][2];
        zero_acc(acc);
        for (int k0 = 0; k0 < JP; k0 += DKC) {
            __syncthreads();
            load_kmajor(Qs, Jb + (size_t)k0 * JP + h * JB, JP);
            __syncthreads();
            mma_chunk(acc, Qs, Xs + k0 * DLD, tp);
        }
        double* out = T + (size_t)((h ? q : p) * JB) * np + c0;
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 2; ++c)
                *reinterpret_cast<double2*>(out + (size_t)tp.row(a) * np + tp.col(c, 0)) = make_double2(acc[a][c][0], acc[a][c][1]);
    }
}

// the pivot is diagonal now: write diag(d) exactly
__global__ void __launch_bounds__(256)
bj_set_pivot_kernel(double* __restrict__ A, int np, int nb, int s, const double* __restrict__ d) {
    const int k = blockIdx.x, b = blockIdx.y, npairs = nb / 2;
    int p, q;
    rr_pair(nb, s, k, p, q);
    double* Ab = A + (size_t)b * np * np;
    const double* db = d + ((size_t)b * npairs + k) * JP;
    for (int e = threadIdx.x; e < JP * JP; e += 256) {
        const int r = e / JP, c = e - r * JP;
        const int gr = (r < JB ? p * JB + r : q * JB + r - JB), gc = (c < JB ? p * JB + c : q * JB + c - JB);
        Ab[(size_t)gr * np + gc] = (r == c) ? db[r] : 0.0;
    }
}

// off[b][0] += sum_{i != j} A_ij^2, off[b][1] += sum_ij A_ij^2 (zeroed by the caller)
__global__ void __launch_bounds__(256)
bj_offnorm_kernel(const double* __restrict__ A, int np, double* __restrict__ off) {
    const int b = blockIdx.y, i = blockIdx.x;
    const double* row = A + ((size_t)b * np + i) * np;
    double so = 0.0, st = 0.0;
    for (int j = threadIdx.x; j < np; j += 256) {
        const double v = row[j] * row[j];
        st += v;
        if (j != i) so += v;
    }
    so = warp_sum(so);
    st = warp_sum(st);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(off + 2 * b, so);
        atomicAdd(off + 2 * b + 1, st);
    }
}

// order[b][r] = index of the r-th largest diagonal entry (ties by index), by counting
__global__ void __launch_bounds__(256)
bj_rank_kernel(const double* __restrict__ A, int np, int32_t* __restrict__ order) {
    const int b = blockIdx.y, k = blockIdx.x * 256 + threadIdx.x;
    if (k >= np) return;
    const double* Ab = A + (size_t)b * np * np;
    const double dk = Ab[(size_t)k * np + k];
    int r = 0;
    for (int j = 0; j < np; ++j) {
        const double dj = Ab[(size_t)j * np + j];
        r += (dj > dk || (dj == dk && j < k)) ? 1 : 0;
    }
    order[(size_t)b * np + r] = k;
}

// row r of the factor: sign sqrt(max(d, 0)) v^T for the r-th largest eigenpair, sign such that <v, w> > 0
// (numpy: u, s, vt = svd(cov); A = sqrt(s)[:, None] * vt, canonical signs of SURVEY 0.1).  rows n..rp-1 are zero.
__global__ void __launch_bounds__(256)
bj_factor_kernel(const double* __restrict__ A, const double* __restrict__ V, int n, int np, int rp,
                 const int32_t* __restrict__ order, const double* __restrict__ w, double* __restrict__ F) {
    __shared__ double red[8];
    const int b = blockIdx.y, r = blockIdx.x;
    double* Fr = F + ((size_t)b * rp + r) * n;
    if (r >= n) {
        for (int j = threadIdx.x; j < n; j += 256) Fr[j] = 0.0;
        return;
    }
    const int k = order[(size_t)b * np + r];
    const double* Vb = V + (size_t)b * np * np + k;
    double s = 0.0;
    for (int j = threadIdx.x; j < n; j += 256) s = fma(Vb[(size_t)j * np], w[j], s);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    double dk = A[(size_t)b * np * np + (size_t)k * np + k];
    if (dk < 0.0) dk = 0.0;
    const double sc = (t < 0.0 ? -1.0 : 1.0) * sqrt(dk);
    for (int j = threadIdx.x; j < n; j += 256) Fr[j] = sc * Vb[(size_t)j * np];
}

static size_t a256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace gpet

using namespace gpet;

static int bj_check(int B, int n, int np) {
    GPET_REQUIRE(B > 0 && n > 1 && np >= n && (np % JP) == 0, "block Jacobi: np must be a multiple of 128 and >= n");
    GPET_SUPPORTED((int64_t)B * (np / JP) <= 65535 && B * 2 <= 65535, "block Jacobi: batch too large");
    return GPET_OK;
}

extern "C" int64_t gpet_block_jacobi_workspace_bytes(int B, int np) {
    if (B <= 0 || np <= 0 || (np % JP) != 0) return 0;
    const int64_t nm = (int64_t)B * (np / JP);      // pivots per step
    return (int64_t)(2 * a256((size_t)nm * JP * JP * 8) + a256((size_t)nm * JP * 8) + a256((size_t)nm * 4) +
                     a256((size_t)B * np * 4) + 512) + gpet_sym_eig_workspace_bytes((int)nm, JP);
}

extern "C" int gpet_block_jacobi_init_f64(const double* cov, int B, int n, int np, double* A, double* V, void* stream) {
    GPET_REQUIRE(cov && A && V, "gpet_block_jacobi_init_f64: null pointer");
    int rc = bj_check(B, n, np);
    if (rc) return rc;
    bj_init_kernel<<<dim3(np, B), 256, 0, (cudaStream_t)stream>>>(cov, n, np, A, V);
    return check_launch("bj_init_kernel");
}

// One sweep over all block pairs; off[b][2] = (off-diagonal, total) squared Frobenius norms of A[b] after it.
extern "C" int gpet_block_jacobi_sweep_f64(double* A, double* V, int B, int np, double* off, void* work, void* stream) {
    GPET_REQUIRE(A && V && off && work, "gpet_block_jacobi_sweep_f64: null pointer");
    int rc = bj_check(B, np, np);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = np / JB, npairs = nb / 2;
    const int64_t nm = (int64_t)B * npairs;
    char* w = (char*)(((uintptr_t)work + 255) & ~(uintptr_t)255);
    double* P = (double*)w;            w += a256((size_t)nm * JP * JP * 8);
    double* Q = (double*)w;            w += a256((size_t)nm * JP * JP * 8);
    double* d = (double*)w;            w += a256((size_t)nm * JP * 8);
    int32_t* sweeps = (int32_t*)w;     w += a256((size_t)nm * 4);
    w += a256((size_t)B * np * 4);     // (order, used by gpet_block_jacobi_factor_f64)
    void* eig_work = (void*)w;
    cudaError_t e = cudaFuncSetAttribute(bj_apply_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BJ_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(bj_apply_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BJ_SMEM);
    if (e != cudaSuccess) {
        set_error("block Jacobi smem attribute: %s", cudaGetErrorString(e));
        return GPET_ERR_CUDA;
    }
    for (int s = 0; s < nb - 1; ++s) {
        bj_gather_kernel<<<dim3(npairs, B), 256, 0, st>>>(A, np, nb, s, P);
        rc = gpet_sym_eig_f64(P, (int)nm, JP, d, Q, sweeps, eig_work, stream);
        if (rc) return rc;
        bj_apply_cols_kernel<<<dim3(np / DB, npairs, 2 * B), DT, BJ_SMEM, st>>>(A, V, B, np, nb, s, Q);
        bj_apply_rows_kernel<<<dim3(np / DB, npairs, B), DT, BJ_SMEM, st>>>(A, np, nb, s, Q);
        bj_set_pivot_kernel<<<dim3(npairs, B), 256, 0, st>>>(A, np, nb, s, d);
    }
    e = cudaMemsetAsync(off, 0, (size_t)B * 2 * sizeof(double), st);
    if (e != cudaSuccess) {
        set_error("block Jacobi memset: %s", cudaGetErrorString(e));
        return GPET_ERR_CUDA;
    }
    bj_offnorm_kernel<<<dim3(np, B), 256, 0, st>>>(A, np, off);
    return check_launch("block Jacobi sweep kernels");
}

// Factor rows from the converged (A, V): F[b][rp][n] (rp >= n rows, the ones beyond n zero), w[n] sign weights
extern "C" int gpet_block_jacobi_factor_f64(const double* A, const double* V, int B, int n, int np, int rp, const double* w,
                                            double* F, void* work, void* stream) {
    GPET_REQUIRE(A && V && w && F && work && rp >= n, "gpet_block_jacobi_factor_f64: bad argument");
    int rc = bj_check(B, n, np);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nm = (int64_t)B * (np / JP);
    char* wk = (char*)(((uintptr_t)work + 255) & ~(uintptr_t)255);
    wk += 2 * a256((size_t)nm * JP * JP * 8) + a256((size_t)nm * JP * 8) + a256((size_t)nm * 4);
    int32_t* order = (int32_t*)wk;
    bj_rank_kernel<<<dim3((np + 255) / 256, B), 256, 0, st>>>(A, np, order);
    bj_factor_kernel<<<dim3(rp, B), 256, 0, st>>>(A, V, n, np, rp, order, w, F);
    return check_launch("block Jacobi factor kernels");
}
