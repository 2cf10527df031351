// Symmetric eigen-decomposition of the full n x n posterior covariance (full-rank kernels: Matern; numpy's
// multivariate_normal factors it by SVD, sklearn_gpr.py:460-464) for matrices beyond the shared-memory eigensolver of
// gpet_factor.cu: two-sided BLOCK Jacobi in HBM.
//
// The matrix is cut into blocks of JB = 32 (default) or 64 columns.  A sweep visits every pair of blocks once (round-robin
// order: nb / 2 disjoint pairs per step, nb - 1 steps); per step
//   gather   the 2 JB x 2 JB pivot sub-matrices [[A_pp, A_pq], [A_qp, A_qq]] of all pairs of all matrices,
//   rotate   them with the batched in-CTA eigensolver of gpet_factor.cu, J = its eigenvector matrix: 64 x 64 pivots by the
//            parallel cyclic Jacobi kernel capped at two inner sweeps (inexact block Jacobi: what is left of a pivot's
//            off-diagonal part stays in A for the next outer sweep), 128 x 128 pivots by Householder + QL,
//   apply    V <- V J on the two block columns (DMMA tiles, the 64-row slab resident in shared memory: in place) and
//            A <- J^T A J - blocks of 32: both sides in one fused kernel on the lower half in pair space
//            (bj_apply_sym_kernel); blocks of 64: a column pass and a row pass.
// Every step removes (most of) the pivot's off-diagonal mass from off(A).  The caller repeats sweeps until
// off(A) <= tol ||A||; from the second iteration of a trace on it starts from the previous eigenvectors
// (gpet_block_jacobi_warm_f64).  The arithmetic is GEMM shaped (8 n^3 flops per sweep on the fp64 tensor instruction) -
// the price of not tridiagonalising an HBM-resident matrix column by column.
#include "gpet_common.cuh"
#include "gpet_dmma_tiles.cuh"

namespace gpet {

// JB = block of columns (64: 128 x 128 pivots, or 32: 64 x 64 pivots - twice the steps, each pivot solve ~4x cheaper);
// the pivot size is JP = 2 JB.  np is a multiple of 128 either way.
constexpr int JPMAX = 128;
__device__ __forceinline__ int pivot_index(int e, int jb, int p, int q) { return e < jb ? p * jb + e : q * jb + e - jb; }

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
    double* Vb = V ? V + ((size_t)b * np + i) * np : nullptr;
    for (int j = threadIdx.x; j < np; j += 256) {
        double v;
        if (i < n && j < n) v = (j <= i) ? cb[(size_t)i * n + j] : cb[(size_t)j * n + i];    // lower triangle mirrored
        else v = (i == j) ? pad : 0.0;
        Ab[j] = v;
        if (V) Vb[j] = (i == j) ? 1.0 : 0.0;
    }
}

// Batched np x np products for the warm start: C[i][j] = sum_k A(i, k) B[k][j].  TA: A row-major [i][k], else A[k][i]
// (i.e. A^T B).  MODE 0: C = acc.  MODE 1: lower tiles only, C = acc mirrored (exactly symmetric).
// MODE 2: C = 1.5 X - 0.5 acc (one Newton-Schulz step towards the nearest orthogonal matrix, X = A operand).
template <bool TA, int MODE>
__global__ void __launch_bounds__(DT)
bj_gemm_kernel(const double* __restrict__ A, const double* __restrict__ Bm, int np, double* __restrict__ C) {
    __shared__ __align__(16) double As[DKC * DLD];
    __shared__ __align__(16) double Bs[DKC * DLD];
    const size_t mat = (size_t)blockIdx.z * np * np;
    int ti = blockIdx.y, tj = blockIdx.x;
    if (MODE == 1) {
        pair_decode(blockIdx.x, ti, tj);
    }
    const TilePos tp;
    double acc[4][2][2];
    zero_acc(acc);
    const double* Ap = TA ? A + mat + (size_t)ti * DB * np : A + mat + ti * DB;
    tile_product<TA, false>(acc, Ap, np, Bm + mat + tj * DB, np, np, As, Bs, tp);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int i = ti * DB + tp.row(a), j = tj * DB + tp.col(c, 0);
            double v0 = acc[a][c][0], v1 = acc[a][c][1];
            if (MODE == 2) {
                const double2 x = *reinterpret_cast<const double2*>(A + mat + (size_t)i * np + j);
                v0 = 1.5 * x.x - 0.5 * v0;
                v1 = 1.5 * x.y - 0.5 * v1;
            }
            if (MODE == 1) {
                if (j <= i) { C[mat + (size_t)i * np + j] = v0; C[mat + (size_t)j * np + i] = v0; }
                if (j + 1 <= i) { C[mat + (size_t)i * np + j + 1] = v1; C[mat + (size_t)(j + 1) * np + i] = v1; }
            } else {
                *reinterpret_cast<double2*>(C + mat + (size_t)i * np + j) = make_double2(v0, v1);
            }
        }
}

// pivot sub-matrices of step s: P[(b * npairs + k)][128][128], exactly symmetric (lower triangle mirrored)
__global__ void __launch_bounds__(256)
bj_gather_kernel(const double* __restrict__ A, int np, int JB, int s, double* __restrict__ P) {
    const int JP = 2 * JB, nb = np / JB;
    const int k = blockIdx.x, b = blockIdx.y, npairs = nb / 2;
    int p, q;
    rr_pair(nb, s, k, p, q);
    const double* Ab = A + (size_t)b * np * np;
    double* Pb = P + ((size_t)b * npairs + k) * JP * JP;
    for (int e = threadIdx.x; e < JP * JP; e += 256) {
        const int r = e / JP, c = e - r * JP;
        const int hi = r >= c ? r : c, lo = r >= c ? c : r;
        Pb[e] = Ab[(size_t)pivot_index(hi, JB, p, q) * np + pivot_index(lo, JB, p, q)];
    }
}


// T[rows, (p | q)] <- T[rows, (p | q)] J for T = A (blockIdx.z < B) and T = V (blockIdx.z >= B); CTA = (64 rows, pair)
template <int JB>
__global__ void __launch_bounds__(DT)
bj_apply_cols_kernel(double* __restrict__ A, double* __restrict__ V, int B, int np, int s, const double* __restrict__ Q) {
    constexpr int JP = 2 * JB;
    extern __shared__ __align__(16) double dsm[];
    double* Xs = dsm;                     // JP x DLD: Xs[k][i] = T[r0 + i][col(k)]
    double* Qs = dsm + JP * DLD;          // DKC x DLD chunk of J
    const int nb = np / JB, k = blockIdx.y, npairs = nb / 2, r0 = blockIdx.x * DB;
    const int b = blockIdx.z % B;
    double* T = (blockIdx.z < B ? A : V) + (size_t)b * np * np;
    int p, q;
    rr_pair(nb, s, k, p, q);
    const double* Jb = Q + ((size_t)b * npairs + k) * JP * JP;
    for (int half = 0; half < 2; ++half)
        for (int kc = 0; kc < JB; kc += DKC)
            load_transposed(Xs + (half * JB + kc) * DLD, T + (size_t)r0 * np + (half ? q : p) * JB + kc, np);
    const TilePos tp;
    for (int h = 0; h < JP / DB; ++h) {
        double acc[4][2][2];
        zero_acc(acc);
        for (int k0 = 0; k0 < JP; k0 += DKC) {
            __syncthreads();
            load_kmajor(Qs, Jb + (size_t)k0 * JP + h * DB, JP);
            __syncthreads();
            mma_chunk(acc, Xs + k0 * DLD, Qs, tp);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int gc = pivot_index(h * DB + tp.col(c, 0), JB, p, q);      // a pair of columns never straddles two blocks
                *reinterpret_cast<double2*>(T + (size_t)(r0 + tp.row(a)) * np + gc) = make_double2(acc[a][c][0], acc[a][c][1]);
            }
    }
}

// A[(p | q), cols] <- J^T A[(p | q), cols]; CTA = (64 columns, pair, matrix)
template <int JB>
__global__ void __launch_bounds__(DT)
bj_apply_rows_kernel(double* __restrict__ A, int np, int s, const double* __restrict__ Q) {
    constexpr int JP = 2 * JB;
    extern __shared__ __align__(16) double dsm[];
    double* Xs = dsm;                     // JP x DLD: Xs[k][j] = A[row(k)][c0 + j]
    double* Qs = dsm + JP * DLD;
    const int nb = np / JB, k = blockIdx.y, npairs = nb / 2, c0 = blockIdx.x * DB, b = blockIdx.z;
    double* T = A + (size_t)b * np * np;
    int p, q;
    rr_pair(nb, s, k, p, q);
    const double* Jb = Q + ((size_t)b * npairs + k) * JP * JP;
    for (int half = 0; half < 2; ++half)
        for (int kc = 0; kc < JB; kc += DKC)
            load_kmajor(Xs + (half * JB + kc) * DLD, T + (size_t)((half ? q : p) * JB + kc) * np + c0, np);
    const TilePos tp;
    for (int h = 0; h < JP / DB; ++h) {
        double acc[4][2][2];
        zero_acc(acc);
        for (int k0 = 0; k0 < JP; k0 += DKC) {
            __syncthreads();
            load_kmajor(Qs, Jb + (size_t)k0 * JP + h * DB, JP);
            __syncthreads();
            mma_chunk(acc, Qs, Xs + k0 * DLD, tp);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int gr = pivot_index(h * DB + tp.row(a), JB, p, q);
#pragma unroll
            for (int c = 0; c < 2; ++c)
                *reinterpret_cast<double2*>(T + (size_t)gr * np + c0 + tp.col(c, 0)) = make_double2(acc[a][c][0], acc[a][c][1]);
        }
    }
}

// Blocks of 32 (64 x 64 pivots): both sides of A <- J^T A J in ONE kernel, on the lower half of A only.  In pair space the
// update is block-wise, A'[k][l] <- J_k^T A'[k][l] J_l for pairs k, l of the step: CTA = (pair k >= pair l, matrix) reads
// its own 64 x 64 block (four 32 x 32 pieces of A), forms T = X J_l and J_k^T T on the tensor instruction with T handed
// over through shared memory, and writes the block and its mirror image.  No CTA reads what another one writes, so the
// update is in place; against the column pass + row pass it does half the flops (4 n^3 per sweep instead of 8 n^3) and
// moves A once instead of twice, and A stays exactly symmetric.
constexpr int BJS_SMEM = (DB + DKC + DB) * DLD * (int)sizeof(double);
__global__ void __launch_bounds__(DT)
bj_apply_sym_kernel(double* __restrict__ A, int np, int s, const double* __restrict__ Q) {
    constexpr int JB = 32, JP = 64;
    extern __shared__ __align__(16) double dsm[];
    double* Xs = dsm;                     // 64 x DLD: Xs[c][i] = X[i][c]  (A operand of X J_l)
    double* Qs = dsm + JP * DLD;          // DKC x DLD chunk of J_l, then of J_k
    double* Ts = Qs + DKC * DLD;          // 64 x DLD: Ts[r][j] = T[r][j]  (B operand of J_k^T T)
    const int nb = np / JB, npairs = nb / 2, b = blockIdx.y;
    int k, l;
    pair_decode(blockIdx.x, k, l);        // k >= l
    int pk, qk, pl, ql;
    rr_pair(nb, s, k, pk, qk);
    rr_pair(nb, s, l, pl, ql);
    double* T = A + (size_t)b * np * np;
    const double* Jk = Q + ((size_t)b * npairs + k) * JP * JP;
    const double* Jl = Q + ((size_t)b * npairs + l) * JP * JP;
    // X[i][c]: rows of pair k, columns of pair l; the 64 rows are two runs of 32 -> load_transposed in two row halves
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const int ii = lane >> 1, kk = (lane & 1) * 2;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int u = warp + 8 * r;                       // 64 units = 4 row groups x 16 column groups of 4
            const int i = (u & 3) * 16 + ii, c = (u >> 2) * 4 + kk;
            const double2 v = *reinterpret_cast<const double2*>(T + (size_t)pivot_index(i, JB, pk, qk) * np + pivot_index(c, JB, pl, ql));
            Xs[c * DLD + i] = v.x;
            Xs[(c + 1) * DLD + i] = v.y;
        }
    }
    const TilePos tp;
    double acc[4][2][2];
    zero_acc(acc);
    for (int k0 = 0; k0 < JP; k0 += DKC) {                    // T = X J_l
        __syncthreads();
        load_kmajor(Qs, Jl + (size_t)k0 * JP, JP);
        __syncthreads();
        mma_chunk(acc, Xs + k0 * DLD, Qs, tp);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 2; ++c)
            *reinterpret_cast<double2*>(Ts + tp.row(a) * DLD + tp.col(c, 0)) = make_double2(acc[a][c][0], acc[a][c][1]);
    zero_acc(acc);
    for (int k0 = 0; k0 < JP; k0 += DKC) {                    // out = J_k^T T
        __syncthreads();                                      // (first trip: Ts complete, Qs free)
        load_kmajor(Qs, Jk + (size_t)k0 * JP, JP);
        __syncthreads();
        mma_chunk(acc, Qs, Ts + k0 * DLD, tp);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = tp.row(a), j = tp.col(c, h);
                if (k == l && j > i) continue;                // the pivot block: lower triangle, mirrored
                const size_t gi = pivot_index(i, JB, pk, qk), gj = pivot_index(j, JB, pl, ql);
                T[gi * np + gj] = acc[a][c][h];
                T[gj * np + gi] = acc[a][c][h];
            }
}

// the pivot is diagonal now: write diag(d) exactly
__global__ void __launch_bounds__(256)
bj_set_pivot_kernel(double* __restrict__ A, int np, int JB, int s, const double* __restrict__ d) {
    const int JP = 2 * JB, nb = np / JB;
    const int k = blockIdx.x, b = blockIdx.y, npairs = nb / 2;
    int p, q;
    rr_pair(nb, s, k, p, q);
    double* Ab = A + (size_t)b * np * np;
    const double* db = d + ((size_t)b * npairs + k) * JP;
    for (int e = threadIdx.x; e < JP * JP; e += 256) {
        const int r = e / JP, c = e - r * JP;
        Ab[(size_t)pivot_index(r, JB, p, q) * np + pivot_index(c, JB, p, q)] = (r == c) ? db[r] : 0.0;
    }
}

// rowsum[b][i][2] = (sum_{j != i} A_ij^2, sum_j A_ij^2), then off[b][2] = their sums over i - both in a fixed order, so
// every rank of a sample-sharded run (which factors the same covariance redundantly) stops after the same sweep
__global__ void __launch_bounds__(256)
bj_offnorm_rows_kernel(const double* __restrict__ A, int np, double* __restrict__ rowsum) {
    __shared__ double red[2][8];
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
        red[0][threadIdx.x >> 5] = so;
        red[1][threadIdx.x >> 5] = st;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
        rowsum[((size_t)b * np + i) * 2 + threadIdx.x] = t;
    }
}
__global__ void __launch_bounds__(256)
bj_offnorm_sum_kernel(const double* __restrict__ rowsum, int np, double* __restrict__ off) {
    __shared__ double red[2][8];
    const int b = blockIdx.x;
    double so = 0.0, st = 0.0;
    for (int i = threadIdx.x; i < np; i += 256) {
        so += rowsum[((size_t)b * np + i) * 2];
        st += rowsum[((size_t)b * np + i) * 2 + 1];
    }
    so = warp_sum(so);
    st = warp_sum(st);
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = so;
        red[1][threadIdx.x >> 5] = st;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
        off[2 * b + threadIdx.x] = t;
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
    GPET_REQUIRE(B > 0 && n > 1 && np >= n && (np % JPMAX) == 0, "block Jacobi: np must be a multiple of 128 and >= n");
    GPET_SUPPORTED((int64_t)B * (np / 64) <= 65535 && B * 2 <= 65535, "block Jacobi: batch too large");
    return GPET_OK;
}

static int bj_block() { return g_tune[GPET_TUNE_JACOBI_BLOCK] == 64 ? 64 : 32; }

struct BjWork {
    double *P, *Q, *d;
    int32_t *sweeps, *order;
    double* rowsum;
    void* eig;
};
// laid out for the larger of the two block sizes, so the tuning knob may change between calls
static int64_t bj_carve(void* work, int B, int np, BjWork* out) {
    size_t off = 0;
    char* base = work ? (char*)(((uintptr_t)work + 255) & ~(uintptr_t)255) : nullptr;
    auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += a256(bytes); return p; };
    const size_t nm32 = (size_t)B * (np / 64), nm64 = (size_t)B * (np / 128);
    const size_t pq = nm64 * 128 * 128 * 8;               // = nm32 * 64 * 64 * 8 * 2: the larger one
    double* P = (double*)take(pq);
    double* Q = (double*)take(pq);
    double* d = (double*)take(nm32 * 64 * 8);
    int32_t* sw = (int32_t*)take(nm32 * 4);
    int32_t* order = (int32_t*)take((size_t)B * np * 4);
    double* rowsum = (double*)take((size_t)B * np * 2 * 8);
    int64_t e32 = gpet_sym_eig_workspace_bytes((int)nm32, 64), e64 = gpet_sym_eig_workspace_bytes((int)nm64, 128);
    void* eig = (void*)take((size_t)(e32 > e64 ? e32 : e64));
    if (out) *out = BjWork{P, Q, d, sw, order, rowsum, eig};
    return (int64_t)off + 512;
}

extern "C" int64_t gpet_block_jacobi_workspace_bytes(int B, int np) {
    if (B <= 0 || np <= 0 || (np % JPMAX) != 0) return 0;
    return bj_carve(nullptr, B, np, nullptr);
}

extern "C" int gpet_block_jacobi_init_f64(const double* cov, int B, int n, int np, double* A, double* V, void* stream) {
    GPET_REQUIRE(cov && A && V, "gpet_block_jacobi_init_f64: null pointer");
    GPET_REQUIRE((((uintptr_t)A | (uintptr_t)V) & 15) == 0, "gpet_block_jacobi_init_f64: A and V must be 16-byte aligned");
    int rc = bj_check(B, n, np);
    if (rc) return rc;
    bj_init_kernel<<<dim3(np, B), 256, 0, (cudaStream_t)stream>>>(cov, n, np, A, V);
    return check_launch("bj_init_kernel");
}

// Warm start from the eigenvectors of a nearby matrix (the previous iteration's covariance): V <- V (1.5 I - 0.5 V^T V)
// (restores orthogonality lost to rounding over many starts), then A = V^T Sigma V, which is nearly diagonal.
// tmp: 2 * B * np * np doubles.
extern "C" int gpet_block_jacobi_warm_f64(const double* cov, int B, int n, int np, double* A, double* V, void* tmp,
                                          void* stream) {
    GPET_REQUIRE(cov && A && V && tmp, "gpet_block_jacobi_warm_f64: null pointer");
    GPET_REQUIRE((((uintptr_t)A | (uintptr_t)V | (uintptr_t)tmp) & 15) == 0, "gpet_block_jacobi_warm_f64: A, V and tmp must be 16-byte aligned");
    int rc = bj_check(B, n, np);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    double* S = (double*)tmp;
    double* T1 = S + (size_t)B * np * np;
    const int nt = np / DB;
    const dim3 full(nt, nt, B), lower(nt * (nt + 1) / 2, 1, B);
    bj_gemm_kernel<false, 0><<<full, DT, 0, st>>>(V, V, np, S);               // S = V^T V
    bj_gemm_kernel<true, 2><<<full, DT, 0, st>>>(V, S, np, T1);               // T1 = 1.5 V - 0.5 V S
    cudaError_t e = cudaMemcpyAsync(V, T1, (size_t)B * np * np * sizeof(double), cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) {
        set_error("gpet_block_jacobi_warm_f64 copy: %s", cudaGetErrorString(e));
        return GPET_ERR_CUDA;
    }
    bj_init_kernel<<<dim3(np, B), 256, 0, st>>>(cov, n, np, S, nullptr);      // S = padded Sigma
    bj_gemm_kernel<true, 0><<<full, DT, 0, st>>>(S, V, np, T1);               // T1 = Sigma V
    bj_gemm_kernel<false, 1><<<lower, DT, 0, st>>>(V, T1, np, A);             // A = V^T T1, symmetric
    return check_launch("block Jacobi warm-start kernels");
}

template <int JB>
static int bj_sweep(double* A, double* V, int B, int np, const BjWork& w, cudaStream_t st) {
    constexpr int JP = 2 * JB;
    const int nb = np / JB, npairs = nb / 2;
    const int64_t nm = (int64_t)B * npairs;
    const int smem = (JP + DKC) * DLD * (int)sizeof(double);
    const bool sym = (JB == 32) && g_tune[GPET_TUNE_JACOBI_SYM] != 0;     // fused two-sided update of the lower half
    cudaError_t e = cudaFuncSetAttribute(bj_apply_cols_kernel<JB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(bj_apply_rows_kernel<JB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(bj_apply_sym_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BJS_SMEM);
    if (e != cudaSuccess) {
        set_error("block Jacobi smem attribute: %s", cudaGetErrorString(e));
        return GPET_ERR_CUDA;
    }
    for (int s = 0; s < nb - 1; ++s) {
        bj_gather_kernel<<<dim3(npairs, B), 256, 0, st>>>(A, np, JB, s, w.P);
        // pivots: 64 x 64 by the in-CTA parallel Jacobi kernel (nearly diagonal after the first sweeps and after a warm
        // start: one or two inner sweeps), 128 x 128 (does not fit that kernel's shared memory) by Householder + QL
        // GPET_TUNE_JACOBI_INNER = k > 0: at most k inner sweeps per pivot (inexact block Jacobi: the pivot is only rotated
        // towards diagonal form; what is left of its off-diagonal part stays in A for the next outer sweep)
        const int inner = (JB == 32 && g_tune[GPET_TUNE_JACOBI_PIVOT] > 0) ? g_tune[GPET_TUNE_JACOBI_INNER] : 0;
        int rc = sym_eig_run(w.P, (int)nm, JP, w.d, w.Q, w.sweeps, w.eig, (void*)st,
                             JB == 32 ? (g_tune[GPET_TUNE_JACOBI_PIVOT] | (inner << 16)) : 0);
        if (rc) return rc;
        if (sym) {
            bj_apply_cols_kernel<JB><<<dim3(np / DB, npairs, B), DT, smem, st>>>(V, V, B, np, s, w.Q);      // V <- V J only
            bj_apply_sym_kernel<<<dim3(npairs * (npairs + 1) / 2, B), DT, BJS_SMEM, st>>>(A, np, s, w.Q);
        } else {
            bj_apply_cols_kernel<JB><<<dim3(np / DB, npairs, 2 * B), DT, smem, st>>>(A, V, B, np, s, w.Q);
            bj_apply_rows_kernel<JB><<<dim3(np / DB, npairs, B), DT, smem, st>>>(A, np, s, w.Q);
        }
        if (inner == 0) bj_set_pivot_kernel<<<dim3(npairs, B), 256, 0, st>>>(A, np, JB, s, w.d);
    }
    return check_launch("block Jacobi sweep kernels");
}

// One sweep over all block pairs; off[b][2] = (off-diagonal, total) squared Frobenius norms of A[b] after it.
extern "C" int gpet_block_jacobi_sweep_f64(double* A, double* V, int B, int np, double* off, void* work, void* stream) {
    GPET_REQUIRE(A && V && off && work, "gpet_block_jacobi_sweep_f64: null pointer");
    GPET_REQUIRE((((uintptr_t)A | (uintptr_t)V) & 15) == 0, "gpet_block_jacobi_sweep_f64: A and V must be 16-byte aligned");
    int rc = bj_check(B, np, np);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    BjWork w;
    bj_carve(work, B, np, &w);
    rc = bj_block() == 64 ? bj_sweep<64>(A, V, B, np, w, st) : bj_sweep<32>(A, V, B, np, w, st);
    if (rc) return rc;
    bj_offnorm_rows_kernel<<<dim3(np, B), 256, 0, st>>>(A, np, w.rowsum);
    bj_offnorm_sum_kernel<<<B, 256, 0, st>>>(w.rowsum, np, off);
    return check_launch("block Jacobi off-norm kernels");
}

// Factor rows from the converged (A, V): F[b][rp][n] (rp >= n rows, the ones beyond n zero), w[n] sign weights
extern "C" int gpet_block_jacobi_factor_f64(const double* A, const double* V, int B, int n, int np, int rp, const double* w,
                                            double* F, void* work, void* stream) {
    GPET_REQUIRE(A && V && w && F && work && rp >= n, "gpet_block_jacobi_factor_f64: bad argument");
    int rc = bj_check(B, n, np);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    BjWork bw;
    bj_carve(work, B, np, &bw);
    bj_rank_kernel<<<dim3((np + 255) / 256, B), 256, 0, st>>>(A, np, bw.order);
    bj_factor_kernel<<<dim3(rp, B), 256, 0, st>>>(A, V, n, np, rp, bw.order, w, F);
    return check_launch("block Jacobi factor kernels");
}
