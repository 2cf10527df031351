// 64 x 64 output tiles on the fp64 tensor instruction (mma.sync.m8n8k4.f64; tcgen05 has no FP64 form): operand chunks
// staged k-major in shared memory, 8 warps as 2 x 4 with a 32 x 16 warp tile.  Shared by the blocked dense kernels
// (gpet_dense.cu) and the block-Jacobi eigensolver (gpet_jacobi.cu).
#pragma once
#include "gpet_common.cuh"

namespace gpet {

constexpr int DB = 64;    // tile edge
constexpr int DKC = 32;   // k extent staged per step
constexpr int DLD = 68;   // leading dimension of the staged operand tiles: == 4 (mod 16) doubles => the 16 lanes of a
                          // half-warp reading one k row of a fragment hit 16 distinct 8-byte banks
constexpr int DT = 256;   // threads per CTA (8 warps as 2 x 4, warp tile 32 x 16 = 4 x 2 DMMA tiles)
constexpr int DLK = DB + 1;

__device__ __forceinline__ int round_up64(int v) { return (v + 63) & ~63; }
static inline int h_round_up64(int v) { return (v + 63) & ~63; }

// rows of matrix b: m[b], or m[sel[b]] with sel[b] < 0 meaning "skip this slot"
__device__ __forceinline__ int dense_rows(const int32_t* __restrict__ m, const int32_t* __restrict__ sel, int b) {
    const int t = sel ? sel[b] : b;
    return t < 0 ? 0 : m[t];
}

__device__ __forceinline__ void dmma8(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// S[k][i] = G[k * ldg + i]   (k < DKC, i < DB): operand stored k-major; one warp per k row, 16-byte loads
__device__ __forceinline__ void load_kmajor(double* S, const double* G, size_t ldg) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int r = 0; r < DKC / 8; ++r) {
        const int k = warp + 8 * r;
        const double2 v = *reinterpret_cast<const double2*>(G + (size_t)k * ldg + 2 * lane);
        *reinterpret_cast<double2*>(S + k * DLD + 2 * lane) = v;
    }
}
// S[k][i] = G[i * ldg + k]: operand stored row-major with k contiguous.  A warp step covers 16 rows x 4 k (every 32-byte
// sector it touches is used in full; its 8-byte shared stores spread over all banks).
__device__ __forceinline__ void load_transposed(double* S, const double* G, size_t ldg) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ii = lane >> 1, kk = (lane & 1) * 2;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int u = warp + 8 * r;      // 32 units = 4 row groups x 8 k groups
        const int i = (u & 3) * 16 + ii, k = (u >> 2) * 4 + kk;
        const double2 v = *reinterpret_cast<const double2*>(G + (size_t)i * ldg + k);
        S[k * DLD + i] = v.x;
        S[(k + 1) * DLD + i] = v.y;
    }
}

struct TilePos {
    int wi, wj, g, t;
    __device__ TilePos() {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        wi = (warp >> 2) * 32;
        wj = (warp & 3) * 16;
        g = lane >> 2;
        t = lane & 3;
    }
    // accumulator (a, c, h) is element (row, col) of the 64 x 64 tile
    __device__ __forceinline__ int row(int a) const { return wi + a * 8 + g; }
    __device__ __forceinline__ int col(int c, int h) const { return wj + c * 8 + 2 * t + h; }
};

// acc[i][j] += sum_{k < DKC} As[k][i] Bs[k][j] for the staged chunks As, Bs (DKC x DLD, k-major)
__device__ __forceinline__ void mma_chunk(double (&acc)[4][2][2], const double* As, const double* Bs, const TilePos& tp) {
#pragma unroll
    for (int kk = 0; kk < DKC; kk += 4) {
        double af[4], bf[2];
        const double* ap = As + (kk + tp.t) * DLD + tp.wi + tp.g;
        const double* bp = Bs + (kk + tp.t) * DLD + tp.wj + tp.g;
#pragma unroll
        for (int a = 0; a < 4; ++a) af[a] = ap[a * 8];
#pragma unroll
        for (int c = 0; c < 2; ++c) bf[c] = bp[c * 8];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 2; ++c) dmma8(acc[a][c][0], acc[a][c][1], af[a], bf[c]);
    }
}

// acc[i][j] += sum_{k < klen} A(i, k) B(j, k).  TA: A is row-major [i][k] (A + i * lda + k), else k-major (A + k * lda + i);
// the same for B with j.  klen is a multiple of DKC.  As / Bs: DKC x DLD doubles each.
template <bool TA, bool TB>
__device__ __forceinline__ void tile_product(double (&acc)[4][2][2], const double* A, size_t lda, const double* B, size_t ldb,
                                             int klen, double* As, double* Bs, const TilePos& tp) {
    for (int k0 = 0; k0 < klen; k0 += DKC) {
        __syncthreads();      // the previous chunk has been consumed
        if (TA) load_transposed(As, A + k0, lda); else load_kmajor(As, A + (size_t)k0 * lda, lda);
        if (TB) load_transposed(Bs, B + k0, ldb); else load_kmajor(Bs, B + (size_t)k0 * ldb, ldb);
        __syncthreads();
        mma_chunk(acc, As, Bs, tp);
    }
}

__device__ __forceinline__ void zero_acc(double (&acc)[4][2][2]) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 2; ++c) acc[a][c][0] = acc[a][c][1] = 0.0;
}

// p -> (ti, tj), ti >= tj >= 0, p = ti (ti + 1) / 2 + tj
__device__ __forceinline__ void pair_decode(int p, int& ti, int& tj) {
    int i = (int)((sqrt(8.0 * (double)p + 1.0) - 1.0) * 0.5);
    while ((i + 1) * (i + 2) / 2 <= p) ++i;
    while (i * (i + 1) / 2 > p) --i;
    ti = i;
    tj = p - i * (i + 1) / 2;
}

}  // namespace gpet
