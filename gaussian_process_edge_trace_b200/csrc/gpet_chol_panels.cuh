// Panel Cholesky of a lower triangle held in shared memory, shared by the posterior kernels (gpet_posterior.cu: the whole
// training matrix of a trace) and the blocked HBM factorisation (gpet_dense.cu: one 64 x 64 diagonal block at a time).
#pragma once
#include "gpet_common.cuh"

namespace gpet {

constexpr int PT = 256;  // threads per CTA

// Storage of the lower triangle of K / L in shared memory: full rows with an odd leading dimension (fast index, the
// default) or packed rows (half the memory: training sets up to GPET_MAX_TRAIN points).
struct FullLowerP {
    int ld;
    __device__ __forceinline__ int operator()(int i, int j) const { return i * ld + j; }
};
struct PackedLowerP {
    __device__ __forceinline__ int operator()(int i, int j) const { return ((i * (i + 1)) >> 1) + j; }
};

// Right-looking Cholesky (lower) in panels of PNB columns: three barriers per PANEL instead of per column.
//   (a) warp 0 factors the diagonal block in registers (every lane the same arithmetic from broadcast loads);
//   (b) one thread per row below solves its PNB entries against the block;
//   (c) one warp per trailing row, lanes across its columns: A_ij -= sum_k L_ik L_jk over the panel's columns.
// Every element receives exactly the fma sequence of the column-at-a-time form (updates from the columns to its left
// in ascending order, then the multiplication by 1 / L_kk), so the factor is bit-identical to it; that form cost
// 3 m barriers and a division per element of every rank-1 update (1.2 ms per 1250 traces at m ~ 100).
constexpr int PNB = 8;
template <class IX>
__device__ void cholesky_panels(int m, const IX ix, double* Ls, double* blk, int* flag) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int k0 = 0; k0 < m; k0 += PNB) {
        const int nb = min(PNB, m - k0), k1 = k0 + nb;
        if (warp == 0) {
            double D[PNB][PNB], iv[PNB];
#pragma unroll
            for (int r = 0; r < PNB; ++r)
#pragma unroll
                for (int cc = 0; cc <= r; ++cc) D[r][cc] = (r < nb) ? Ls[ix(k0 + r, k0 + cc)] : (r == cc ? 1.0 : 0.0);
            bool bad = false;
#pragma unroll
            for (int cc = 0; cc < PNB; ++cc) {
                double d = D[cc][cc];
#pragma unroll
                for (int k = 0; k < cc; ++k) d = fma(-D[cc][k], D[cc][k], d);
                if (!(d > 0.0)) { bad = true; d = 1.0; }
                const double sq = sqrt(d);
                D[cc][cc] = sq;
                iv[cc] = 1.0 / sq;
#pragma unroll
                for (int r = cc + 1; r < PNB; ++r) {
                    double v = D[r][cc];
#pragma unroll
                    for (int k = 0; k < cc; ++k) v = fma(-D[r][k], D[cc][k], v);
                    D[r][cc] = v * iv[cc];
                }
            }
            if (lane == 0) {
                if (bad) *flag = 1;
#pragma unroll
                for (int r = 0; r < PNB; ++r) {
#pragma unroll
                    for (int cc = 0; cc <= r; ++cc) {
                        blk[r * PNB + cc] = D[r][cc];
                        if (r < nb) Ls[ix(k0 + r, k0 + cc)] = D[r][cc];
                    }
                    blk[PNB * PNB + r] = iv[r];
                }
            }
        }
        __syncthreads();
        for (int r = k1 + tid; r < m; r += PT) {      // nb == PNB whenever rows exist below
            double a[PNB];
#pragma unroll
            for (int cc = 0; cc < PNB; ++cc) a[cc] = Ls[ix(r, k0 + cc)];
#pragma unroll
            for (int cc = 0; cc < PNB; ++cc) {
                double v = a[cc];
#pragma unroll
                for (int k = 0; k < cc; ++k) v = fma(-a[k], blk[cc * PNB + k], v);
                a[cc] = v * blk[PNB * PNB + cc];
                Ls[ix(r, k0 + cc)] = a[cc];
            }
        }
        __syncthreads();
        for (int i = k1 + warp; i < m; i += PT / 32) {
            double li[PNB];
#pragma unroll
            for (int kk = 0; kk < PNB; ++kk) li[kk] = -Ls[ix(i, k0 + kk)];
            for (int j = k1 + lane; j <= i; j += 32) {
                double v = Ls[ix(i, j)];
#pragma unroll
                for (int kk = 0; kk < PNB; ++kk) v = fma(li[kk], Ls[ix(j, k0 + kk)], v);
                Ls[ix(i, j)] = v;
            }
        }
        __syncthreads();
    }
}

}  // namespace gpet
