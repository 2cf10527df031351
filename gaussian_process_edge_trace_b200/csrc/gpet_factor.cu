// Factor of the posterior covariance, device provider ("throughput mode" of SURVEY.md H1).
//
// numpy's legacy multivariate_normal (called at sklearn_gpr.py:464) draws  Z @ (sqrt(s)[:,None] * Vt) + mean
// with u, s, Vt = svd(Sigma).  Sigma = U_r Mr U_r^T (gpet_posterior.cu), so with Mr = Q diag(d) Q^T:
// Vt = (U_r Q)^T, s = d.  This file: (1) batched symmetric eigensolver (parallel cyclic two-sided Jacobi in
// shared memory, one CTA per matrix, eigenvalues sorted descending like LAPACK's singular values);
// (2) assembly of A = diag(sqrt(d)) Q^T U_r^T with the canonical sign rule <Vt[k], w> > 0.
#include "gpet_common.cuh"

namespace gpet {

constexpr int J_MAX_SWEEPS = 40;
constexpr double J_REL_TOL = 1e-16;     // |a_pq| <= J_REL_TOL sqrt(a_pp a_qq): converged to rounding level

// Round-robin ("chess tournament") ordering: rp/2 disjoint pairs per round, rp-1 rounds per sweep.  All rotations of
// a round commute, so a round is:  (A) rp/2 threads compute (c, s) from the current matrix;  (B) every 2x2 block
// A[{p,q}][{r,s}] of (row pair P, column pair R >= P) is replaced by J_P^T * block * J_R in one step and mirrored,
// and the eigenvector columns {r,s} are rotated by J_R.  Two barriers per round.
__global__ void __launch_bounds__(1024)
jacobi_eig_kernel(double* __restrict__ Mr, int rp, double* __restrict__ d_out, double* __restrict__ Q_out,
                  int32_t* __restrict__ sweeps_out, int max_sweeps) {
    extern __shared__ double sm[];
    const int ld = rp + 1;  // odd leading dimension: conflict-free column walks
    double* A = sm;                    // rp x ld
    double* Q = A + (size_t)rp * ld;   // rp x ld
    double* cs = Q + (size_t)rp * ld;  // rp/2 x 2 (c, s)
    int* pq = (int*)(cs + rp);         // rp/2 x 2 (p, q)
    int* rank = pq + rp;               // rp
    __shared__ int n_rot;
    __shared__ double dmax_s;
    const int JT = blockDim.x;
    const int b = blockIdx.x, tid = threadIdx.x;
    const double* src = Mr + (size_t)b * rp * rp;
    for (int p = tid; p < rp * rp; p += JT) {
        int i = p / rp, j = p - i * rp;
        // symmetrise on load (the producer is symmetric up to rounding)
        A[i * ld + j] = 0.5 * (src[p] + src[(size_t)j * rp + i]);
        Q[i * ld + j] = (i == j) ? 1.0 : 0.0;
    }
    __syncthreads();
    if (tid == 0) {
        double mx = 0.0;
        for (int i = 0; i < rp; ++i) mx = fmax(mx, fabs(A[i * ld + i]));
        dmax_s = mx;
        n_rot = 0;
    }
    __syncthreads();
    const double abs_floor = 1e-20 * dmax_s;
    const int half = rp / 2, nm1 = rp - 1;
    // Work items of a round are fixed positions of the pairing, not matrix rows: (P, R >= P) block pairs for the
    // two-sided update of A and (R, row) for the eigenvector columns; both are dealt out flat over the threads so
    // that every thread has the same number of items (a warp-per-P mapping leaves half of the lanes idle).
    const int n_blk = half * (half + 1) / 2;
    constexpr int MAX_BLK = 8;                     // 64 * 65 / 2 block pairs (rp = 128) over >= 256 threads: <= 9 each
    int blkP[MAX_BLK + 1], blkR[MAX_BLK + 1];
    int n_mine = 0;
    for (int t = tid; t < n_blk && n_mine <= MAX_BLK; t += JT) {
        int R = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
        if ((R + 1) * (R + 2) / 2 <= t) ++R;
        if (R * (R + 1) / 2 > t) --R;
        blkR[n_mine] = R;
        blkP[n_mine] = t - R * (R + 1) / 2;
        ++n_mine;
    }
    int sweep = 0;
    for (; sweep < max_sweeps; ++sweep) {
        for (int round = 0; round < nm1; ++round) {
            if (tid < half) {
                int p, q;
                if (tid == 0) {
                    p = nm1;
                    q = round;
                } else {
                    p = (round + tid) % nm1;
                    q = (round - tid + nm1) % nm1;
                }
                if (p > q) { int t = p; p = q; q = t; }
                const double app = A[p * ld + p], aqq = A[q * ld + q], apq = A[p * ld + q];
                double c = 1.0, s = 0.0;
                const double aabs = fabs(apq);
                // rotate while the off-diagonal entry is above rounding level relative to its diagonal pair
                if (aabs > abs_floor && aabs > J_REL_TOL * sqrt(fabs(app) * fabs(aqq))) {
                    const double tau = (aqq - app) / (2.0 * apq);
                    const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(fma(tau, tau, 1.0)));
                    c = rsqrt(fma(t, t, 1.0));
                    s = t * c;
                    atomicAdd(&n_rot, 1);
                }
                cs[2 * tid] = c;
                cs[2 * tid + 1] = s;
                pq[2 * tid] = p;
                pq[2 * tid + 1] = q;
            }
            __syncthreads();
            // 2x2 blocks of A, upper half P <= R only (A is symmetric; the mirror block is written too, which also
            // keeps A exactly symmetric)
            for (int k = 0; k < n_mine; ++k) {
                const int P = blkP[k], R = blkR[k];
                const double c1 = cs[2 * P], s1 = cs[2 * P + 1];
                const double c2 = cs[2 * R], s2 = cs[2 * R + 1];
                if (s1 == 0.0 && s2 == 0.0) continue;
                const int p = pq[2 * P], q = pq[2 * P + 1];
                const int r = pq[2 * R], s = pq[2 * R + 1];
                const double apr = A[p * ld + r], aps = A[p * ld + s], aqr = A[q * ld + r], aqs = A[q * ld + s];
                // rows: J_P^T
                const double bpr = c1 * apr - s1 * aqr, bqr = s1 * apr + c1 * aqr;
                const double bps = c1 * aps - s1 * aqs, bqs = s1 * aps + c1 * aqs;
                // columns: J_R
                const double npr = c2 * bpr - s2 * bps, nps = s2 * bpr + c2 * bps;
                const double nqr = c2 * bqr - s2 * bqs, nqs = s2 * bqr + c2 * bqs;
                if (P == R) {   // the annihilated pair: exact zero off the diagonal
                    A[p * ld + p] = npr;
                    A[q * ld + q] = nqs;
                    A[p * ld + q] = 0.0;
                    A[q * ld + p] = 0.0;
                } else {
                    A[p * ld + r] = npr; A[r * ld + p] = npr;
                    A[p * ld + s] = nps; A[s * ld + p] = nps;
                    A[q * ld + r] = nqr; A[r * ld + q] = nqr;
                    A[q * ld + s] = nqs; A[s * ld + q] = nqs;
                }
            }
            // eigenvector columns {r, s} <- columns * J_R: items (R, row) flat over the threads, consecutive lanes walk
            // down the rows (stride ld, odd: conflict free)
            for (int t = tid; t < half * rp; t += JT) {
                const int R = t / rp, i = t - R * rp;
                const double c2 = cs[2 * R], s2 = cs[2 * R + 1];
                if (s2 == 0.0) continue;
                const int r = pq[2 * R], s = pq[2 * R + 1];
                const double qr = Q[i * ld + r], qs = Q[i * ld + s];
                Q[i * ld + r] = c2 * qr - s2 * qs;
                Q[i * ld + s] = s2 * qr + c2 * qs;
            }
            __syncthreads();
        }
        const int rot = n_rot;
        __syncthreads();
        if (tid == 0) n_rot = 0;
        if (rot == 0) break;
        __syncthreads();
    }
    // sort descending (rank by counting; ties broken by index => deterministic)
    __syncthreads();
    for (int k = tid; k < rp; k += JT) {
        const double dk = A[k * ld + k];
        int r = 0;
        for (int l = 0; l < rp; ++l) {
            const double dl = A[l * ld + l];
            r += (dl > dk) || (dl == dk && l < k);
        }
        rank[k] = r;
        d_out[(size_t)b * rp + r] = dk;
    }
    __syncthreads();
    double* Qo = Q_out + (size_t)b * rp * rp;
    for (int p = tid; p < rp * rp; p += JT) {
        int i = p / rp, k = p - i * rp;
        Qo[(size_t)i * rp + rank[k]] = Q[i * ld + k];
    }
    if (tid == 0) sweeps_out[b] = sweep;
}

// ---------------------------------------------------------------------------------------------------------------------
// Householder tridiagonalisation + implicit-shift QL (the EISPACK tred2 / tql2 pair), one CTA of 128 threads per matrix.
// ~9 rp^3 flops instead of the ~60 rp^3 of eight Jacobi sweeps, and the matrix is touched O(rp) times instead of
// O(rp) times PER SWEEP with scattered 2x2 accesses.  Parallel structure:
//   tred2: per Householder step the scale / norm / u^T p reductions are block reductions, the matrix-vector product is
//          one row-and-column walk per thread, the rank-2 update is dealt out flat over the lower triangle; the
//          accumulation of the reflectors gives every thread its own column;
//   tql2:  per QL sweep one thread runs the scalar (c, s) recurrence (division free: one rsqrt per rotation) and
//          stores the rotations, then every thread applies the whole rotation sequence to its own row of Z.
// The eigenvalues are absolutely accurate (eps * ||M||); the sampler scales eigenvector k by sqrt(d_k), so components
// below 1e-15 * lambda_max are irrelevant at the 1e-6 bar (they are the reference's own null-space noise).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int TQ_T = 128;

// NT threads per matrix: 128 for n <= 128 (the sizes this solver was tuned on; their summation order is unchanged), 256
// for larger matrices - several steps below give thread j row / column j, so NT >= n is required.
template <int NT>
__device__ __forceinline__ double tq_block_sum(double v, double* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = (red[0] + red[1]) + (red[2] + red[3]);
    if (NT == 256) s += (red[4] + red[5]) + (red[6] + red[7]);
    return s;
}

template <int NT>
__global__ void __launch_bounds__(NT)
tridiag_reduce_kernel(const double* __restrict__ Mr, int n, double* __restrict__ d_ws, double* __restrict__ e_ws,
                      double* __restrict__ Q_out) {
    extern __shared__ double sm[];
    const int ld = n | 1;
    double* a = sm;                         // n x ld: matrix -> Householder vectors -> Q -> eigenvectors
    double* d = a + (size_t)n * ld;         // n
    double* e = d + n;                      // n
    __shared__ double red[8];
    __shared__ double sc_s[4];              // scalars broadcast by thread 0
    const int b = blockIdx.x, tid = threadIdx.x;
    const double* src = Mr + (size_t)b * n * n;
    for (int p = tid; p < n * n; p += NT) {
        const int i = p / n, j = p - i * n;
        a[i * ld + j] = 0.5 * (src[p] + src[(size_t)j * n + i]);     // symmetrise on load
    }
    __syncthreads();
    // ---- tred2 -----------------------------------------------------------------------------------------------------
    for (int i = n - 1; i >= 1; --i) {
        const int l = i - 1;
        double h = 0.0;
        if (l > 0) {
            const double scale = tq_block_sum<NT>(tid <= l ? fabs(a[i * ld + tid]) : 0.0, red);
            if (scale == 0.0) {
                if (tid == 0) e[i] = a[i * ld + l];
            } else {
                double v = 0.0;
                if (tid <= l) {
                    v = a[i * ld + tid] / scale;
                    a[i * ld + tid] = v;
                }
                h = tq_block_sum<NT>(v * v, red);
                if (tid == 0) {
                    const double f = a[i * ld + l];
                    const double g = (f >= 0.0) ? -sqrt(h) : sqrt(h);
                    e[i] = scale * g;
                    sc_s[0] = h - f * g;
                    a[i * ld + l] = f - g;
                }
                __syncthreads();
                h = sc_s[0];
                // p = A u / H  (lower triangle of A: row j up to the diagonal, then down column j)
                double ej = 0.0;
                if (tid <= l) {
                    const int j = tid;
                    a[j * ld + i] = a[i * ld + j] / h;
                    double g0 = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0;     // four independent chains hide the FMA/LDS latency
                    const double* rj = a + j * ld;
                    const double* ri = a + i * ld;
                    int k = 0;
                    for (; k + 3 <= j; k += 4) {
                        g0 = fma(rj[k], ri[k], g0);
                        g1 = fma(rj[k + 1], ri[k + 1], g1);
                        g2 = fma(rj[k + 2], ri[k + 2], g2);
                        g3 = fma(rj[k + 3], ri[k + 3], g3);
                    }
                    for (; k <= j; ++k) g0 = fma(rj[k], ri[k], g0);
                    k = j + 1;
                    for (; k + 3 <= l; k += 4) {
                        g0 = fma(a[k * ld + j], ri[k], g0);
                        g1 = fma(a[(k + 1) * ld + j], ri[k + 1], g1);
                        g2 = fma(a[(k + 2) * ld + j], ri[k + 2], g2);
                        g3 = fma(a[(k + 3) * ld + j], ri[k + 3], g3);
                    }
                    for (; k <= l; ++k) g1 = fma(a[k * ld + j], ri[k], g1);
                    g0 += g2;
                    g1 += g3;
                    ej = (g0 + g1) / h;
                }
                const double f = tq_block_sum<NT>(tid <= l ? ej * a[i * ld + tid] : 0.0, red);
                const double hh = f / (h + h);
                if (tid <= l) e[tid] = ej - hh * a[i * ld + tid];      // q = p - K u
                __syncthreads();
                // A -= u q^T + q u^T on the lower triangle 0 <= k <= j <= l: one warp per row j, lanes across k (the flat
                // index over the triangle cost a square root and two fix-ups per element: 39 % of the kernel's instructions)
                {
                    const double* ui = a + i * ld;
                    for (int j = l - (tid >> 5); j >= 0; j -= NT / 32) {      // longest rows first
                        const double uj = ui[j], ej2 = e[j];
                        double* rj = a + j * ld;
                        for (int k = tid & 31; k <= j; k += 32) rj[k] -= uj * e[k] + ej2 * ui[k];
                    }
                }
            }
        } else {
            if (tid == 0) e[i] = a[i * ld + l];
        }
        if (tid == 0) d[i] = h;
        __syncthreads();
    }
    if (tid == 0) { d[0] = 0.0; e[0] = 0.0; }
    __syncthreads();
    // accumulate the reflectors: thread j owns column j
    for (int i = 0; i < n; ++i) {
        const int l = i - 1;
        if (d[i] != 0.0 && tid <= l) {
            const int j = tid;
            double g0 = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0;
            const double* ri = a + i * ld;
            int k = 0;
            for (; k + 3 <= l; k += 4) {
                g0 = fma(ri[k], a[k * ld + j], g0);
                g1 = fma(ri[k + 1], a[(k + 1) * ld + j], g1);
                g2 = fma(ri[k + 2], a[(k + 2) * ld + j], g2);
                g3 = fma(ri[k + 3], a[(k + 3) * ld + j], g3);
            }
            for (; k <= l; ++k) g0 = fma(ri[k], a[k * ld + j], g0);
            const double g = (g0 + g1) + (g2 + g3);
#pragma unroll 4
            for (k = 0; k <= l; ++k) a[k * ld + j] = fma(-g, a[k * ld + i], a[k * ld + j]);
        }
        __syncthreads();
        if (tid == 0) {
            d[i] = a[i * ld + i];
            a[i * ld + i] = 1.0;
        }
        if (tid <= l) {
            a[tid * ld + i] = 0.0;
            a[i * ld + tid] = 0.0;
        }
        __syncthreads();
    }
    // tridiagonal matrix (off-diagonal shifted down by one, tql2 convention) and Q = product of the reflectors
    for (int i = tid; i < n; i += NT) {
        d_ws[(size_t)b * n + i] = d[i];
        e_ws[(size_t)b * n + i] = (i + 1 < n) ? e[i + 1] : 0.0;
    }
    double* Qo = Q_out + (size_t)b * n * n;
    for (int p = tid; p < n * n; p += NT) {
        const int i = p / n, j = p - i * n;
        Qo[p] = a[i * ld + j];
    }
}

// QL with implicit shifts on the tridiagonal matrix: ONE WARP per matrix.  Lane 0 runs the scalar recurrence (it does not
// depend on the eigenvectors), all lanes search for the deflation point.  Every rotation (c, s) is appended to a list in
// global memory together with one (m, lo) header per sweep; tridiag_apply_kernel replays the list on the eigenvectors.
constexpr int QL_WARPS = 4;

__global__ void __launch_bounds__(QL_WARPS * 32)
tridiag_ql_kernel(int B, int n, double* __restrict__ d_ws, const double* __restrict__ e_ws, double2* __restrict__ rot,
                  int2* __restrict__ hdr, int32_t* __restrict__ counts, int cap_rot, int cap_sweeps) {
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * QL_WARPS + warp;
    if (b >= B) return;
    double* d = sm + (size_t)warp * 2 * n;
    double* e = d + n;
    for (int i = lane; i < n; i += 32) {
        d[i] = d_ws[(size_t)b * n + i];
        e[i] = e_ws[(size_t)b * n + i];
    }
    __syncwarp();
    double2* rb = rot + (size_t)b * cap_rot;
    int2* hb = hdr + (size_t)b * cap_sweeps;
    int n_rot = 0, n_sw = 0;
    bool fail = false;
    for (int l = 0; l < n && !fail; ++l) {
        for (int guard = 0;; ++guard) {
            if (guard == 64) { fail = true; break; }         // tql2 gives up after 30
            // first m >= l whose off-diagonal entry is negligible (m = n-1 always qualifies)
            int m = n - 1;
            for (int base = l; base < n - 1; base += 32) {
                const int mm = base + lane;
                bool neg = false;
                if (mm < n - 1) {
                    const double dd = fabs(d[mm]) + fabs(d[mm + 1]);
                    neg = (fabs(e[mm]) + dd == dd);
                }
                const unsigned bal = __ballot_sync(0xffffffffu, neg);
                if (bal) { m = base + __ffs(bal) - 1; break; }
            }
            if (m == l) break;
            if (n_sw >= cap_sweeps || n_rot + (m - l) > cap_rot) { fail = true; break; }
            if (lane == 0) {
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = sqrt(fma(g, g, 1.0));
                g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? fabs(r) : -fabs(r)));
                double s = 1.0, c = 1.0, p = 0.0;
                int i = m - 1, k = n_rot;
                bool under = false;
                for (; i >= l; --i) {
                    const double f = s * e[i], bb = c * e[i];
                    // r = hypot(f, g), s = f / r, c = g / r from ONE reciprocal square root (MUFU seed + a cubic
                    // Newton step), no division: this recurrence is the serial critical path of the whole solver.
                    // h2 == 0 (both operands below 1e-154: only in converged, negligible tails) takes tql2's underflow
                    // recovery branch.
                    const double h2 = fma(f, f, g * g);
                    if (!(h2 > 2.3e-308)) {
                        e[i + 1] = 0.0;
                        d[i + 1] -= p;
                        e[m] = 0.0;
                        under = true;
                        break;
                    }
                    double inv;
                    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(inv) : "d"(h2));
                    const double er = fma(-h2 * inv, inv, 1.0);
                    inv = fma(inv * er, fma(0.375, er, 0.5), inv);
                    e[i + 1] = h2 * inv;
                    s = f * inv;
                    c = g * inv;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * s + 2.0 * c * bb;
                    p = s * r;
                    d[i + 1] = g + p;
                    g = c * r - bb;
                    rb[k++] = make_double2(c, s);             // rotation of columns (i, i+1), in order i = m-1 .. lo
                }
                if (!under) {
                    d[l] -= p;
                    e[l] = g;
                    e[m] = 0.0;
                }
                hb[n_sw] = make_int2(m, under ? i + 1 : l);   // (m, lo): exactly m - lo rotations were appended
                n_rot = k;
            }
            n_rot = __shfl_sync(0xffffffffu, n_rot, 0);
            ++n_sw;
        }
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) d_ws[(size_t)b * n + i] = d[i];
    if (lane == 0) counts[b] = fail ? -1 : n_sw;
}

// Replays the rotation list on the rows of Q (one thread per row, rotations broadcast from L1), then sorts the
// eigenvalues descending (rank by counting; ties broken by index => deterministic) and writes d and Q.
constexpr int AP_STAGE = 512, AP_HDR = 32;

template <int NT>
__global__ void __launch_bounds__(NT)
tridiag_apply_kernel(int n, const double* __restrict__ d_ws, const double2* __restrict__ rot, const int2* __restrict__ hdr,
                     const int32_t* __restrict__ counts, int cap_rot, int cap_sweeps, double* __restrict__ d_out,
                     double* __restrict__ Q, int32_t* __restrict__ iters_out) {
    extern __shared__ double sm[];
    const int ld = n | 1;
    double* a = sm;                         // n x ld
    double* d = a + (size_t)n * ld;         // n
    int* rank = (int*)(d + n);              // n (+ pad), then the rotation staging buffer
    __shared__ int grp_s[2];
    const int b = blockIdx.x, tid = threadIdx.x;
    double* Qb = Q + (size_t)b * n * n;
    for (int p = tid; p < n * n; p += NT) {
        const int i = p / n, j = p - i * n;
        a[i * ld + j] = Qb[p];
    }
    for (int i = tid; i < n; i += NT) d[i] = d_ws[(size_t)b * n + i];
    __syncthreads();
    const int n_sw = counts[b];
    {
        // sweeps are replayed in groups: the group's rotations are staged in shared memory by all threads (coalesced),
        // then every row thread consumes them from there (a dependent global load per rotation costs ~250 cycles)
        double2* stage = reinterpret_cast<double2*>(rank + ((n + 3) & ~3));      // AP_STAGE rotations, 16-byte aligned
        __shared__ int2 hs[AP_HDR];
        const double2* rb = rot + (size_t)b * cap_rot;
        const int2* hb = hdr + (size_t)b * cap_sweeps;
        double* z = a + (size_t)(tid < n ? tid : 0) * ld;
        int t0 = 0, k0 = 0;
        while (t0 < n_sw) {
            // group = as many whole sweeps as fit AP_STAGE rotations / AP_HDR headers (a sweep has < n <= 256 rotations)
            __syncthreads();
            if (tid == 0) {
                int cnt = 0, t1 = t0;
                while (t1 < n_sw && t1 - t0 < AP_HDR) {
                    const int2 h = hb[t1];
                    const int c = max(h.x - h.y, 0);
                    if (cnt + c > AP_STAGE) break;
                    hs[t1 - t0] = h;
                    cnt += c;
                    ++t1;
                }
                grp_s[0] = t1 - t0;
                grp_s[1] = cnt;
            }
            __syncthreads();
            const int ns = grp_s[0], cnt = grp_s[1];
            for (int k = tid; k < cnt; k += NT) stage[k] = rb[k0 + k];
            __syncthreads();
            if (tid < n) {
                int k = 0;
                for (int t2 = 0; t2 < ns; ++t2) {
                    const int m = hs[t2].x, lo = hs[t2].y;
                    if (m - 1 < lo) continue;
                    double hi = z[m];
#pragma unroll 4
                    for (int i = m - 1; i >= lo; --i, ++k) {
                        const double2 cs = stage[k];
                        const double zi = z[i];
                        z[i + 1] = fma(cs.y, zi, cs.x * hi);
                        hi = fma(cs.x, zi, -cs.y * hi);
                    }
                    z[lo] = hi;
                }
            }
            t0 += ns;
            k0 += cnt;
        }
    }
    __syncthreads();
    for (int k = tid; k < n; k += NT) {
        const double dk = d[k];
        int r = 0;
        for (int l2 = 0; l2 < n; ++l2) {
            const double dl = d[l2];
            r += (dl > dk) || (dl == dk && l2 < k);
        }
        rank[k] = r;
        d_out[(size_t)b * n + r] = dk;
    }
    __syncthreads();
    for (int p = tid; p < n * n; p += NT) {
        const int i = p / n, k = p - i * n;
        Qb[(size_t)i * n + rank[k]] = a[i * ld + k];
    }
    if (tid == 0) iters_out[b] = n_sw;
}

// A[b][k][j] = sign_k sqrt(max(d_k, 0)) sum_i Q[i][k] Ur[j][i]
template <int AS_TJ>      // grid columns per CTA: 64 (8 warps x 8 columns) or 32 (ranks whose Q + 64 columns exceed shared memory)
__global__ void __launch_bounds__(256)
factor_assemble_kernel(const double* __restrict__ d, const double* __restrict__ Q, const double* __restrict__ Ur,
                       const double* __restrict__ uw, int rp, int n, double* __restrict__ A) {
    extern __shared__ double sm[];
    double* Qs = sm;                          // rp x rp  [i][k]
    double* Us = Qs + (size_t)rp * rp;        // AS_TJ x (rp+1)  [jj][i]
    double* scale = Us + (size_t)AS_TJ * (rp + 1);  // rp
    const int b = blockIdx.y, j0 = blockIdx.x * AS_TJ, tid = threadIdx.x;
    const double* Qb = Q + (size_t)b * rp * rp;
    for (int p = tid; p < rp * rp; p += 256) Qs[p] = Qb[p];
    for (int p = tid; p < AS_TJ * rp; p += 256) {
        int jj = p / rp, i = p - jj * rp;
        Us[jj * (rp + 1) + i] = (j0 + jj < n) ? Ur[(size_t)(j0 + jj) * rp + i] : 0.0;
    }
    __syncthreads();
    for (int k = tid; k < rp; k += 256) {
        double t = 0.0;
        for (int i = 0; i < rp; ++i) t = fma(Qs[i * rp + k], uw[i], t);
        const double dk = d[(size_t)b * rp + k];
        scale[k] = (t < 0.0 ? -1.0 : 1.0) * sqrt(dk > 0.0 ? dk : 0.0);
    }
    __syncthreads();
    // C[k][jj] = sum_i Qs[i][k] Us[jj][i] on the fp64 tensor instruction: m = k (8 per tile), n = jj (warp w owns the
    // 8 columns 8w .. 8w+7), K = i in steps of 4; two shared-memory loads per DMMA instead of two per scalar FMA
    double* Ab = A + (size_t)b * rp * n;
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int nks = rp >> 2;                      // rp is a multiple of 4
    if (8 * warp >= AS_TJ) return;                // narrow tile: the upper warps only helped with the loads
    const double* up = Us + (size_t)(8 * warp + g) * (rp + 1) + t4;
    for (int mt = 0; mt * 8 < rp; ++mt) {
        const int ka = mt * 8 + g;
        const bool kok = ka < rp;
        const double* qp = Qs + (size_t)t4 * rp + (kok ? ka : 0);
        double c0 = 0.0, c1 = 0.0;
#pragma unroll 4
        for (int ks = 0; ks < nks; ++ks) {
            const double av = kok ? qp[(size_t)(4 * ks) * rp] : 0.0;
            const double bv = up[4 * ks];
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c0), "+d"(c1)
                         : "d"(av), "d"(bv));
        }
        const int j = j0 + 8 * warp + 2 * t4;
        if (kok) {
            const double sc = scale[ka];
            if (j < n) Ab[(size_t)ka * n + j] = sc * c0;
            if (j + 1 < n) Ab[(size_t)ka * n + j + 1] = sc * c1;
        }
    }
}

}  // namespace gpet

using namespace gpet;

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" int64_t gpet_sym_eig_workspace_bytes(int B, int rp) {
    const size_t cap_rot = 2 * (size_t)rp * rp, cap_sw = 8 * (size_t)rp;
    return (int64_t)(2 * align256((size_t)B * rp * 8) + align256((size_t)B * 4) + align256((size_t)B * cap_sw * 8) +
                     align256((size_t)B * cap_rot * 16) + 256);
}

extern "C" int gpet_sym_eig_f64(double* Mr, int B, int rp, double* d, double* Q, int32_t* sweeps, void* work, void* stream) {
    return gpet::sym_eig_run(Mr, B, rp, d, Q, sweeps, work, stream, g_tune[GPET_TUNE_EIG_THREADS]);
}

// jt == 0: Householder + QL; otherwise threads per CTA of the parallel cyclic Jacobi kernel
int gpet::sym_eig_run(double* Mr, int B, int rp, double* d, double* Q, int32_t* sweeps, void* work, void* stream, int jt) {
    GPET_REQUIRE(Mr && d && Q && sweeps && B > 0, "gpet_sym_eig_f64: bad argument");
    GPET_SUPPORTED(rp >= 2 && (rp % 2) == 0 && rp <= GPET_MAX_RANK, "gpet_sym_eig_f64: rp=%d must be even and <= %d", rp,
                   GPET_MAX_RANK);
    cudaError_t e;
    if (jt == 0) {      // Householder + QL (default): reduce -> serial QL recurrences (one warp each) -> replay
        GPET_REQUIRE(work != nullptr, "gpet_sym_eig_f64: workspace required (gpet_sym_eig_workspace_bytes)");
        const int cap_rot = 2 * rp * rp, cap_sw = 8 * rp;
        char* w = (char*)work;
        double* d_ws = (double*)w;                 w += align256((size_t)B * rp * 8);
        double* e_ws = (double*)w;                 w += align256((size_t)B * rp * 8);
        int32_t* counts = (int32_t*)w;             w += align256((size_t)B * 4);
        int2* hdr = (int2*)w;                      w += align256((size_t)B * cap_sw * 8);
        double2* rot = (double2*)w;
        const size_t smem_r = ((size_t)rp * (rp | 1) + 2 * (size_t)rp) * sizeof(double);
        const size_t smem_a = ((size_t)rp * (rp | 1) + (size_t)rp) * sizeof(double) + (size_t)((rp + 3) & ~3) * sizeof(int) +
                              (size_t)AP_STAGE * sizeof(double2);
        GPET_SUPPORTED(smem_r <= 227 * 1024 && smem_a <= 227 * 1024, "gpet_sym_eig_f64: rp=%d needs %zu B shared memory", rp,
                       smem_r > smem_a ? smem_r : smem_a);
        const bool wide = rp > 128;           // thread j owns row / column j in several steps: NT >= rp
        GPET_SUPPORTED(rp <= 256, "gpet_sym_eig_f64: rp=%d > 256", rp);
        e = wide ? cudaFuncSetAttribute(tridiag_reduce_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r)
                 : cudaFuncSetAttribute(tridiag_reduce_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r);
        if (e == cudaSuccess)
            e = wide ? cudaFuncSetAttribute(tridiag_apply_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a)
                     : cudaFuncSetAttribute(tridiag_apply_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a);
        if (e != cudaSuccess) {
            set_error("tridiag smem attribute: %s", cudaGetErrorString(e));
            return GPET_ERR_CUDA;
        }
        cudaStream_t st = (cudaStream_t)stream;
        if (wide) tridiag_reduce_kernel<256><<<B, 256, smem_r, st>>>(Mr, rp, d_ws, e_ws, Q);
        else tridiag_reduce_kernel<128><<<B, 128, smem_r, st>>>(Mr, rp, d_ws, e_ws, Q);
        tridiag_ql_kernel<<<(B + QL_WARPS - 1) / QL_WARPS, QL_WARPS * 32, (size_t)QL_WARPS * 2 * rp * sizeof(double), st>>>(
            B, rp, d_ws, e_ws, rot, hdr, counts, cap_rot, cap_sw);
        if (wide) tridiag_apply_kernel<256><<<B, 256, smem_a, st>>>(rp, d_ws, rot, hdr, counts, cap_rot, cap_sw, d, Q, sweeps);
        else tridiag_apply_kernel<128><<<B, 128, smem_a, st>>>(rp, d_ws, rot, hdr, counts, cap_rot, cap_sw, d, Q, sweeps);
        return check_launch("tridiag_eig kernels");
    }
    const int jt_in = jt;
    jt &= 0xffff;
    jt = jt < 256 ? 256 : (jt > 1024 ? 1024 : (jt / 32) * 32);     // >= 256: see MAX_BLK in the kernel
    const size_t smem = (2 * (size_t)rp * (rp + 1) + rp) * sizeof(double) + 2 * (size_t)rp * sizeof(int);
    GPET_SUPPORTED(smem <= 227 * 1024, "gpet_sym_eig_f64: the Jacobi solver needs %zu B shared memory at rp=%d", smem, rp);
    e = cudaFuncSetAttribute(jacobi_eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("jacobi smem attribute: %s", cudaGetErrorString(e));
        return GPET_ERR_CUDA;
    }
    // jt >> 16 (block Jacobi pivots only): cap on the inner sweeps - the pivot is then rotated towards diagonal form, not
    // diagonalised to rounding level, and the caller must not assume Q^T M Q = diag(d)
    const int cap = (jt_in >> 16) > 0 ? (jt_in >> 16) : J_MAX_SWEEPS;
    jacobi_eig_kernel<<<B, jt, smem, (cudaStream_t)stream>>>(Mr, rp, d, Q, sweeps, cap);
    return check_launch("jacobi_eig_kernel");
}

extern "C" int gpet_factor_assemble_f64(const double* d, const double* Q, const double* Ur, const double* uw, int B, int rp,
                                        int n, double* A, void* stream) {
    GPET_REQUIRE(d && Q && Ur && uw && A && B > 0 && n > 0, "gpet_factor_assemble_f64: bad argument");
    GPET_SUPPORTED(rp >= 4 && (rp % 4) == 0 && rp <= GPET_MAX_RANK, "gpet_factor_assemble_f64: rp=%d must be a multiple of 4, <= %d",
                   rp, GPET_MAX_RANK);
    cudaStream_t st = (cudaStream_t)stream;
    auto smem_for = [rp](int tj) { return ((size_t)rp * rp + (size_t)tj * (rp + 1) + rp) * sizeof(double); };
    if (smem_for(64) <= 227 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(factor_assemble_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(64));
        if (e != cudaSuccess) {
            set_error("assemble smem attribute: %s", cudaGetErrorString(e));
            return GPET_ERR_CUDA;
        }
        dim3 grid((n + 63) / 64, B);
        factor_assemble_kernel<64><<<grid, 256, smem_for(64), st>>>(d, Q, Ur, uw, rp, n, A);
    } else {
        GPET_SUPPORTED(smem_for(32) <= 227 * 1024, "gpet_factor_assemble_f64: rp=%d needs %zu B shared memory", rp, smem_for(32));
        cudaError_t e = cudaFuncSetAttribute(factor_assemble_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(32));
        if (e != cudaSuccess) {
            set_error("assemble smem attribute: %s", cudaGetErrorString(e));
            return GPET_ERR_CUDA;
        }
        dim3 grid((n + 31) / 32, B);
        factor_assemble_kernel<32><<<grid, 256, smem_for(32), st>>>(d, Q, Ur, uw, rp, n, A);
    }
    return check_launch("factor_assemble_kernel");
}
