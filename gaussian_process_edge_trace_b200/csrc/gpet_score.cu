// Posterior-curve scoring (cost_funct) and top-N_keep selection.
//
// Reference seams: gpet.py:371-410 (cost_funct: bilinear gather of the gradient image along the curve,
// arc-length abscissa, two composite Simpson integrals; scipy _basic_simpson non-uniform branch),
// gpet.py:336-367 (forward difference), gpet.py:437-449 (loop over curves, argsort, keep N_keep).
//
// Layout: Y[b][j][s] with s contiguous => one THREAD per curve walks down j, a warp reads 32 consecutive
// doubles per column (coalesced 256 B).  The gradient image is read through a column-major float32 copy
// gradT[b][x][y], so the two bilinear taps of a point are adjacent and all curves of a CTA gather from
// the same 4*M-byte column at a time (L1/L2 resident).  Algorithmic HBM bytes per curve: 8 n + 8.
#include "gpet_common.cuh"
#include "gpet_npsum.cuh"

namespace gpet {


// floor of a double in [0, 2^31) without the conversion (XU) pipe: round-to-nearest through the 2^52 trick, then
// fix up; the integer falls out of the low mantissa word.
__device__ __forceinline__ double floor_nonneg(double x, int& i) {
    const double magic = 6755399441055744.0;  // 1.5 * 2^52
    double r = (x + magic) - magic;           // nearest integer (ties to even)
    if (r > x) r -= 1.0;
    i = __double2loint(r + magic);
    return r;
}

// float (>= 0, finite) -> double by integer arithmetic (ALU pipe instead of the conversion pipe)
__device__ __forceinline__ double f32_to_f64_nonneg(float v) {
    const unsigned int u = __float_as_uint(v);
    const unsigned int e = u >> 23;
    if (e == 0u) return (double)v;  // zero / subnormal: rare, take the slow path
    const unsigned int hi = (u >> 3) + (896u << 20);
    const unsigned int lo = u << 29;
    return __hiloint2double((int)hi, (int)lo);
}

// The two bilinear taps of one curve point, fetched early and combined later (FITPACK bispeu with kx=ky=1 on
// integer knots == clamped 2-tap lerp).
struct Taps {
    float g0, g1;
    double f;
};

__device__ __forceinline__ Taps fetch_taps(const float* __restrict__ col, double y, double ymax, double fmaxrow) {
    const double yc = fmin(fmax(y, 0.0), ymax);
    int i0;
    double fl = floor_nonneg(yc, i0);
    if (fl > fmaxrow) { fl = fmaxrow; i0 = (int)fmaxrow; }   // yc == M-1 exactly: rows M-2, M-1 with f = 1
    Taps t;
    t.f = yc - fl;
    t.g0 = __ldg(col + i0);
    t.g1 = __ldg(col + i0 + 1);
    return t;
}

__device__ __forceinline__ double finish_taps(const Taps& t) {
    const double g0 = f32_to_f64_nonneg(t.g0), g1 = f32_to_f64_nonneg(t.g1);
    return fma(t.f, g1 - g0, g0) + 1e-3;
}

__device__ __forceinline__ double simpson_term(double y0, double y1, double y2, double h0, double h1) {
    // hs/6 * ( y0 (2 - h1/h0) + y1 hs^2/(h0 h1) + y2 (2 - h0/h1) ) with a single reciprocal
    const double hs = h0 + h1, hp = h0 * h1;
    const double num = y0 * (fma(2.0, h0, -h1) * h1) + y1 * (hs * hs) + y2 * (fma(2.0, h1, -h0) * h0);
    return (hs * num) * __drcp_rn(6.0 * hp);
}

// THREADS curves per CTA (one thread each).  PIPE: the taps of the next Simpson pair are issued one iteration ahead
// of their use, on top of the curve values that are always prefetched two pairs ahead.
template <int THREADS, bool PIPE>
__global__ void __launch_bounds__(THREADS)
score_kernel(const double* __restrict__ Y, const float* __restrict__ gradT, int n, int S, int M, int N, int x_st,
             double* __restrict__ cost) {
    const int b = blockIdx.y;
    const int s = blockIdx.x * THREADS + threadIdx.x;
    if (s >= S) return;
    const size_t Sz = (size_t)S;
    const double* yp = Y + (size_t)b * n * Sz + s;
    const float* gt = gradT + ((size_t)b * N + x_st) * M;
    const double ymax = (double)(M - 1), fmaxrow = (double)(M - 2);
    // K = n - 1 Simpson samples (j = 0 .. n-2), K odd <=> n even; pairs p = 0 .. (K-1)/2 - 1
    const int P = (n - 2) / 2;
    double y0 = __ldg(yp), y1 = __ldg(yp + Sz);
    double pa = __ldg(yp + 2 * Sz), pb = __ldg(yp + 3 * Sz);
    double pc = 0.0, pd = 0.0;
    if (P > 1) { pc = __ldg(yp + 4 * Sz); pd = __ldg(yp + 5 * Sz); }
    double d = y1 - y0;
    double seg0 = sqrt(fma(d, d, 1.0));
    double t0 = seg0;  // cumsum abscissa of sample 0
    double g0 = finish_taps(fetch_taps(gt, y0, ymax, fmaxrow));
    double AL = 0.0, LI = 0.0;
    const double* ynext = yp + 6 * Sz;
    const float* col = gt + M;
    Taps ta, tb;
    if (PIPE) {
        ta = fetch_taps(col, y1, ymax, fmaxrow);
        tb = fetch_taps(col + M, pa, ymax, fmaxrow);
    }
    for (int p = 0; p < P; ++p) {
        const double y2 = pa, y3 = pb;
        pa = pc;
        pb = pd;
        if (p + 2 < P) { pc = __ldg(ynext); pd = __ldg(ynext + Sz); }
        ynext += 2 * Sz;
        Taps na, nb;
        if (PIPE) {
            if (p + 1 < P) {   // samples 2p+3 (y3) and 2p+4 (the new pa)
                na = fetch_taps(col + 2 * M, y3, ymax, fmaxrow);
                nb = fetch_taps(col + 3 * M, pa, ymax, fmaxrow);
            }
        } else {
            ta = fetch_taps(col, y1, ymax, fmaxrow);
            tb = fetch_taps(col + M, y2, ymax, fmaxrow);
        }
        d = y2 - y1;
        const double seg1 = sqrt(fma(d, d, 1.0));
        const double t1 = t0 + seg1;
        d = y3 - y2;
        const double seg2 = sqrt(fma(d, d, 1.0));
        const double t2 = t1 + seg2;
        const double g1 = finish_taps(ta);
        const double g2 = finish_taps(tb);
        col += 2 * M;
        LI += simpson_term(g0, g1, g2, t1 - t0, t2 - t1);
        AL += fma(4.0, seg1, seg0) + seg2;
        y1 = y3;
        seg0 = seg2;
        t0 = t2;
        g0 = g2;
        if (PIPE) { ta = na; tb = nb; }
    }
    cost[(size_t)b * S + s] = (AL * (2.0 / 6.0)) / LI;
}

template <int THREADS, bool PIPE>
static void launch_score(const double* Y, const float* gradT, int B, int n, int S, int M, int N, int x_st, double* cost,
                         cudaStream_t st) {
    dim3 grid((S + THREADS - 1) / THREADS, B);
    score_kernel<THREADS, PIPE><<<grid, THREADS, 0, st>>>(Y, gradT, n, S, M, N, x_st, cost);
}

// ---- top-N_keep: one CTA per trace, bitonic sort of (cost, index) in shared memory -----------------
constexpr int TK_THREADS = 512;

__device__ __forceinline__ bool pair_less(double ka, int ia, double kb, int ib) {
    return (ka < kb) || (ka == kb && ia < ib);
}

__global__ void __launch_bounds__(TK_THREADS)
topk_kernel(const double* __restrict__ cost, int S, int P2, int Kp, int32_t* __restrict__ idx_out,
            double* __restrict__ cost_out, double* __restrict__ wts_out) {
    extern __shared__ double sm[];
    double* key = sm;
    int* val = (int*)(sm + P2);
    const int b = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < P2; i += TK_THREADS) {
        double c = (i < S) ? cost[(size_t)b * S + i] : __longlong_as_double(0x7ff0000000000000LL);
        if (c != c) c = __longlong_as_double(0x7ff0000000000000LL);  // NaN sorts last
        key[i] = c;
        val[i] = (i < S) ? i : 0x7fffffff;
    }
    __syncthreads();
    for (int k = 2; k <= P2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < P2; i += TK_THREADS) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const bool up = ((i & k) == 0);
                    const double ka = key[i], kb = key[ixj];
                    const int ia = val[i], ib = val[ixj];
                    const bool sw = up ? pair_less(kb, ib, ka, ia) : pair_less(ka, ia, kb, ib);
                    if (sw) {
                        key[i] = kb; key[ixj] = ka;
                        val[i] = ib; val[ixj] = ia;
                    }
                }
            }
            __syncthreads();
        }
    }
    // weights (gpet.py:492-493): inv = 1/cost; w = inv / np.sum(inv)   (numpy pairwise sum, one thread)
    double* inv = sm + P2 + (P2 + 1) / 2;
    for (int c = tid; c < Kp; c += TK_THREADS) inv[c] = 1.0 / key[c];
    __syncthreads();
    __shared__ double total;
    if (tid == 0) total = np_pairwise_sum(inv, Kp);
    __syncthreads();
    for (int c = tid; c < Kp; c += TK_THREADS) {
        idx_out[(size_t)b * Kp + c] = val[c];
        cost_out[(size_t)b * Kp + c] = key[c];
        wts_out[(size_t)b * Kp + c] = inv[c] / total;
    }
}

}  // namespace gpet

using namespace gpet;

extern "C" int gpet_score_f64(const double* Y, const float* gradT, int B, int n, int S, int M, int N, int x_st, double* cost,
                              void* stream) {
    GPET_REQUIRE(Y && gradT && cost && B > 0 && S > 0 && M >= 2, "gpet_score_f64: bad argument");
    GPET_REQUIRE(x_st >= 0 && x_st + n <= N, "gpet_score_f64: edge span outside the image");
    GPET_SUPPORTED(n >= 4 && (n % 2) == 0,
                   "gpet_score_f64: edge_length=%d must be even (scipy's Simpson end correction for an even sample "
                   "count is version dependent)", n);
    GPET_SUPPORTED(B <= 65535, "gpet_score_f64: B too large for one launch");
    cudaStream_t st = (cudaStream_t)stream;
    const int th = g_tune[GPET_TUNE_SCORE_THREADS];
    const bool pipe = g_tune[GPET_TUNE_SCORE_PIPELINE] != 0;
    if (th == 128) { if (pipe) launch_score<128, true>(Y, gradT, B, n, S, M, N, x_st, cost, st); else launch_score<128, false>(Y, gradT, B, n, S, M, N, x_st, cost, st); }
    else if (th == 512) { if (pipe) launch_score<512, true>(Y, gradT, B, n, S, M, N, x_st, cost, st); else launch_score<512, false>(Y, gradT, B, n, S, M, N, x_st, cost, st); }
    else { if (pipe) launch_score<256, true>(Y, gradT, B, n, S, M, N, x_st, cost, st); else launch_score<256, false>(Y, gradT, B, n, S, M, N, x_st, cost, st); }
    return check_launch("score_kernel");
}

extern "C" int gpet_topk_f64(const double* cost, int B, int S, int Kp, int32_t* idx, double* best_cost, double* wts,
                             void* stream) {
    GPET_REQUIRE(cost && idx && best_cost && wts && B > 0 && S > 0 && Kp > 0 && Kp <= S, "gpet_topk_f64: bad argument");
    GPET_SUPPORTED(S <= 8192, "gpet_topk_f64: S=%d > 8192 not supported by the shared-memory sort", S);
    int P2 = 1;
    while (P2 < S) P2 <<= 1;
    const size_t smem = ((size_t)P2 + (P2 + 1) / 2 + Kp + 2) * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("topk smem attribute: %s", cudaGetErrorString(e));
        return GPET_ERR_CUDA;
    }
    topk_kernel<<<B, TK_THREADS, smem, (cudaStream_t)stream>>>(cost, S, P2, Kp, idx, best_cost, wts);
    return check_launch("topk_kernel");
}
