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
#include "gpet_score_math.cuh"
#include <type_traits>

namespace gpet {


// ---- general variant: one thread per curve, curve values prefetched in registers ------------------------------------
// Used when the bulk-copy path below is not applicable (odd S / unaligned Y).  SCAN: the Simpson abscissa is the
// running sum of the segment lengths as in the reference (h = t[j+1] - t[j]); without it h = seg (differs by
// ulp(t) ~ 1e-13 relative).
template <bool SCAN, bool PREFETCH>
__device__ __forceinline__ void simpson_pair(CurveState& c, const double y2, const double y3, double& next2,
                                             double& next3, const double*& ynext, const size_t Sz,
                                             const float*& col, const int Mm1) {
    if (PREFETCH) {
        next2 = __ldg(ynext);
        next3 = __ldg(ynext + Sz);
        ynext += 2 * Sz;
    }
    const int Mp = Mm1 + 3;
    const Taps ta = fetch_taps(col, c.y1, Mm1);
    const Taps tb = fetch_taps(col + Mp, y2, Mm1);
    col += 2 * Mp;
    simpson_pair_math<SCAN>(c, y2, y3, ta, tb);
}

template <int THREADS, bool SCAN>
__global__ void __launch_bounds__(THREADS)
score_kernel(const double* __restrict__ Y, const float* __restrict__ gradT, const int32_t* __restrict__ img_index,
             int n, int S, int M, int N, int x_st, double* __restrict__ cost) {
    const int b = blockIdx.y;
    const int s = blockIdx.x * THREADS + threadIdx.x;
    if (s >= S) return;
    const size_t Sz = (size_t)S;
    const int Mp = M + 2, Mm1 = M - 1;
    const double* yp = Y + (size_t)b * n * Sz + s;
    const int img = img_index ? img_index[b] : b;
    const float* col = gradT + ((size_t)img * N + x_st) * Mp + 1;
    // K = n - 1 Simpson samples (j = 0 .. n-2), K odd <=> n even; pairs p = 0 .. (K-1)/2 - 1
    const int P = (n - 2) / 2;
    CurveState c;
    double tfirst;
    curve_begin<SCAN>(c, __ldg(yp), __ldg(yp + Sz), col, Mm1, tfirst);
    col += Mp;
    double a2 = __ldg(yp + 2 * Sz), a3 = __ldg(yp + 3 * Sz);   // pair 0
    double b2 = 0.0, b3 = 0.0;
    const double* ynext = yp + 4 * Sz;
    int p = 0;
#pragma unroll 1
    for (; p + 2 < P; p += 2) {   // two pairs per trip: the prefetch registers ping-pong without moves
        simpson_pair<SCAN, true>(c, a2, a3, b2, b3, ynext, Sz, col, Mm1);
        simpson_pair<SCAN, true>(c, b2, b3, a2, a3, ynext, Sz, col, Mm1);
    }
    if (p + 1 < P) {
        simpson_pair<SCAN, true>(c, a2, a3, b2, b3, ynext, Sz, col, Mm1);
        simpson_pair<SCAN, false>(c, b2, b3, a2, a3, ynext, Sz, col, Mm1);
    } else {
        simpson_pair<SCAN, false>(c, a2, a3, b2, b3, ynext, Sz, col, Mm1);
    }
    cost[(size_t)b * S + s] = curve_cost<SCAN>(c, tfirst);
}

// ---- odd edge_length: an EVEN number K = n - 1 of Simpson samples --------------------------------------------------
// scipy.integrate.simpson (>= 1.11, the version the oracle pins; the reference's `simps` alias) then applies composite
// Simpson to the first K - 1 samples and Cartwright's correction to the last interval:
//   + alpha f[K-1] + beta f[K-2] - eta f[K-3],  alpha = (2 h1^2 + 3 h0 h1) / (6 (h0 + h1)),  beta = (h1^2 + 3 h0 h1) / (6 h0),
//   eta = h1^3 / (6 h0 (h0 + h1)),  h0 = x[K-2] - x[K-3], h1 = x[K-1] - x[K-2]      (uniform grid: 5/12, 2/3, 1/12).
// Plain one-thread-per-curve form (all BASELINE configurations have an even edge_length and take the kernels above).
__global__ void __launch_bounds__(128)
score_odd_kernel(const double* __restrict__ Y, const float* __restrict__ gradT, const int32_t* __restrict__ img_index,
                 int n, int S, int M, int N, int x_st, double* __restrict__ cost) {
    const int b = blockIdx.y;
    const int s = blockIdx.x * 128 + threadIdx.x;
    if (s >= S) return;
    const size_t Sz = (size_t)S;
    const int Mp = M + 2, Mm1 = M - 1;
    const double* yp = Y + (size_t)b * n * Sz + s;
    const int img = img_index ? img_index[b] : b;
    const float* col = gradT + ((size_t)img * N + x_st) * Mp + 1;
    const int P = (n - 3) / 2;                   // Simpson pairs over samples 0 .. 2P = K - 2
    double ya = __ldg(yp), yb = __ldg(yp + Sz);
    double d = yb - ya;
    double s0 = sqrt(fma(d, d, 1.0));            // seg[0]
    double t0 = s0;                              // t[0] = cumsum(seg)[0]
    const double tfirst = t0;
    double g0 = finish_taps(fetch_taps(col, ya, Mm1));
    double gm = 0.0, tm = 0.0, sm_ = 0.0;        // sample 2p - 1 of the latest pair
    double LI6 = 0.0, AL3 = 0.0;
    ya = yb;                                     // y[1]
    for (int p = 0; p < P; ++p) {
        const double y2 = __ldg(yp + (size_t)(2 * p + 2) * Sz), y3 = __ldg(yp + (size_t)(2 * p + 3) * Sz);
        d = y2 - ya;
        const double seg1 = sqrt(fma(d, d, 1.0));
        d = y3 - y2;
        const double seg2 = sqrt(fma(d, d, 1.0));
        const double t1 = t0 + seg1, t2 = t1 + seg2;
        const double g1 = finish_taps(fetch_taps(col + (size_t)(2 * p + 1) * Mp, ya, Mm1));
        const double g2 = finish_taps(fetch_taps(col + (size_t)(2 * p + 2) * Mp, y2, Mm1));
        LI6 += simpson6_term(g0, g1, g2, t1 - t0, t2 - t1);
        AL3 += (s0 + 4.0 * seg1) + seg2;
        gm = g1; tm = t1; sm_ = seg1;
        g0 = g2; t0 = t2; s0 = seg2;
        ya = y3;
    }
    // last interval: samples K-3 = 2P-1 (gm, tm, sm_), K-2 = 2P (g0, t0, s0), K-1 = 2P+1
    const double yl = __ldg(yp + (size_t)(2 * P + 2) * Sz);
    d = yl - ya;
    const double sl = sqrt(fma(d, d, 1.0));
    const double tl = t0 + sl;
    const double gl = finish_taps(fetch_taps(col + (size_t)(2 * P + 1) * Mp, ya, Mm1));
    const double h0 = t0 - tm, h1 = tl - t0;
    const double alpha = (2.0 * h1 * h1 + 3.0 * h0 * h1) / (6.0 * (h1 + h0));
    const double beta = (h1 * h1 + 3.0 * h0 * h1) / (6.0 * h0);
    const double eta = (h1 * h1 * h1) / (6.0 * h0 * (h0 + h1));
    // the +1e-3 of the integrand: Simpson and the correction are exact on constants
    const double LI = LI6 * (1.0 / 6.0) + (alpha * gl + beta * g0 - eta * gm) + 1e-3 * (tl - tfirst);
    const double AL = AL3 * (1.0 / 3.0) + ((5.0 / 12.0) * sl + (2.0 / 3.0) * s0 - (1.0 / 12.0) * sm_);
    cost[(size_t)b * S + s] = AL / LI;
}

template <int THREADS, bool SCAN>
static void launch_score(const double* Y, const float* gradT, const int32_t* ii, int B, int n, int S, int M, int N, int x_st, double* cost,
                         cudaStream_t st) {
    dim3 grid((S + THREADS - 1) / THREADS, B);
    score_kernel<THREADS, SCAN><<<grid, THREADS, 0, st>>>(Y, gradT, ii, n, S, M, N, x_st, cost);
}

// ---- streamed variant: curve values staged through shared memory by the bulk-copy (TMA) engine ----------------------
// The register-prefetching kernel above is latency bound: with 40 registers/thread an SM holds 48 warps x 512 B of
// curve values in flight, less than the ~45 KB/SM that HBM latency x bandwidth requires.  Here one elected thread
// streams [4 rows] x [128 curves] tiles of Y (each row segment = 1 KB contiguous) into a STAGES-deep ring with
// cp.async.bulk + mbarrier transaction counts, so STAGES-1 tiles per CTA are always in flight at no register cost,
// and because the curve values of the NEXT Simpson pair are already in shared memory, the dependent gradient gathers
// are issued one pair ahead of their use.  The tile loop is unrolled STAGES times so every ring slot is a constant.
constexpr int SC_T = 128;     // curves per CTA
constexpr int SC_ROWS = 4;    // rows per tile = two Simpson pairs

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// Warp-specialised: warps 0..3 hold one curve per thread, warp 4 is the producer.  Per tile the producer waits for
// the ring slot to be released (empty barrier, one arrival per consumer warp), arms the full barrier with the byte
// count and issues the row copies; with its otherwise idle lanes it also pulls the four gradient columns of that tile
// into L2, STAGES tiles (2 STAGES Simpson pairs) before the consumers gather from them.  Consumers never meet a
// CTA-wide barrier: they copy their four values of a tile to registers, release the slot and go on.
constexpr int SC_THREADS_STREAM = SC_T + 32;

template <bool SCAN, int STAGES, int MINB>
__global__ void __launch_bounds__(SC_THREADS_STREAM, MINB)
score_stream_kernel(const double* __restrict__ Y, const float* __restrict__ gradT,
                    const int32_t* __restrict__ img_index, int n, int S, int M, int N, int x_st,
                    double* __restrict__ cost) {
    __shared__ __align__(128) double ring[STAGES][SC_ROWS][SC_T];
    __shared__ __align__(8) unsigned long long full[STAGES], empty[STAGES];
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    const int s0 = blockIdx.x * SC_T;
    const int cnt = min(SC_T, S - s0);                  // curves of this CTA (even: S is even)
    const int nchunks = (n + SC_ROWS - 1) / SC_ROWS;    // n even => the last tile has 4 or 2 rows
    const uint32_t ring0 = smem_u32(&ring[0][0][0]), full0 = smem_u32(&full[0]), empty0 = smem_u32(&empty[0]);
    constexpr uint32_t TILE_BYTES = SC_ROWS * SC_T * 8, ROW_BYTES = SC_T * 8;
    const int Mp = M + 2, Mm1 = M - 1;
    const int img = img_index ? img_index[b] : b;
    const float* gt = gradT + ((size_t)img * N + x_st) * Mp + 1;    // row 0 of the column of sample 0
    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(full0 + 8 * i, 1);
            mbar_init(empty0 + 8 * i, SC_T / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= SC_T) {
        // ------------------------------------------------ producer warp ----------------------------------------------
        const uint32_t rowbytes = (uint32_t)cnt * 8u;
        const double* src = Y + (size_t)b * n * S + s0;
        const char* gcol = reinterpret_cast<const char*>(gt - 1);           // guarded start of the column of sample 0
        const size_t tile_cols_bytes = (size_t)SC_ROWS * Mp * sizeof(float);
        int slot = 0;
        uint32_t phase = 1;          // first pass over the ring: the slots are free (wait on the preceding phase)
#pragma unroll 1
        for (int q = 0; q < nchunks; ++q) {
            const int rows = min(SC_ROWS, n - q * SC_ROWS);
            mbar_wait(empty0 + 8 * slot, phase);
            if (lane == 0) {
                const uint32_t bar = full0 + 8 * slot, dst = ring0 + TILE_BYTES * slot;
                mbar_expect_tx(bar, rows * rowbytes);
                for (int r = 0; r < rows; ++r) bulk_g2s(dst + ROW_BYTES * r, src + (size_t)r * S, rowbytes, bar);
            }
            src += (size_t)SC_ROWS * S;
            for (size_t o = (size_t)lane * 128; o < tile_cols_bytes; o += 32 * 128) prefetch_l2(gcol + o);
            gcol += tile_cols_bytes;
            if (++slot == STAGES) { slot = 0; phase ^= 1u; }
        }
        return;
    }

    // ---------------------------------------------------- consumer warps ---------------------------------------------
    static_assert((STAGES & (STAGES - 1)) == 0, "STAGES must be a power of two");
    const double* mine = &ring[0][0][tid];
    CurveState c;
    double tfirst;
    Taps ta, tb, ua, ub;   // taps ping-pong: the pair on rows 0,1 of a tile consumes t*, the pair on rows 2,3 consumes u*
    double r0, r1, r2, r3; // the curve values of the current tile, in registers: its ring slot is released right away
    mbar_wait(full0, 0);
    r0 = mine[0]; r1 = mine[SC_T]; r2 = mine[2 * SC_T]; r3 = mine[3 * SC_T];
    __syncwarp();
    if (lane == 0) mbar_arrive(empty0);
    curve_begin<SCAN>(c, r0, r1, gt, Mm1, tfirst);
    int off = Mp;                                  // element offset of the column of sample 2p+1 (p = current pair)
    ua = fetch_taps_off(gt, off, r1, Mm1);         // pair 0 lives on rows 2,3 of tile 0
    ub = fetch_taps_off(gt, off + Mp, r2, Mm1);
    int slot = 0;
    uint32_t phase = 0;
    // ---- tile q -> q+1: the pair on rows 2,3 of tile q, then the pair on rows 0,1 of tile q+1 --------------------------
    // FOUR: tile q+1 has rows 2,3 (every tile except possibly the last one)
    auto advance = [&](auto four_tag) {
        constexpr bool FOUR = decltype(four_tag)::value;
        slot = (slot + 1) & (STAGES - 1);
        phase ^= (slot == 0) ? 1u : 0u;
        const double* nx = mine + slot * (SC_ROWS * SC_T);
        mbar_wait(full0 + 8 * slot, phase);
        const double n0 = nx[0], n1 = nx[SC_T];
        double n2 = 0.0, n3 = 0.0;
        if (FOUR) { n2 = nx[2 * SC_T]; n3 = nx[3 * SC_T]; }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * slot);
        // pair 2q (rows 2,3 of tile q); request the taps of pair 2q+1 = samples 4q+3 (y = r3), 4q+4 (y = n0)
        ta = fetch_taps_off(gt, off + 2 * Mp, r3, Mm1);
        tb = fetch_taps_off(gt, off + 3 * Mp, n0, Mm1);
        simpson_pair_math<SCAN>(c, r2, r3, ua, ub);
        // pair 2q+1 (rows 0,1 of tile q+1); request the taps of pair 2q+2 = samples 4q+5 (y = n1), 4q+6 (y = n2)
        if (FOUR) {
            ua = fetch_taps_off(gt, off + 4 * Mp, n1, Mm1);
            ub = fetch_taps_off(gt, off + 5 * Mp, n2, Mm1);
        }
        off += 4 * Mp;
        simpson_pair_math<SCAN>(c, n0, n1, ta, tb);
        r2 = n2; r3 = n3;
    };
#pragma unroll 1
    for (int q = 0; q + 2 < nchunks; ++q) advance(std::true_type{});
    if (nchunks > 1) {
        if ((n & 3) == 0) advance(std::true_type{});
        else advance(std::false_type{});
    }
    // ---- last tile: its pair on rows 2,3, if present -----------------------------------------------------------------
    if ((n & 3) == 0) simpson_pair_math<SCAN>(c, r2, r3, ua, ub);
    if (tid < cnt) cost[(size_t)b * S + s0 + tid] = curve_cost<SCAN>(c, tfirst);
}

template <bool SCAN, int STAGES, int MINB>
static void launch_score_stream(const double* Y, const float* gradT, const int32_t* ii, int B, int n, int S, int M, int N, int x_st,
                                double* cost, cudaStream_t st) {
    dim3 grid((S + SC_T - 1) / SC_T, B);
    score_stream_kernel<SCAN, STAGES, MINB><<<grid, SC_THREADS_STREAM, 0, st>>>(Y, gradT, ii, n, S, M, N, x_st, cost);
}

// ---- streamed variant with CPT curves per consumer thread --------------------------------------------------------------
// Same ring / producer-warp structure as score_stream_kernel, but every consumer thread carries CPT independent curves
// (tid, tid + 128, ...): the dependent FP64 chain of one Simpson pair (~25 D-ops deep for 39 D-ops of work) is
// interleaved with the chains of the thread's other curves, and the per-tile overhead (barrier wait, slot/phase and
// offset arithmetic, loop control) is paid once per CPT curves.  Tile = [4 rows] x [128 CPT curves].
template <bool SCAN, int STAGES, int MINB, int CPT>
__global__ void __launch_bounds__(SC_THREADS_STREAM, MINB)
score_streamN_kernel(const double* __restrict__ Y, const float* __restrict__ gradT,
                     const int32_t* __restrict__ img_index, int n, int S, int M, int N, int x_st,
                     double* __restrict__ cost) {
    constexpr int TW = SC_T * CPT;                       // curves per CTA
    extern __shared__ __align__(128) unsigned char sc_smem[];
    double* ring = reinterpret_cast<double*>(sc_smem);                                          // [STAGES][SC_ROWS][TW]
    unsigned long long* full = reinterpret_cast<unsigned long long*>(ring + STAGES * SC_ROWS * TW);
    unsigned long long* empty = full + STAGES;
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    const int s0 = blockIdx.x * TW;
    const int cnt = min(TW, S - s0);
    const int nchunks = (n + SC_ROWS - 1) / SC_ROWS;
    const uint32_t ring0 = smem_u32(ring), full0 = smem_u32(full), empty0 = smem_u32(empty);
    constexpr uint32_t TILE_BYTES = SC_ROWS * TW * 8, ROW_BYTES = TW * 8;
    const int Mp = M + 2, Mm1 = M - 1;
    const int img = img_index ? img_index[b] : b;
    const float* gt = gradT + ((size_t)img * N + x_st) * Mp + 1;
    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(full0 + 8 * i, 1);
            mbar_init(empty0 + 8 * i, SC_T / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= SC_T) {
        const uint32_t rowbytes = (uint32_t)cnt * 8u;
        const double* src = Y + (size_t)b * n * S + s0;
        const char* gcol = reinterpret_cast<const char*>(gt - 1);
        const size_t tile_cols_bytes = (size_t)SC_ROWS * Mp * sizeof(float);
        int slot = 0;
        uint32_t phase = 1;
#pragma unroll 1
        for (int q = 0; q < nchunks; ++q) {
            const int rows = min(SC_ROWS, n - q * SC_ROWS);
            mbar_wait(empty0 + 8 * slot, phase);
            if (lane == 0) {
                const uint32_t bar = full0 + 8 * slot, dst = ring0 + TILE_BYTES * slot;
                mbar_expect_tx(bar, rows * rowbytes);
                for (int r = 0; r < rows; ++r) bulk_g2s(dst + ROW_BYTES * r, src + (size_t)r * S, rowbytes, bar);
            }
            src += (size_t)SC_ROWS * S;
            // whole gradient columns into L2 (limiting the prefetch to the band of rows the CTA's curves occupy was
            // measured slower: at S = 1000 the curves of a column still span half the image, and a miss costs more
            // than the bytes saved)
            for (size_t o = (size_t)lane * 128; o < tile_cols_bytes; o += 32 * 128) prefetch_l2(gcol + o);
            gcol += tile_cols_bytes;
            if (++slot == STAGES) { slot = 0; phase ^= 1u; }
        }
        return;
    }

    static_assert((STAGES & (STAGES - 1)) == 0, "STAGES must be a power of two");
    const double* mine = ring + tid;
    CurveState c[CPT];
    double tfirst[CPT];
    Taps ta[CPT], tb[CPT], ua[CPT], ub[CPT];
    double r2[CPT], r3[CPT];
    mbar_wait(full0, 0);
    {
        double r0[CPT], r1[CPT];
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            r0[k] = mine[k * SC_T]; r1[k] = mine[TW + k * SC_T]; r2[k] = mine[2 * TW + k * SC_T]; r3[k] = mine[3 * TW + k * SC_T];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0);
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            curve_begin<SCAN>(c[k], r0[k], r1[k], gt, Mm1, tfirst[k]);
            ua[k] = fetch_taps_off(gt, Mp, r1[k], Mm1);
            ub[k] = fetch_taps_off(gt, 2 * Mp, r2[k], Mm1);
        }
    }
    int off = Mp;
    int slot = 0;
    uint32_t phase = 0;
    auto advance = [&](auto four_tag) {
        constexpr bool FOUR = decltype(four_tag)::value;
        slot = (slot + 1) & (STAGES - 1);
        phase ^= (slot == 0) ? 1u : 0u;
        const double* nx = mine + slot * (SC_ROWS * TW);
        mbar_wait(full0 + 8 * slot, phase);
        double n0[CPT], n1[CPT], n2[CPT], n3[CPT];
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            n0[k] = nx[k * SC_T]; n1[k] = nx[TW + k * SC_T];
            n2[k] = 0.0; n3[k] = 0.0;
            if (FOUR) { n2[k] = nx[2 * TW + k * SC_T]; n3[k] = nx[3 * TW + k * SC_T]; }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * slot);
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            ta[k] = fetch_taps_off(gt, off + 2 * Mp, r3[k], Mm1);
            tb[k] = fetch_taps_off(gt, off + 3 * Mp, n0[k], Mm1);
        }
#pragma unroll
        for (int k = 0; k < CPT; ++k) simpson_pair_math<SCAN>(c[k], r2[k], r3[k], ua[k], ub[k]);
        if (FOUR) {
#pragma unroll
            for (int k = 0; k < CPT; ++k) {
                ua[k] = fetch_taps_off(gt, off + 4 * Mp, n1[k], Mm1);
                ub[k] = fetch_taps_off(gt, off + 5 * Mp, n2[k], Mm1);
            }
        }
        off += 4 * Mp;
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            simpson_pair_math<SCAN>(c[k], n0[k], n1[k], ta[k], tb[k]);
            r2[k] = n2[k]; r3[k] = n3[k];
        }
    };
#pragma unroll 1
    for (int q = 0; q + 2 < nchunks; ++q) advance(std::true_type{});
    if (nchunks > 1) {
        if ((n & 3) == 0) advance(std::true_type{});
        else advance(std::false_type{});
    }
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        if ((n & 3) == 0) simpson_pair_math<SCAN>(c[k], r2[k], r3[k], ua[k], ub[k]);
        if (tid + k * SC_T < cnt) cost[(size_t)b * S + s0 + tid + k * SC_T] = curve_cost<SCAN>(c[k], tfirst[k]);
    }
}

template <bool SCAN, int STAGES, int MINB, int CPT>
static void launch_score_streamN(const double* Y, const float* gradT, const int32_t* ii, int B, int n, int S, int M, int N,
                                 int x_st, double* cost, cudaStream_t st) {
    constexpr int TW = SC_T * CPT;
    constexpr size_t smem = (size_t)STAGES * SC_ROWS * TW * 8 + 2 * STAGES * 8;
    // function attributes are per device (and cheap to set): no per-process "done once" flag
    cudaFuncSetAttribute(score_streamN_kernel<SCAN, STAGES, MINB, CPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid((S + TW - 1) / TW, B);
    score_streamN_kernel<SCAN, STAGES, MINB, CPT><<<grid, SC_THREADS_STREAM, smem, st>>>(Y, gradT, ii, n, S, M, N, x_st, cost);
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

// ---- top-N_keep for large sample counts (S > 8192, e.g. BASELINE config 2: S = 100 000, N_keep = 10 000) ------------
// One CTA per trace: (1) 8-pass radix select of the N_keep-th smallest key over the costs in global memory,
// (2) compaction of the N_keep selected (cost, index) pairs into shared memory, (3) the same bitonic sort as above on
// just those, (4) weights.  Keys are the order-preserving uint64 image of the doubles; NaN sorts last.
constexpr int TKL_THREADS = 1024;

__device__ __forceinline__ unsigned long long order_key(double c) {
    if (c != c) c = __longlong_as_double(0x7ff0000000000000LL);
    const unsigned long long u = (unsigned long long)__double_as_longlong(c);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ULL);
}

__global__ void __launch_bounds__(TKL_THREADS)
topk_large_kernel(const double* __restrict__ cost, int S, int P2, int Kp, int32_t* __restrict__ idx_out,
                  double* __restrict__ cost_out, double* __restrict__ wts_out) {
    extern __shared__ double sm[];
    double* key = sm;                    // P2
    int* val = (int*)(sm + P2);          // P2
    __shared__ unsigned int hist[256];
    __shared__ unsigned long long prefix_s;
    __shared__ int need_s, n_sel, n_eq;
    __shared__ double total;
    const int b = blockIdx.x, tid = threadIdx.x;
    const double* cb = cost + (size_t)b * S;
    if (tid == 0) { prefix_s = 0ULL; need_s = Kp; }
    __syncthreads();
    for (int pass = 7; pass >= 0; --pass) {
        for (int i = tid; i < 256; i += TKL_THREADS) hist[i] = 0u;
        __syncthreads();
        const unsigned long long prefix = prefix_s;
        const int sh = 8 * pass;
        for (int i = tid; i < S; i += TKL_THREADS) {
            const unsigned long long u = order_key(cb[i]);
            if (pass == 7 || (u >> (sh + 8)) == prefix) atomicAdd(&hist[(unsigned)(u >> sh) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            int need = need_s, bin = 0;
            unsigned int cum = 0;
            for (; bin < 256; ++bin) {
                if (cum + hist[bin] >= (unsigned)need) break;
                cum += hist[bin];
            }
            need_s = need - (int)cum;
            prefix_s = (prefix << 8) | (unsigned long long)bin;
        }
        __syncthreads();
    }
    const unsigned long long T = prefix_s;     // key of the Kp-th smallest cost; need_s of the elements equal to it are kept
    const int need_eq = need_s;
    if (tid == 0) { n_sel = 0; n_eq = 0; }
    for (int i = tid; i < P2; i += TKL_THREADS) {
        key[i] = __longlong_as_double(0x7ff0000000000000LL);
        val[i] = 0x7fffffff;
    }
    __syncthreads();
    for (int i = tid; i < S; i += TKL_THREADS) {
        const double c = cb[i];
        const unsigned long long u = order_key(c);
        if (u < T) {
            const int slot = atomicAdd(&n_sel, 1);
            key[slot] = (c != c) ? __longlong_as_double(0x7ff0000000000000LL) : c;
            val[slot] = i;
        } else if (u == T) {
            atomicAdd(&n_eq, 1);
        }
    }
    __syncthreads();
    if (tid == 0) {     // ties at the threshold: lowest indices first (exactly equal costs are practically impossible)
        int slot = n_sel, taken = 0;
        for (int i = 0; i < S && taken < need_eq; ++i) {
            const double c = cb[i];
            if (order_key(c) == T) {
                key[slot] = (c != c) ? __longlong_as_double(0x7ff0000000000000LL) : c;
                val[slot] = i;
                ++slot;
                ++taken;
            }
        }
    }
    __syncthreads();
    for (int k = 2; k <= P2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < P2; i += TKL_THREADS) {
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
    // weights (gpet.py:492-493): inv = 1/cost; w = inv / np.sum(inv)   (numpy pairwise sum, one thread; inv staged in wts)
    double* wb = wts_out + (size_t)b * Kp;
    for (int c = tid; c < Kp; c += TKL_THREADS) wb[c] = 1.0 / key[c];
    __syncthreads();
    if (tid == 0) total = np_pairwise_sum(wb, Kp);
    __syncthreads();
    for (int c = tid; c < Kp; c += TKL_THREADS) {
        idx_out[(size_t)b * Kp + c] = val[c];
        cost_out[(size_t)b * Kp + c] = key[c];
        wb[c] = wb[c] / total;
    }
}

}  // namespace gpet

using namespace gpet;

extern "C" int gpet_score_f64(const double* Y, const float* gradT, const int32_t* img_index, int B, int n, int S, int M,
                              int N, int x_st, double* cost, void* stream) {
    GPET_REQUIRE(Y && gradT && cost && B > 0 && S > 0 && M >= 2, "gpet_score_f64: bad argument");
    GPET_REQUIRE(x_st >= 0 && x_st + n <= N, "gpet_score_f64: edge span outside the image");
    GPET_SUPPORTED(n >= 4, "gpet_score_f64: edge_length=%d must be at least 4", n);
    GPET_SUPPORTED(B <= 65535, "gpet_score_f64: B too large for one launch");
    cudaStream_t st = (cudaStream_t)stream;
    if (n % 2) {          // even Simpson sample count: composite rule + Cartwright's last-interval correction
        dim3 grid((S + 127) / 128, B);
        score_odd_kernel<<<grid, 128, 0, st>>>(Y, gradT, img_index, n, S, M, N, x_st, cost);
        return check_launch("score_odd_kernel");
    }
    const int th = g_tune[GPET_TUNE_SCORE_THREADS];
    const bool scan = g_tune[GPET_TUNE_SCORE_SCAN] != 0;
    const int stages = g_tune[GPET_TUNE_SCORE_STAGES];
    // the bulk-copy path needs 16-byte aligned row segments: S even and Y 16-byte aligned
    if (stages > 0 && (S % 2) == 0 && ((uintptr_t)Y % 16) == 0) {
#define GPET_SC_STREAM(ST, MB) \
    do { if (scan) launch_score_stream<true, ST, MB>(Y, gradT, img_index, B, n, S, M, N, x_st, cost, st); \
         else launch_score_stream<false, ST, MB>(Y, gradT, img_index, B, n, S, M, N, x_st, cost, st); } while (0)
        const int mb = g_tune[GPET_TUNE_SCORE_MINBLOCKS];
        // curves per consumer thread: two when the launch fills the GPU several times over (>= 5 waves of 4 CTAs/SM at 148
        // SMs), one for small launches (sub-batches, converged traces dropped), where half as many CTAs leave a longer tail
        int cpt = g_tune[GPET_TUNE_SCORE_CPT];
        if (cpt == 0) cpt = ((long long)B * ((S + 2 * SC_T - 1) / (2 * SC_T)) >= 5 * 4 * 148) ? 2 : 1;
        if (cpt >= 2) {
#define GPET_SC_STREAMN(ST, MB, C) \
    do { if (scan) launch_score_streamN<true, ST, MB, C>(Y, gradT, img_index, B, n, S, M, N, x_st, cost, st); \
         else launch_score_streamN<false, ST, MB, C>(Y, gradT, img_index, B, n, S, M, N, x_st, cost, st); } while (0)
            if (cpt == 2) { if (mb == 3) GPET_SC_STREAMN(4, 3, 2); else GPET_SC_STREAMN(4, 4, 2); }
            else { GPET_SC_STREAMN(4, 3, 3); }
#undef GPET_SC_STREAMN
            return check_launch("score_streamN_kernel");
        }
        if (stages <= 4) { if (mb <= 4) GPET_SC_STREAM(4, 4); else if (mb <= 5) GPET_SC_STREAM(4, 5); else GPET_SC_STREAM(4, 6); }
        else { if (mb <= 4) GPET_SC_STREAM(8, 4); else if (mb <= 5) GPET_SC_STREAM(8, 5); else GPET_SC_STREAM(8, 6); }
#undef GPET_SC_STREAM
        return check_launch("score_stream_kernel");
    }
    if (th == 128) { if (scan) launch_score<128, true>(Y, gradT, img_index, B, n, S, M, N, x_st, cost, st); else launch_score<128, false>(Y, gradT, img_index, B, n, S, M, N, x_st, cost, st); }
    else if (th == 512) { if (scan) launch_score<512, true>(Y, gradT, img_index, B, n, S, M, N, x_st, cost, st); else launch_score<512, false>(Y, gradT, img_index, B, n, S, M, N, x_st, cost, st); }
    else { if (scan) launch_score<256, true>(Y, gradT, img_index, B, n, S, M, N, x_st, cost, st); else launch_score<256, false>(Y, gradT, img_index, B, n, S, M, N, x_st, cost, st); }
    return check_launch("score_kernel");
}

extern "C" int gpet_topk_f64(const double* cost, int B, int S, int Kp, int32_t* idx, double* best_cost, double* wts,
                             void* stream) {
    GPET_REQUIRE(cost && idx && best_cost && wts && B > 0 && S > 0 && Kp > 0 && Kp <= S, "gpet_topk_f64: bad argument");
    if (S > 8192) {     // radix select of the N_keep smallest first, then the shared-memory sort on those only
        int P2 = 1;
        while (P2 < Kp) P2 <<= 1;
        const size_t smem_l = (size_t)P2 * (sizeof(double) + sizeof(int));
        GPET_SUPPORTED(smem_l <= 200 * 1024, "gpet_topk_f64: N_keep=%d > 16384 not supported (S=%d)", Kp, S);
        cudaError_t el = cudaFuncSetAttribute(topk_large_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l);
        if (el != cudaSuccess) {
            set_error("topk smem attribute: %s", cudaGetErrorString(el));
            return GPET_ERR_CUDA;
        }
        topk_large_kernel<<<B, TKL_THREADS, smem_l, (cudaStream_t)stream>>>(cost, S, P2, Kp, idx, best_cost, wts);
        return check_launch("topk_large_kernel");
    }
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
