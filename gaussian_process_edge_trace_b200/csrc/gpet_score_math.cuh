// Arithmetic of the curve cost (cost_funct, gpet.py:371-410; scipy _basic_simpson non-uniform branch), shared by the
// scoring kernels (gpet_score.cu) and the fused sampling + scoring kernel (gpet_sample_score.cu).
#pragma once
#include "gpet_common.cuh"

namespace gpet {

// ---- arithmetic building blocks -----------------------------------------------------------------------------------
// The kernel is bound by the FP64 pipe (64 lanes/SM/clk on sm_100: one warp-wide D-instruction every 2 clk per SM
// quarter), not by HBM, unless the per-point operation count is cut hard.  Every helper below is therefore the
// shortest sequence that is still accurate to ~1 ulp; bit-exactness with numpy is not attainable anyway because
// numpy's pairwise summation order differs (costs agree with the reference to ~1e-14 relative).

// sqrt(q) for finite q >= 1: MUFU.RSQ64H seed (~2^-21), then one cubically convergent Goldschmidt step (5 D-ops,
// no slow-path branch):  s0 = q r0, e = 1 - s0 r0, sqrt(q) = s0 / sqrt(1 - e) = s0 + s0 e (1/2 + 3/8 e) + O(e^3).
__device__ __forceinline__ double sqrt_ge1(double q) {
    double r0;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(q));
    const double s0 = q * r0;
    const double e = fma(-s0, r0, 1.0);
    const double p = fma(0.375, e, 0.5);
    return fma(s0 * e, p, s0);
}

// 1/d for finite normal d > 0: MUFU.RCP64H seed, one cubic Newton step (3 D-ops, no slow-path branch).
__device__ __forceinline__ double rcp_pos(double d) {
    double x0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x0) : "d"(d));
    const double e = fma(-d, x0, 1.0);
    const double e2 = fma(e, e, e);
    return fma(x0, e2, x0);
}

// float (>= 0, finite) -> double with two ALU instructions instead of the conversion pipe.  Exact for normal
// floats; zero and subnormals map to <= 2^-126, which vanishes against the +1e-3 offset of the integrand.
__device__ __forceinline__ double f32_to_f64_nonneg(float v) {
    const unsigned int u = __float_as_uint(v);
    return __hiloint2double((int)((u >> 3) + (896u << 20)), (int)(u << 29));
}

// The two bilinear taps of one curve point, fetched early and combined later (FITPACK bispeu with kx=ky=1 on
// integer knots == clamped 2-tap lerp):  yc = clip(y, 0, M-1), i0 = min(floor(yc), M-2), f = yc - i0.
// floor and its way back run on the conversion pipe (F2I/I2F), which is otherwise idle.  `col` points at row 0 of a
// guarded column (gpet_transpose_f32: entries -1 and M repeat rows 0 and M-1), so only the integer row is clamped,
// to [-1, M-1]: a clamped point reads the same value twice and its (unclamped, finite) weight multiplies zero.
struct Taps {
    float g0, g1;
    double f;
};

__device__ __forceinline__ Taps fetch_taps(const float* __restrict__ col, double y, int Mm1) {
    const int i0 = __double2int_rd(y);                 // NaN -> 0 (f carries the NaN), +-inf saturate
    Taps t;
    t.f = y - __int2double_rn(i0);
    const float* p = col + min(max(i0, -1), Mm1);
    t.g0 = __ldg(p);
    t.g1 = __ldg(p + 1);
    return t;
}

// same, addressed as base[off + row] with a 32-bit element offset (one integer add + one wide multiply-add)
__device__ __forceinline__ Taps fetch_taps_off(const float* __restrict__ base, int off, double y, int Mm1) {
    const int i0 = __double2int_rd(y);
    Taps t;
    t.f = y - __int2double_rn(i0);
    const float* p = base + (off + min(max(i0, -1), Mm1));
    t.g0 = __ldg(p);
    t.g1 = __ldg(p + 1);
    return t;
}

// g0 + f (g1 - g0); the +1e-3 of the reference integrand is added once per curve (Simpson is exact on constants)
__device__ __forceinline__ double finish_taps(const Taps& t) {
    const double g0 = f32_to_f64_nonneg(t.g0), g1 = f32_to_f64_nonneg(t.g1);
    return fma(t.f, g1 - g0, g0);
}

// 6 x one composite Simpson pair on a non-uniform abscissa (scipy _basic_simpson):
//   hs/6 (y0 (2 - h1/h0) + y1 hs^2/(h0 h1) + y2 (2 - h0/h1)) = hs/(6 h0 h1) (y0 h1 (2h0-h1) + y1 hs^2 + y2 h0 (2h1-h0))
__device__ __forceinline__ double simpson6_term(double y0, double y1, double y2, double h0, double h1) {
    const double hs = h0 + h1, hp = h0 * h1, hp2 = hp + hp;
    // h1 (2h0 - h1) = 2 h0 h1 - h1^2,   h0 (2h1 - h0) = 2 h0 h1 - h0^2
    const double num = fma(y2, fma(-h0, h0, hp2), fma(y1, hs * hs, y0 * fma(-h1, h1, hp2)));
    return (hs * rcp_pos(hp)) * num;
}

// Running state of one curve between Simpson pairs.
struct CurveState {
    double y1;             // curve value at sample 2p+1 (first interior sample of the next pair)
    double g0;             // integrand (without the 1e-3 offset) at sample 2p
    double t0;             // cumsum abscissa at sample 2p (SCAN) / running span (no SCAN)
    double seg_last;       // segment length of the last sample processed
    double AL4, LI6;       // sum of the odd segments, 6 x line integral
};

// Arithmetic of one Simpson pair: samples 2p, 2p+1, 2p+2 with curve values y1 = y[2p+1] (state), y2 = y[2p+2],
// y3 = y[2p+3] and the gradient taps of samples 2p+1, 2p+2.
template <bool SCAN>
__device__ __forceinline__ void simpson_pair_math(CurveState& c, const double y2, const double y3, const Taps& ta,
                                                  const Taps& tb) {
    double d = y2 - c.y1;
    const double seg1 = sqrt_ge1(fma(d, d, 1.0));
    d = y3 - y2;
    const double seg2 = sqrt_ge1(fma(d, d, 1.0));
    double h0 = seg1, h1 = seg2;
    if (SCAN) {
        const double t1 = c.t0 + seg1;
        const double t2 = t1 + seg2;
        h0 = t1 - c.t0;
        h1 = t2 - t1;
        c.t0 = t2;
    } else {
        c.t0 += seg1 + seg2;
    }
    const double g1 = finish_taps(ta);
    const double g2 = finish_taps(tb);
    c.LI6 += simpson6_term(c.g0, g1, g2, h0, h1);
    c.AL4 += seg1;     // odd samples; the even ones follow from the total length at the end
    c.y1 = y3;
    c.g0 = g2;
    c.seg_last = seg2;
}

template <bool SCAN>
__device__ __forceinline__ void curve_begin(CurveState& c, const double y0, const double y1, const float* col,
                                            const int Mm1, double& tfirst) {
    c.y1 = y1;
    const double d = y1 - y0;
    tfirst = sqrt_ge1(fma(d, d, 1.0));
    c.t0 = SCAN ? tfirst : 0.0;
    c.seg_last = tfirst;
    c.g0 = finish_taps(fetch_taps(col, y0, Mm1));
    c.AL4 = c.LI6 = 0.0;
}

// 3 AL = seg_first + 4 sum(odd) + 2 sum(even interior) + seg_last = 2 sum(odd) + 2 T - seg_first - seg_last with
// T = sum of all segments;   LI = LI6/6 + 1e-3 (t_last - t_first)
template <bool SCAN>
__device__ __forceinline__ double curve_cost(const CurveState& c, const double tfirst) {
    const double span = SCAN ? c.t0 - tfirst : c.t0;
    const double T = SCAN ? c.t0 : c.t0 + tfirst;
    const double AL = (2.0 * (c.AL4 + T) - (tfirst + c.seg_last)) * (1.0 / 3.0);
    const double LI = fma(c.LI6, 1.0 / 6.0, 1e-3 * span);
    return AL / LI;
}


}  // namespace gpet
