// Standard-normal draws of numpy's legacy generator on the device.
//
// Reference seam: sklearn_gpr.py:460-464 -> numpy RandomState(seed).multivariate_normal -> standard_normal((S, n)):
// MT19937 seeded by init_genrand(seed), doubles from two outputs ((a >> 5) * 2^26 + (b >> 6)) / 2^53, and the polar
// (Marsaglia) method of legacy_gauss: attempts (x1, x2) = 2 u - 1 are rejected unless 0 < r2 = x1^2 + x2^2 < 1; an
// accepted attempt yields f x2 then f x1 with f = sqrt(-2 log(r2) / r2).  The stream is sequential by definition;
// here it is produced in three data-parallel stages:
//   1. mt19937_stream_kernel: ONE CTA advances the 624-word state block by block (each block = three dependent
//      phases of <= 227 independent words) and writes the tempered outputs;
//   2. polar_count_kernel + scan: every attempt (4 outputs) is tested in parallel, accepted attempts are counted per
//      CTA and the counts scanned, which gives every accepted attempt its position in the output sequence;
//   3. polar_write_kernel: accepted attempts compute their two normals and store them, already transposed and
//      restricted to the columns / sample block the caller consumes (Zt[j][s - s0], j < kcols).
// Integer part and the rejection test are exact, so the SAME attempts are accepted as on the host, and every floating
// point operation but one is correctly rounded and FMA-free on both sides.  The one is log(r2): glibc's log is not
// correctly rounded but errs by less than 0.52 ulp, so it returns the correctly rounded value whenever the true logarithm
// is not within ~0.02 ulp of the midpoint of two doubles.  log_cr() below evaluates the logarithm in double-double
// arithmetic (error ~2^-64 relative), returns its correct rounding and FLAGS the inputs whose logarithm lies within
// 0.03 ulp of a midpoint (6 % of them); the flagged attempts are listed (r2, x1, x2, element index) and the host
// recomputes just those with libm's own log (gpet_host_log_f64) and patches them in.  With the patch the normals are
// numpy's bit for bit (checked on millions of draws); without it (fixups == NULL) ~0.05 % of them differ by 1-3 ulp.
#include <math.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "gpet_common.cuh"
#include "gpet_log_table.cuh"

namespace gpet {

constexpr int MT_N = 624, MT_M = 397;
constexpr uint32_t MT_UPPER = 0x80000000u, MT_LOWER = 0x7fffffffu, MT_MATRIX = 0x9908b0dfu;
constexpr int PL_THREADS = 1024;

__device__ __forceinline__ uint32_t mt_twist(uint32_t cur, uint32_t nxt, uint32_t far) {
    const uint32_t y = (cur & MT_UPPER) | (nxt & MT_LOWER);
    return far ^ (y >> 1) ^ ((y & 1u) ? MT_MATRIX : 0u);
}

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

__global__ void __launch_bounds__(256)
mt19937_stream_kernel(uint32_t seed, long long nblocks, uint32_t* __restrict__ out) {
    __shared__ uint32_t mt[MT_N];
    const int tid = threadIdx.x;
    if (tid == 0) {     // init_genrand: a sequential recurrence of 623 steps
        uint32_t v = seed;
        mt[0] = v;
        for (int i = 1; i < MT_N; ++i) {
            v = 1812433253u * (v ^ (v >> 30)) + (uint32_t)i;
            mt[i] = v;
        }
    }
    __syncthreads();
    for (long long b = 0; b < nblocks; ++b) {
        uint32_t v = 0;
        // words 0..226: inputs are all old
        if (tid < MT_N - MT_M) v = mt_twist(mt[tid], mt[tid + 1], mt[tid + MT_M]);
        __syncthreads();
        if (tid < MT_N - MT_M) mt[tid] = v;
        __syncthreads();
        // words 227..453: the far word is new (written above)
        if (tid < MT_N - MT_M) v = mt_twist(mt[227 + tid], mt[228 + tid], mt[tid]);
        __syncthreads();
        if (tid < MT_N - MT_M) mt[227 + tid] = v;
        __syncthreads();
        // words 454..622
        if (tid < 169) v = mt_twist(mt[454 + tid], mt[455 + tid], mt[227 + tid]);
        __syncthreads();
        if (tid < 169) mt[454 + tid] = v;
        __syncthreads();
        if (tid == 0) mt[623] = mt_twist(mt[623], mt[0], mt[396]);
        __syncthreads();
        uint32_t* o = out + b * MT_N;
        for (int i = tid; i < MT_N; i += 256) o[i] = mt_temper(mt[i]);
        // the next block's first phase only reads mt before its own barrier: no hazard with the loop above
    }
}

// ---- correctly rounded log(x), 2^-104 <= x < 1 ---------------------------------------------------------------------
struct DD {
    double h, l;
};
__device__ __forceinline__ DD two_sum(double a, double b) {
    const double s = __dadd_rn(a, b);
    const double bb = __dadd_rn(s, -a);
    return DD{s, __dadd_rn(__dadd_rn(a, -__dadd_rn(s, -bb)), __dadd_rn(b, -bb))};
}
__device__ __forceinline__ DD fast_two_sum(double a, double b) {      // |a| >= |b| (or a == 0)
    const double s = __dadd_rn(a, b);
    return DD{s, __dadd_rn(b, -__dadd_rn(s, -a))};
}
__device__ __forceinline__ DD dd_add(DD a, DD b) {
    DD s = two_sum(a.h, b.h);
    const DD t = two_sum(a.l, b.l);
    s = fast_two_sum(s.h, __dadd_rn(s.l, t.h));
    return fast_two_sum(s.h, __dadd_rn(s.l, t.l));
}

// x = 2^e m, m in [0.75, 1.5);  log x = e ln2 + logc[i] + log1p(r),  r = m invc[i] - 1 exactly in two doubles,
// |r| < 0.0106;  log1p(r) = r - r^2/2 (double-double) + r^3 (1/3 - r/4 + ... + r^10/13) (double: <= 2^-14 of the result).
__device__ __forceinline__ double log_cr(double x, bool& hard) {
    const long long bits = __double_as_longlong(x);
    int e = (int)((bits >> 52) & 0x7ff) - 1023;
    double m = __longlong_as_double((bits & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);
    if (m >= 1.5) {
        m = __dmul_rn(m, 0.5);
        e += 1;
    }
    const int i = (int)__dadd_rn(__dmul_rn(__dadd_rn(m, -0.75), 64.0), 0.5);
    const double invc = c_logt_invc[i];
    const double ph = __dmul_rn(m, invc);
    const double pl = __fma_rn(m, invc, -ph);
    const DD r = two_sum(__dadd_rn(ph, -1.0), pl);
    const double z = r.h;
    const double t2h = __dmul_rn(z, z);
    const double t2l = __dadd_rn(__fma_rn(z, z, -t2h), __dmul_rn(__dmul_rn(2.0, z), r.l));
    double q = 1.0 / 13.0;
    q = __fma_rn(q, z, -1.0 / 12.0);
    q = __fma_rn(q, z, 1.0 / 11.0);
    q = __fma_rn(q, z, -1.0 / 10.0);
    q = __fma_rn(q, z, 1.0 / 9.0);
    q = __fma_rn(q, z, -1.0 / 8.0);
    q = __fma_rn(q, z, 1.0 / 7.0);
    q = __fma_rn(q, z, -1.0 / 6.0);
    q = __fma_rn(q, z, 1.0 / 5.0);
    q = __fma_rn(q, z, -1.0 / 4.0);
    q = __fma_rn(q, z, 1.0 / 3.0);
    q = __dmul_rn(q, __dmul_rn(t2h, z));
    DD s = dd_add(r, DD{__dmul_rn(-0.5, t2h), __dmul_rn(-0.5, t2l)});
    s = dd_add(s, DD{q, 0.0});
    const double ed = (double)e;
    DD t = two_sum(__dmul_rn(ed, LOGT_LN2H), __dmul_rn(ed, LOGT_LN2L));      // e ln2h is exact
    t = dd_add(t, DD{c_logt_hi[i], c_logt_lo[i]});
    t = dd_add(t, s);
    // t.h = the correctly rounded logarithm unless t.l is (almost) half a unit in the last place of t.h
    const long long hb = __double_as_longlong(t.h);
    const double ulp = __longlong_as_double(((hb >> 52) & 0x7ff) - 52 << 52);
    hard = (fabs(t.l) >= 0.47 * ulp) || ((hb & 0x000fffffffffffffLL) == 0);
    return t.h;
}

// Fix-up list (device): i64 head[2] = {flagged attempts, 0} | f64 r2[cap] | f64 lg[cap] | f64 x1[cap] | f64 x2[cap] | i64 e0[cap].
// The host reads r2[0 .. count), writes libm's log of it to lg[0 .. count), and fixup_apply_kernel rewrites the normals.
struct FixupView {
    long long* head;
    double *r2, *lg, *x1, *x2;
    long long* e0;
    long long cap;
};
__host__ __device__ inline FixupView fixup_view(void* p, long long cap) {
    FixupView v;
    v.head = (long long*)p;
    v.r2 = (double*)(v.head + 2);
    v.lg = v.r2 + cap;
    v.x1 = v.lg + cap;
    v.x2 = v.x1 + cap;
    v.e0 = (long long*)(v.x2 + cap);
    v.cap = cap;
    return v;
}

__device__ __forceinline__ double legacy_double(uint32_t a, uint32_t b) {
    return __dmul_rn(__dadd_rn(__dmul_rn((double)(a >> 5), 67108864.0), (double)(b >> 6)), 1.0 / 9007199254740992.0);
}

// attempt k uses outputs 4k .. 4k+3; returns acceptance and (x1, x2, r2)
__device__ __forceinline__ bool polar_attempt(const uint32_t* __restrict__ u, long long k, double& x1, double& x2,
                                              double& r2) {
    const uint4 w = *reinterpret_cast<const uint4*>(u + 4 * k);
    x1 = __dadd_rn(__dmul_rn(2.0, legacy_double(w.x, w.y)), -1.0);
    x2 = __dadd_rn(__dmul_rn(2.0, legacy_double(w.z, w.w)), -1.0);
    r2 = __dadd_rn(__dmul_rn(x1, x1), __dmul_rn(x2, x2));
    return !(r2 >= 1.0 || r2 == 0.0);
}

__global__ void __launch_bounds__(PL_THREADS)
polar_count_kernel(const uint32_t* __restrict__ u, long long nattempts, int32_t* __restrict__ block_counts) {
    const long long k = (long long)blockIdx.x * PL_THREADS + threadIdx.x;
    double x1, x2, r2;
    const bool acc = (k < nattempts) && polar_attempt(u, k, x1, x2, r2);
    const int c = __syncthreads_count(acc);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
}

// exclusive scan of the per-CTA counts (one CTA, running carry); offsets[nb] = total
__global__ void __launch_bounds__(1024)
polar_scan_kernel(const int32_t* __restrict__ counts, long long nb, long long* __restrict__ offsets) {
    __shared__ long long warp_tot[32];
    __shared__ long long carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (long long base = 0; base < nb; base += 1024) {
        const long long i = base + tid;
        const long long v = (i < nb) ? counts[i] : 0;
        long long x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += t;
        }
        if (lane == 31) warp_tot[warp] = x;
        __syncthreads();
        if (warp == 0) {
            long long w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_tot[lane] = w;     // inclusive over warps
        }
        __syncthreads();
        const long long before = carry + (warp ? warp_tot[warp - 1] : 0) + (x - v);
        if (i < nb) offsets[i] = before;
        __syncthreads();
        if (tid == 1023) carry = before + v;
        __syncthreads();
    }
    if (tid == 0) offsets[nb] = carry;
}

__global__ void __launch_bounds__(PL_THREADS)
polar_write_kernel(const uint32_t* __restrict__ u, long long nattempts, const long long* __restrict__ offsets,
                   long long total_elems, int n, int kcols, long long s0, long long S_loc, double* __restrict__ Zt,
                   FixupView fix) {
    __shared__ int warp_cnt[PL_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long k = (long long)blockIdx.x * PL_THREADS + tid;
    double x1 = 0.0, x2 = 0.0, r2 = 0.5;
    const bool acc = (k < nattempts) && polar_attempt(u, k, x1, x2, r2);
    const unsigned bal = __ballot_sync(0xffffffffu, acc);
    if (lane == 0) warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int before = 0;
    for (int w2 = 0; w2 < warp; ++w2) before += warp_cnt[w2];
    const long long q = offsets[blockIdx.x] + before + __popc(bal & ((1u << lane) - 1u));
    if (!acc) return;
    const long long e0 = 2 * q;
    if (e0 >= total_elems) return;
    bool hard;
    const double f = __dsqrt_rn(__ddiv_rn(__dmul_rn(-2.0, log_cr(r2, hard)), r2));
    if (hard && fix.head != nullptr) {       // the host redoes this attempt with libm's log (head[0] counts all of them)
        const long long pos = (long long)atomicAdd((unsigned long long*)fix.head, 1ull);
        if (pos < fix.cap) {
            fix.r2[pos] = r2;
            fix.x1[pos] = x1;
            fix.x2[pos] = x2;
            fix.e0[pos] = e0;
        }
    }
    // element e of standard_normal((S, n)) in C order: sample s = e / n, grid column j = e % n
    {
        const long long s = e0 / n;
        const int j = (int)(e0 - s * n);
        if (j < kcols && s >= s0 && s < s0 + S_loc) Zt[(size_t)j * S_loc + (s - s0)] = __dmul_rn(f, x2);
    }
    const long long e1 = e0 + 1;
    if (e1 < total_elems) {
        const long long s = e1 / n;
        const int j = (int)(e1 - s * n);
        if (j < kcols && s >= s0 && s < s0 + S_loc) Zt[(size_t)j * S_loc + (s - s0)] = __dmul_rn(f, x1);
    }
}

// the flagged attempts again, with the logarithm the host's libm returned
__global__ void __launch_bounds__(256)
fixup_apply_kernel(FixupView fix, long long count, long long total_elems, int n, int kcols, long long s0, long long S_loc,
                   double* __restrict__ Zt) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const double r2 = fix.r2[k];
    const double f = __dsqrt_rn(__ddiv_rn(__dmul_rn(-2.0, fix.lg[k]), r2));
    const long long e0 = fix.e0[k];
    {
        const long long s = e0 / n;
        const int j = (int)(e0 - s * n);
        if (j < kcols && s >= s0 && s < s0 + S_loc) Zt[(size_t)j * S_loc + (s - s0)] = __dmul_rn(f, fix.x2[k]);
    }
    const long long e1 = e0 + 1;
    if (e1 < total_elems) {
        const long long s = e1 / n;
        const int j = (int)(e1 - s * n);
        if (j < kcols && s >= s0 && s < s0 + S_loc) Zt[(size_t)j * S_loc + (s - s0)] = __dmul_rn(f, fix.x1[k]);
    }
}

__global__ void polar_check_kernel(const long long* __restrict__ offsets, long long nb, long long need_pairs,
                                   int32_t* __restrict__ ok) {
    ok[0] = offsets[nb] >= need_pairs ? 1 : 0;
}

static void plan(long long S, int n, long long& nattempts, long long& nmt, long long& ncount) {
    const long long pairs = (S * n + 1) / 2;
    // acceptance probability pi/4; mean + 8 sigma + slack attempts
    const double mean = (double)pairs / 0.7853981633974483;
    nattempts = (long long)(mean + 8.0 * sqrt((double)pairs * 0.2146 / (0.7853981633974483 * 0.7853981633974483)) + 256.0);
    nmt = (4 * nattempts + MT_N - 1) / MT_N;
    nattempts = nmt * MT_N / 4;                       // use every generated word
    ncount = (nattempts + PL_THREADS - 1) / PL_THREADS;
}

}  // namespace gpet

using namespace gpet;

extern "C" int64_t gpet_standard_normal_workspace_bytes(int64_t S, int n) {
    long long na, nmt, nc;
    plan(S, n, na, nmt, nc);
    return nmt * MT_N * 4 + 16 + (nc + 1) * 4 + 16 + (nc + 2) * 8 + 256;
}

static long long fixup_capacity(int64_t S, int n) {
    const long long pairs = (S * n + 1) / 2;
    return pairs / 10 + 4096;          // ~6 % of the attempts are flagged
}

extern "C" int64_t gpet_standard_normal_fixup_bytes(int64_t S, int n) { return 16 + fixup_capacity(S, n) * 40; }

// libm's own log on the host (what numpy's legacy generator calls), for the flagged attempts; large counts are split
// over a few threads
extern "C" int gpet_host_log_f64(const double* x, double* out, int64_t count) {
    GPET_REQUIRE(count >= 0 && (count == 0 || (x && out)), "gpet_host_log_f64: bad argument");
    unsigned nt = count >= 200000 ? std::min(8u, std::max(1u, std::thread::hardware_concurrency())) : 1u;
    if (nt <= 1) {
        for (int64_t i = 0; i < count; ++i) out[i] = log(x[i]);
        return GPET_OK;
    }
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nt; ++t) {
        const int64_t a = count * t / nt, b = count * (t + 1) / nt;
        pool.emplace_back([=]() { for (int64_t i = a; i < b; ++i) out[i] = log(x[i]); });
    }
    for (auto& th : pool) th.join();
    return GPET_OK;
}

// Second half of an exact draw: the host has read r2[0 .. count) of the fix-up list and stored log(r2) in lg[0 .. count).
extern "C" int gpet_standard_normal_fixup_apply_f64(int64_t S, int n, int kcols, int64_t s0, int64_t S_loc, double* Zt,
                                                    void* fixups, int64_t count, void* stream) {
    GPET_REQUIRE(Zt && fixups && S > 0 && n > 0 && kcols > 0 && kcols <= n && s0 >= 0 && S_loc > 0 && s0 + S_loc <= S,
                 "gpet_standard_normal_fixup_apply_f64: bad argument");
    const long long cap = fixup_capacity(S, n);
    GPET_REQUIRE(count >= 0 && count <= cap, "gpet_standard_normal_fixup_apply_f64: %lld records, room for %lld", (long long)count, cap);
    if (count == 0) return GPET_OK;
    fixup_apply_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(fixup_view(fixups, cap), count,
                                                                                       (long long)S * n, n, kcols, s0, S_loc, Zt);
    return check_launch("fixup_apply_kernel");
}

extern "C" int gpet_standard_normal_t_f64(uint32_t seed, int64_t S, int n, int kcols, int64_t s0, int64_t S_loc, double* Zt,
                                          int32_t* ok, void* fixups, void* work, void* stream) {
    GPET_REQUIRE(Zt && ok && work && S > 0 && n > 0 && kcols > 0 && kcols <= n && s0 >= 0 && S_loc > 0 && s0 + S_loc <= S,
                 "gpet_standard_normal_t_f64: bad argument");
    long long na, nmt, nc;
    plan(S, n, na, nmt, nc);
    GPET_SUPPORTED(nc < 0x7fffffffLL, "gpet_standard_normal_t_f64: too many draws for one call");
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t* u = (uint32_t*)work;
    size_t off = ((size_t)nmt * MT_N * 4 + 15) & ~(size_t)15;
    int32_t* counts = (int32_t*)((char*)work + off);
    off = (off + (size_t)(nc + 1) * 4 + 15) & ~(size_t)15;
    long long* offsets = (long long*)((char*)work + off);
    mt19937_stream_kernel<<<1, 256, 0, st>>>(seed, nmt, u);
    polar_count_kernel<<<(unsigned)nc, PL_THREADS, 0, st>>>(u, na, counts);
    polar_scan_kernel<<<1, 1024, 0, st>>>(counts, nc, offsets);
    FixupView fix{};
    if (fixups != nullptr) {
        fix = fixup_view(fixups, fixup_capacity(S, n));
        cudaError_t e = cudaMemsetAsync(fixups, 0, 16, st);
        if (e != cudaSuccess) {
            set_error("gpet_standard_normal_t_f64: fix-up header: %s", cudaGetErrorString(e));
            return GPET_ERR_CUDA;
        }
    }
    polar_write_kernel<<<(unsigned)nc, PL_THREADS, 0, st>>>(u, na, offsets, (long long)S * n, n, kcols, s0, S_loc, Zt, fix);
    polar_check_kernel<<<1, 1, 0, st>>>(offsets, nc, ((long long)S * n + 1) / 2, ok);
    return check_launch("gpet_standard_normal_t_f64");
}
