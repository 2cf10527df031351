// L-BFGS-B state machines of the final hyper-parameter fit (sklearn_gpr.py:254-295, 587-607: 13 runs of
// scipy.optimize.minimize(method='L-BFGS-B') per trace), advanced on the device in lock step with the objective
// kernel (gpet_lml_f64): one THREAD per run, the state of run e interleaved with stride E (gpet_lbfgsb.cuh), so the
// whole fit is a chain [advance -> objective] of kernel launches with no host arithmetic and a 4-byte read-back per
// round (the number of runs still active).  The host entry points below run the same code on contiguous state and
// exist for the CPU differential test against scipy's own setulb (tests/test_lbfgsb_host.py).
#include "gpet_common.cuh"
#include "gpet_lbfgsb.cuh"
#include "gpet_b200.h"

namespace gpet {

using namespace gpet_lb;

__global__ void __launch_bounds__(64)
lbfgsb_init_kernel(double* __restrict__ dstate, int32_t* __restrict__ istate, int E, const double* __restrict__ x0,
                   const double* __restrict__ lo, const double* __restrict__ up) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    Solver<MemStrided> s(MemStrided{dstate + e, istate + e, (long long)E});
    const double xs[3] = {x0[3 * e], x0[3 * e + 1], x0[3 * e + 2]};
    const double l3[3] = {lo[0], lo[1], lo[2]}, u3[3] = {up[0], up[1], up[2]};
    s.init(xs, l3, u3);
    s.store();
}

// One round: runs whose objective was evaluated in the previous round take (f, g) and advance to their next
// evaluation point or to the end.  trace_eval[e] = trace_of[e] while run e waits for an evaluation at theta[e][3],
// -1 once it has ended (the objective kernel skips those).  n_active[0] counts the waiting runs of this round (reset by
// the caller), n_active[1] all evaluations requested so far, n_active[2] the rounds that had at least one waiting run.
__global__ void __launch_bounds__(128)
lbfgsb_advance_kernel(double* __restrict__ dstate, int32_t* __restrict__ istate, int E, int first,
                      const int32_t* __restrict__ trace_of, const double* __restrict__ f, const double* __restrict__ g,
                      double* __restrict__ theta, int32_t* __restrict__ trace_eval, int32_t* __restrict__ n_active) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    if (!first && trace_eval[e] < 0) return;
    Solver<MemStrided> s(MemStrided{dstate + e, istate + e, (long long)E});
    s.load();
    if (!first) {
        s.F() = f[e];
        s.G(1) = g[3 * e];
        s.G(2) = g[3 * e + 1];
        s.G(3) = g[3 * e + 2];
        ++s.nfev;
    }
    const bool need = s.advance();
    s.store();
    if (need) {
        theta[3 * e] = s.X(1);
        theta[3 * e + 1] = s.X(2);
        theta[3 * e + 2] = s.X(3);
        trace_eval[e] = trace_of[e];
        if (atomicAdd(n_active, 1) == 0) atomicAdd(n_active + 2, 1);    // first waiting run of this round: one more round with work
        atomicAdd(n_active + 1, 1);                                     // evaluations so far
    } else {
        trace_eval[e] = -1;
    }
}

__global__ void __launch_bounds__(64)
lbfgsb_result_kernel(const double* __restrict__ dstate, const int32_t* __restrict__ istate, int E,
                     double* __restrict__ x, double* __restrict__ fval, int32_t* __restrict__ nfev,
                     int32_t* __restrict__ task) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    MemStrided mem{const_cast<double*>(dstate) + e, const_cast<int32_t*>(istate) + e, (long long)E};
    x[3 * e] = mem.d(O_X);
    x[3 * e + 1] = mem.d(O_X + 1);
    x[3 * e + 2] = mem.d(O_X + 2);
    fval[e] = mem.d(O_F);
    nfev[e] = mem.i(I_NFEV);
    task[e] = mem.i(I_TASK);
}


// ---- one run per WARP (lane 0 works), state of run e contiguous at dstate + e * ND ------------------------------------
// ncu on the thread-per-run kernel (profiles/r01_lbfgsb_advance_ncu.csv): 2.35 of 32 lanes active per issued
// instruction - the runs of a warp are in different phases (line-search step / new iteration, different numbers of
// correction pairs and breakpoints), so a warp walks through up to 32 different paths one after the other.  With one
// run per warp nothing diverges, 32 times more warps hide each other's latency, and a run's arrays are contiguous
// (a 128-byte line holds 16 consecutive elements instead of one).
constexpr int LBW_WARPS = 8;      // runs per CTA

__global__ void __launch_bounds__(32 * LBW_WARPS)
lbfgsb_init_warp_kernel(double* __restrict__ dstate, int32_t* __restrict__ istate, int E, const double* __restrict__ x0,
                        const double* __restrict__ lo, const double* __restrict__ up) {
    const int e = blockIdx.x * LBW_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (e >= E) return;
    double* dp = dstate + (size_t)e * ND;
    int32_t* ip = istate + (size_t)e * NI;
    for (int k = lane; k < ND; k += 32) dp[k] = 0.0;
    for (int k = lane; k < NI; k += 32) ip[k] = 0;
    __syncwarp();
    if (lane != 0) return;
    Solver<MemFlat> s(MemFlat{dp, ip});
    const double xs[3] = {x0[3 * e], x0[3 * e + 1], x0[3 * e + 2]};
    const double l3[3] = {lo[0], lo[1], lo[2]}, u3[3] = {up[0], up[1], up[2]};
    s.init(xs, l3, u3);
    s.store();
}

__global__ void __launch_bounds__(32 * LBW_WARPS, 1024 / (32 * LBW_WARPS))     // 64 registers, 8 warps per scheduler
lbfgsb_advance_warp_kernel(double* __restrict__ dstate, int32_t* __restrict__ istate, int E, int first,
                           const int32_t* __restrict__ trace_of, const double* __restrict__ f,
                           const double* __restrict__ g, double* __restrict__ theta, int32_t* __restrict__ trace_eval,
                           int32_t* __restrict__ n_active) {
    const int e = blockIdx.x * LBW_WARPS + (threadIdx.x >> 5);
    if ((threadIdx.x & 31) != 0 || e >= E) return;
    if (!first && trace_eval[e] < 0) return;
    Solver<MemFlat> s(MemFlat{dstate + (size_t)e * ND, istate + (size_t)e * NI});
    s.load();
    if (!first) {
        s.F() = f[e];
        s.G(1) = g[3 * e];
        s.G(2) = g[3 * e + 1];
        s.G(3) = g[3 * e + 2];
        ++s.nfev;
    }
    const bool need = s.advance();
    s.store();
    if (need) {
        theta[3 * e] = s.X(1);
        theta[3 * e + 1] = s.X(2);
        theta[3 * e + 2] = s.X(3);
        trace_eval[e] = trace_of[e];
        if (atomicAdd(n_active, 1) == 0) atomicAdd(n_active + 2, 1);
        atomicAdd(n_active + 1, 1);
    } else {
        trace_eval[e] = -1;
    }
}

__global__ void __launch_bounds__(64)
lbfgsb_result_warp_kernel(const double* __restrict__ dstate, const int32_t* __restrict__ istate, int E,
                          double* __restrict__ x, double* __restrict__ fval, int32_t* __restrict__ nfev,
                          int32_t* __restrict__ task) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const double* dp = dstate + (size_t)e * ND;
    const int32_t* ip = istate + (size_t)e * NI;
    x[3 * e] = dp[O_X];
    x[3 * e + 1] = dp[O_X + 1];
    x[3 * e + 2] = dp[O_X + 2];
    fval[e] = dp[O_F];
    nfev[e] = ip[I_NFEV];
    task[e] = ip[I_TASK];
}

}  // namespace gpet

using namespace gpet;
using namespace gpet_lb;

extern "C" int64_t gpet_lbfgsb_state_doubles(void) { return ND; }
extern "C" int64_t gpet_lbfgsb_state_ints(void) { return NI; }

extern "C" int gpet_lbfgsb_init_f64(double* dstate, int32_t* istate, int E, const double* x0, const double* lo,
                                    const double* up, void* stream) {
    GPET_REQUIRE(dstate && istate && x0 && lo && up && E > 0, "gpet_lbfgsb_init_f64: bad argument");
    if (g_tune[GPET_TUNE_LBFGSB_THREADS] == 0) {     // one run per warp, contiguous state
        lbfgsb_init_warp_kernel<<<(E + LBW_WARPS - 1) / LBW_WARPS, 32 * LBW_WARPS, 0, (cudaStream_t)stream>>>(dstate, istate,
                                                                                                         E, x0, lo, up);
        return check_launch("lbfgsb_init_warp_kernel");
    }
    lbfgsb_init_kernel<<<(E + 63) / 64, 64, 0, (cudaStream_t)stream>>>(dstate, istate, E, x0, lo, up);
    return check_launch("lbfgsb_init_kernel");
}

extern "C" int gpet_lbfgsb_advance_f64(double* dstate, int32_t* istate, int E, int first, const int32_t* trace_of,
                                       const double* f, const double* g, double* theta, int32_t* trace_eval,
                                       int32_t* n_active, void* stream) {
    GPET_REQUIRE(dstate && istate && trace_of && f && g && theta && trace_eval && n_active && E > 0,
                 "gpet_lbfgsb_advance_f64: bad argument");
    cudaError_t err = cudaMemsetAsync(n_active, 0, sizeof(int32_t), (cudaStream_t)stream);
    if (err != cudaSuccess) {
        set_error("lbfgsb advance memset: %s", cudaGetErrorString(err));
        return GPET_ERR_CUDA;
    }
    // One run per thread, serial and latency bound (up to ~1.3 ms per round once the 10 correction pairs are in use,
    // whatever the CTA size: 1.28 / 1.28 / 1.28 ms at 32 / 64 / 128 threads, tools/bench_lbfgsb.py), so all runs must be
    // resident at once.  Keeping the 2m x 2m factor of formk in shared memory (3.2 KB per thread) was measured SLOWER
    // (1.9 ms per round at E = 16250): it limits an SM to 64 runs and the launch then needs two waves.
    int nt = g_tune[GPET_TUNE_LBFGSB_THREADS];
    if (nt == 0) {
        lbfgsb_advance_warp_kernel<<<(E + LBW_WARPS - 1) / LBW_WARPS, 32 * LBW_WARPS, 0, (cudaStream_t)stream>>>(
            dstate, istate, E, first, trace_of, f, g, theta, trace_eval, n_active);
        return check_launch("lbfgsb_advance_warp_kernel");
    }
    nt = nt <= 32 ? 32 : (nt <= 64 ? 64 : 128);
    lbfgsb_advance_kernel<<<(E + nt - 1) / nt, nt, 0, (cudaStream_t)stream>>>(dstate, istate, E, first, trace_of, f, g, theta,
                                                                            trace_eval, n_active);
    return check_launch("lbfgsb_advance_kernel");
}

// n_rounds x [advance -> objective] enqueued back to back, then one copy of the three counters to (pinned) host memory.
extern "C" int gpet_fit_rounds_f64(const double* X, const double* y, const double* w, const int32_t* xcol,
                                   const int32_t* m, int mmax, int kind, double gp_alpha, double* dstate,
                                   int32_t* istate, int E, int first, int n_rounds, const int32_t* trace_of, double* f,
                                   double* g, double* theta, int32_t* trace_eval, int32_t* counters,
                                   int32_t* counters_host, void* stream) {
    GPET_REQUIRE(n_rounds > 0 && counters && counters_host, "gpet_fit_rounds_f64: bad argument");
    for (int r = 0; r < n_rounds; ++r) {
        int rc = gpet_lbfgsb_advance_f64(dstate, istate, E, (first && r == 0) ? 1 : 0, trace_of, f, g, theta, trace_eval,
                                         counters, stream);
        if (rc != GPET_OK) return rc;
        rc = gpet_lml_f64(X, y, w, xcol, m, mmax, trace_eval, theta, E, kind, gp_alpha, f, g, stream);
        if (rc != GPET_OK) return rc;
    }
    cudaError_t err = cudaMemcpyAsync(counters_host, counters, 3 * sizeof(int32_t), cudaMemcpyDeviceToHost,
                                      (cudaStream_t)stream);
    if (err != cudaSuccess) {
        set_error("gpet_fit_rounds_f64 copy: %s", cudaGetErrorString(err));
        return GPET_ERR_CUDA;
    }
    return GPET_OK;
}

extern "C" int gpet_lbfgsb_result_f64(const double* dstate, const int32_t* istate, int E, double* x, double* fval,
                                      int32_t* nfev, int32_t* task, void* stream) {
    GPET_REQUIRE(dstate && istate && x && fval && nfev && task && E > 0, "gpet_lbfgsb_result_f64: bad argument");
    if (g_tune[GPET_TUNE_LBFGSB_THREADS] == 0) {
        lbfgsb_result_warp_kernel<<<(E + 63) / 64, 64, 0, (cudaStream_t)stream>>>(dstate, istate, E, x, fval, nfev, task);
        return check_launch("lbfgsb_result_warp_kernel");
    }
    lbfgsb_result_kernel<<<(E + 63) / 64, 64, 0, (cudaStream_t)stream>>>(dstate, istate, E, x, fval, nfev, task);
    return check_launch("lbfgsb_result_kernel");
}

// ---- host twins (contiguous state: run e at dstate + e * ND, istate + e * NI) -----------------------------------------
extern "C" int gpet_lbfgsb_host_init(double* dstate, int32_t* istate, int E, const double* x0, const double* lo,
                                     const double* up) {
    GPET_REQUIRE(dstate && istate && x0 && lo && up && E > 0, "gpet_lbfgsb_host_init: bad argument");
    for (int e = 0; e < E; ++e) {
        Solver<MemFlat> s(MemFlat{dstate + (size_t)e * ND, istate + (size_t)e * NI});
        s.init(x0 + 3 * e, lo, up);
        s.store();
    }
    return GPET_OK;
}

// give[e] != 0: run e takes (f[e], g[e][3]) first.  need[e] = 1 if run e now waits for an evaluation at x[e][3].
extern "C" int gpet_lbfgsb_host_advance(double* dstate, int32_t* istate, int E, const int32_t* give, const double* f,
                                        const double* g, int32_t* need, double* x) {
    GPET_REQUIRE(dstate && istate && give && f && g && need && x && E > 0, "gpet_lbfgsb_host_advance: bad argument");
    for (int e = 0; e < E; ++e) {
        Solver<MemFlat> s(MemFlat{dstate + (size_t)e * ND, istate + (size_t)e * NI});
        s.load();
        if (s.task >= T_CONV) { need[e] = 0; continue; }
        if (give[e]) {
            s.F() = f[e];
            for (int i = 0; i < 3; ++i) s.G(i + 1) = g[3 * e + i];
            ++s.nfev;
        }
        need[e] = s.advance() ? 1 : 0;
        s.store();
        for (int i = 0; i < 3; ++i) x[3 * e + i] = s.X(i + 1);
    }
    return GPET_OK;
}
