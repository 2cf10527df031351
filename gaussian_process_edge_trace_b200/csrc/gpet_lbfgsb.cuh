// L-BFGS-B (Byrd, Lu, Nocedal, Zhu 1995; version 3.0 of Morales & Nocedal 2011 with the projected subspace step)
// for the final hyper-parameter fit: n = 3 variables (log constant, log length-scale, log noise), every variable
// bounded on both sides, m = 10 correction pairs - the constants scipy.optimize.minimize(method='L-BFGS-B') uses when
// the reference calls it at sklearn_gpr.py:589-595 (ftol = 2.22e-9, gtol = 1e-5, maxls = 20, maxiter = maxfun = 15000).
//
// Written from the published algorithm as a reverse-communication state machine (like scipy's `setulb`): `advance`
// runs one instance until it needs the objective and gradient at `x`, or stops.  The same source compiles for the
// host (tests/: differential test against scipy's own setulb, instance by instance, iterate by iterate) and for the
// device (gpet_finalfit.cu: one thread per instance, the state of instance i interleaved with stride E so that a
// warp's accesses coalesce).  Every decision (breakpoint order, free-set changes, skipped updates, memory refreshes,
// line-search cases, stopping tests) follows the algorithm scipy runs.  Rounding: a one-ulp difference decides whether
// a step that runs to a bound lands on it, so the products scipy hands to BLAS (ddot / daxpy: fused multiply-adds on
// current CPUs) are written as fma chains and everything else must stay unfused - compile this file with -fmad=false.
// The small factorizations (LAPACK in scipy's build) may still sum in another order: iterates agree to rounding.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define LB_HD __host__ __device__ __forceinline__
#else
#define LB_HD inline
#endif

namespace gpet_lb {

constexpr int N = 3, M = 10, M2 = 20;
constexpr double FTOL = 2.220446049250313e-09;   // factr * epsmch
constexpr double PGTOL = 1e-5;
constexpr double EPSMCH = 2.220446049250313e-16;
constexpr int MAXLS = 20, MAXITER = 15000, MAXFUN = 15000;

// offsets into the per-instance double state
constexpr int O_X = 0, O_F = 3, O_G = 4, O_L = 7, O_U = 10, O_WS = 13, O_WY = O_WS + N * M, O_SY = O_WY + N * M,
              O_SS = O_SY + M * M, O_WT = O_SS + M * M, O_WN = O_WT + M * M, O_SND = O_WN + M2 * M2,
              O_Z = O_SND + M2 * M2, O_R = O_Z + N, O_D = O_R + N, O_T = O_D + N, O_XP = O_T + N, O_WA = O_XP + N,
              O_SC = O_WA + 8 * M, N_SC = 24, ND = O_SC + N_SC;
// offsets into the per-instance int state
constexpr int I_INDEX = 0, I_IWHERE = 3, I_INDX2 = 6, I_SC = 9, N_ISC = 26, NI = I_SC + N_ISC;
constexpr int I_TASK = I_SC + 20, I_NFEV = I_SC + 22;     // positions of `task` and `nfev` (see load/store)

enum Task { T_START = 0, T_FG_START = 1, T_FG_LN = 2, T_NEW_X = 3, T_CONV = 4, T_ABNORMAL = 5, T_STOP = 6 };
enum Ls { LS_START = 0, LS_FG = 1, LS_CONV = 2, LS_WARN = 3 };

// contiguous (host) and strided (device) views of one instance
struct MemFlat {
    double* dp;
    int* ip;
    LB_HD double& d(int k) const { return dp[k]; }
    LB_HD int& i(int k) const { return ip[k]; }
};
struct MemStrided {
    double* dp;
    int* ip;
    long long stride;
    LB_HD double& d(int k) const { return dp[(long long)k * stride]; }
    LB_HD int& i(int k) const { return ip[(long long)k * stride]; }
};

template <class A>
struct Solver {
    A a;
    // scalars, kept in registers between load() and store()
    double theta, fold, dnorm, gd, stpmx, sbgnrm, stp, gdold, dtd;
    double ginit, gtest, gx, gy, finit, fx, fy, stx, sty, stmin, stmax, width, width1;
    int head, col, itail, iter, iupdat, nseg, nfgv, info, ifun, iword, nfree, nact, ileave, nenter, iback, nskip;
    int updatd, brackt, stage, ls, task, nit, nfev, wrk;

    LB_HD explicit Solver(const A& acc) : a(acc) {}

    // ---- 1-based views (as in the published pseudo-code / Fortran) -------------------------------------------------
    LB_HD double& X(int i) const { return a.d(O_X + i - 1); }
    LB_HD double& G(int i) const { return a.d(O_G + i - 1); }
    LB_HD double& F() const { return a.d(O_F); }
    LB_HD double& Lo(int i) const { return a.d(O_L + i - 1); }
    LB_HD double& Up(int i) const { return a.d(O_U + i - 1); }
    LB_HD double& WS(int i, int j) const { return a.d(O_WS + (j - 1) * N + i - 1); }
    LB_HD double& WY(int i, int j) const { return a.d(O_WY + (j - 1) * N + i - 1); }
    LB_HD double& SY(int i, int j) const { return a.d(O_SY + (j - 1) * M + i - 1); }
    LB_HD double& SS(int i, int j) const { return a.d(O_SS + (j - 1) * M + i - 1); }
    LB_HD double& WT(int i, int j) const { return a.d(O_WT + (j - 1) * M + i - 1); }
    LB_HD double& WN(int i, int j) const { return a.d(O_WN + (j - 1) * M2 + i - 1); }
    LB_HD double& WN1(int i, int j) const { return a.d(O_SND + (j - 1) * M2 + i - 1); }
    LB_HD double& Z(int i) const { return a.d(O_Z + i - 1); }
    LB_HD double& R(int i) const { return a.d(O_R + i - 1); }
    LB_HD double& D(int i) const { return a.d(O_D + i - 1); }
    LB_HD double& T(int i) const { return a.d(O_T + i - 1); }
    LB_HD double& XP(int i) const { return a.d(O_XP + i - 1); }
    LB_HD double& WA(int i) const { return a.d(O_WA + i - 1); }
    LB_HD int& INDEX(int i) const { return a.i(I_INDEX + i - 1); }
    LB_HD int& IWHERE(int i) const { return a.i(I_IWHERE + i - 1); }
    LB_HD int& INDX2(int i) const { return a.i(I_INDX2 + i - 1); }

    LB_HD void load() {
        const int o = O_SC;
        theta = a.d(o + 0); fold = a.d(o + 1); dnorm = a.d(o + 2); gd = a.d(o + 3); stpmx = a.d(o + 4);
        sbgnrm = a.d(o + 5); stp = a.d(o + 6); gdold = a.d(o + 7); dtd = a.d(o + 8);
        ginit = a.d(o + 9); gtest = a.d(o + 10); gx = a.d(o + 11); gy = a.d(o + 12); finit = a.d(o + 13);
        fx = a.d(o + 14); fy = a.d(o + 15); stx = a.d(o + 16); sty = a.d(o + 17); stmin = a.d(o + 18);
        stmax = a.d(o + 19); width = a.d(o + 20); width1 = a.d(o + 21);
        const int q = I_SC;
        head = a.i(q + 0); col = a.i(q + 1); itail = a.i(q + 2); iter = a.i(q + 3); iupdat = a.i(q + 4);
        nseg = a.i(q + 5); nfgv = a.i(q + 6); info = a.i(q + 7); ifun = a.i(q + 8); iword = a.i(q + 9);
        nfree = a.i(q + 10); nact = a.i(q + 11); ileave = a.i(q + 12); nenter = a.i(q + 13); iback = a.i(q + 14);
        nskip = a.i(q + 15); updatd = a.i(q + 16); brackt = a.i(q + 17); stage = a.i(q + 18); ls = a.i(q + 19);
        task = a.i(q + 20); nit = a.i(q + 21); nfev = a.i(q + 22); wrk = a.i(q + 23);
    }
    LB_HD void store() const {
        const int o = O_SC;
        a.d(o + 0) = theta; a.d(o + 1) = fold; a.d(o + 2) = dnorm; a.d(o + 3) = gd; a.d(o + 4) = stpmx;
        a.d(o + 5) = sbgnrm; a.d(o + 6) = stp; a.d(o + 7) = gdold; a.d(o + 8) = dtd;
        a.d(o + 9) = ginit; a.d(o + 10) = gtest; a.d(o + 11) = gx; a.d(o + 12) = gy; a.d(o + 13) = finit;
        a.d(o + 14) = fx; a.d(o + 15) = fy; a.d(o + 16) = stx; a.d(o + 17) = sty; a.d(o + 18) = stmin;
        a.d(o + 19) = stmax; a.d(o + 20) = width; a.d(o + 21) = width1;
        const int q = I_SC;
        a.i(q + 0) = head; a.i(q + 1) = col; a.i(q + 2) = itail; a.i(q + 3) = iter; a.i(q + 4) = iupdat;
        a.i(q + 5) = nseg; a.i(q + 6) = nfgv; a.i(q + 7) = info; a.i(q + 8) = ifun; a.i(q + 9) = iword;
        a.i(q + 10) = nfree; a.i(q + 11) = nact; a.i(q + 12) = ileave; a.i(q + 13) = nenter; a.i(q + 14) = iback;
        a.i(q + 15) = nskip; a.i(q + 16) = updatd; a.i(q + 17) = brackt; a.i(q + 18) = stage; a.i(q + 19) = ls;
        a.i(q + 20) = task; a.i(q + 21) = nit; a.i(q + 22) = nfev; a.i(q + 23) = wrk;
    }

    // ---- start: x0 is projected onto the box; every variable has two finite bounds ---------------------------------
    LB_HD void init(const double* x0, const double* lo, const double* up) {
        // the whole workspace starts at zero, as scipy allocates it: the update of the 2m x 2m matrix in formk adds
        // to entries that no earlier call has written when a correction pair was stored in an iteration that skipped
        // the subspace step (no free variable at the Cauchy point), and scipy's iterates are the ones with zeros there
        for (int k = 0; k < ND; ++k) a.d(k) = 0.0;
        for (int k = 0; k < NI; ++k) a.i(k) = 0;
        for (int i = 1; i <= N; ++i) {
            Lo(i) = lo[i - 1];
            Up(i) = up[i - 1];
            double v = x0[i - 1];
            v = v < lo[i - 1] ? lo[i - 1] : (v > up[i - 1] ? up[i - 1] : v);     // np.clip of scipy's wrapper
            X(i) = v;
            G(i) = 0.0;
        }
        F() = 0.0;
        task = T_START;
        nit = nfev = 0;
        theta = 1.0; fold = dnorm = gd = stpmx = sbgnrm = stp = gdold = dtd = 0.0;
        ginit = gtest = gx = gy = finit = fx = fy = stx = sty = stmin = stmax = width = width1 = 0.0;
        head = 1; col = itail = iter = iupdat = nseg = nfgv = info = ifun = iword = nact = ileave = nenter = 0;
        iback = nskip = updatd = brackt = stage = wrk = 0;
        nfree = N;
        ls = LS_START;
    }

    LB_HD void refresh() {      // discard the correction pairs
        info = 0; col = 0; head = 1; theta = 1.0; iupdat = 0; updatd = 0;
    }

    // ---- small dense helpers on column-major upper-triangular factors -----------------------------------------------
    // Cholesky R'R of the leading nn x nn block stored at (r0, c0) of a matrix accessed through `at`; 0 or the
    // 1-based index of the failing pivot (LINPACK dpofa order)
    template <class At>
    LB_HD int chol_upper(At at, int nn) const {
        for (int j = 1; j <= nn; ++j) {
            double s = 0.0;
            for (int k = 1; k <= j - 1; ++k) {
                double dot = 0.0;
                for (int q = 1; q <= k - 1; ++q) dot = fma(at(q, k), at(q, j), dot);
                double t = at(k, j) - dot;
                t = t / at(k, k);
                at(k, j) = t;
                s += t * t;
            }
            s = at(j, j) - s;
            if (!(s > 0.0)) return j;
            at(j, j) = sqrt(s);
        }
        return 0;
    }
    // R' x = b (trans = true) or R x = b (trans = false), R upper nn x nn; b accessed through `bt` (1-based)
    template <class At, class Bt>
    LB_HD int tri_solve(At at, int nn, Bt bt, bool trans) const {
        for (int j = 1; j <= nn; ++j)
            if (at(j, j) == 0.0) return j;
        if (trans) {
            bt(1) = bt(1) / at(1, 1);
            for (int j = 2; j <= nn; ++j) {
                double dot = 0.0;
                for (int q = 1; q <= j - 1; ++q) dot = fma(at(q, j), bt(q), dot);
                bt(j) = (bt(j) - dot) / at(j, j);
            }
        } else {
            bt(nn) = bt(nn) / at(nn, nn);
            for (int jj = 2; jj <= nn; ++jj) {
                const int j = nn - jj + 1;
                const double temp = -bt(j + 1);
                for (int q = 1; q <= j; ++q) bt(q) = fma(temp, at(q, j + 1), bt(q));
                bt(j) = bt(j) / at(j, j);
            }
        }
        return 0;
    }

    // ---- projected gradient norm ---------------------------------------------------------------------------------
    LB_HD void projgr() {
        sbgnrm = 0.0;
        for (int i = 1; i <= N; ++i) {
            double gi = G(i);
            if (gi < 0.0) gi = fmax(X(i) - Up(i), gi);
            else gi = fmin(X(i) - Lo(i), gi);
            sbgnrm = fmax(sbgnrm, fabs(gi));
        }
    }

    // ---- product of the 2col x 2col middle matrix of the compact L-BFGS formula with v -> p ---------------------------
    // v, p: offsets (0-based) into WA
    LB_HD int bmv(int ov, int op) {
        if (col == 0) return 0;
        auto wt = [&](int i, int j) -> double& { return WT(i, j); };
        WA(op + col + 1) = WA(ov + col + 1);
        for (int i = 2; i <= col; ++i) {
            const int i2 = col + i;
            double sum = 0.0;
            for (int k = 1; k <= i - 1; ++k) sum += SY(i, k) * WA(ov + k) / SY(k, k);
            WA(op + i2) = WA(ov + i2) + sum;
        }
        auto p2 = [&](int i) -> double& { return WA(op + col + i); };
        if (tri_solve(wt, col, p2, true) != 0) return 1;
        for (int i = 1; i <= col; ++i) WA(op + i) = WA(ov + i) / sqrt(SY(i, i));
        if (tri_solve(wt, col, p2, false) != 0) return 1;
        for (int i = 1; i <= col; ++i) WA(op + i) = -WA(op + i) / sqrt(SY(i, i));
        for (int i = 1; i <= col; ++i) {
            double sum = 0.0;
            for (int k = i + 1; k <= col; ++k) sum += SY(k, i) * WA(op + col + k) / SY(i, i);
            WA(op + i) += sum;
        }
        return 0;
    }

    // ---- heap of breakpoints (t = T, iorder = INDX2) ---------------------------------------------------------------
    LB_HD void hpsolb(int nn, int iheap) {
        if (iheap == 0) {
            for (int k = 2; k <= nn; ++k) {
                const double ddum = T(k);
                const int indxin = INDX2(k);
                int i = k;
                while (i > 1) {
                    const int j = i / 2;
                    if (ddum < T(j)) { T(i) = T(j); INDX2(i) = INDX2(j); i = j; }
                    else break;
                }
                T(i) = ddum;
                INDX2(i) = indxin;
            }
        }
        if (nn > 1) {
            int i = 1;
            const double out = T(1);
            const int indxou = INDX2(1);
            const double ddum = T(nn);
            const int indxin = INDX2(nn);
            for (;;) {
                int j = i + i;
                if (j <= nn - 1) {
                    if (T(j + 1) < T(j)) j = j + 1;
                    if (T(j) < ddum) { T(i) = T(j); INDX2(i) = INDX2(j); i = j; }
                    else break;
                } else break;
            }
            T(i) = ddum;
            INDX2(i) = indxin;
            T(nn) = out;
            INDX2(nn) = indxou;
        }
    }

    // ---- generalized Cauchy point -> Z; p, c, wbp, v = WA(1..), WA(2M+1..), WA(4M+1..), WA(6M+1..) ------------------
    LB_HD void cauchy() {
        const int OP = 0, OC = 2 * M, OW = 4 * M, OV = 6 * M;
        if (sbgnrm <= 0.0) {
            for (int i = 1; i <= N; ++i) Z(i) = X(i);
            return;
        }
        bool bnded = true;
        int nfree_c = N + 1, nbreak = 0, ibkmin = 0;
        double bkmin = 0.0;
        const int col2 = 2 * col;
        double f1 = 0.0;
        for (int i = 1; i <= col2; ++i) WA(OP + i) = 0.0;
        for (int i = 1; i <= N; ++i) {
            const double neggi = -G(i);
            double tl = 0.0, tu = 0.0;
            if (IWHERE(i) != 3 && IWHERE(i) != -1) {
                tl = X(i) - Lo(i);
                tu = Up(i) - X(i);
                const bool xlower = tl <= 0.0, xupper = tu <= 0.0;
                IWHERE(i) = 0;
                if (xlower) { if (neggi <= 0.0) IWHERE(i) = 1; }
                else if (xupper) { if (neggi >= 0.0) IWHERE(i) = 2; }
                else { if (fabs(neggi) <= 0.0) IWHERE(i) = -3; }
            }
            int pointr = head;
            if (IWHERE(i) != 0 && IWHERE(i) != -1) {
                D(i) = 0.0;
            } else {
                D(i) = neggi;
                f1 -= neggi * neggi;
                for (int j = 1; j <= col; ++j) {
                    WA(OP + j) += WY(i, pointr) * neggi;
                    WA(OP + col + j) += WS(i, pointr) * neggi;
                    pointr = pointr % M + 1;
                }
                if (neggi < 0.0) {
                    ++nbreak;
                    INDX2(nbreak) = i;
                    T(nbreak) = tl / (-neggi);
                    if (nbreak == 1 || T(nbreak) < bkmin) { bkmin = T(nbreak); ibkmin = nbreak; }
                } else if (neggi > 0.0) {
                    ++nbreak;
                    INDX2(nbreak) = i;
                    T(nbreak) = tu / neggi;
                    if (nbreak == 1 || T(nbreak) < bkmin) { bkmin = T(nbreak); ibkmin = nbreak; }
                } else {
                    --nfree_c;
                    INDX2(nfree_c) = i;
                    if (fabs(neggi) > 0.0) bnded = false;
                }
            }
        }
        if (theta != 1.0)
            for (int j = 1; j <= col; ++j) WA(OP + col + j) *= theta;
        for (int i = 1; i <= N; ++i) Z(i) = X(i);
        if (nbreak == 0 && nfree_c == N + 1) return;
        for (int j = 1; j <= col2; ++j) WA(OC + j) = 0.0;
        double f2 = -theta * f1;
        const double f2_org = f2;
        if (col > 0) {
            if (bmv(OP, OV) != 0) { info = 1; return; }
            double s = 0.0;
            for (int j = 1; j <= col2; ++j) s = fma(WA(OV + j), WA(OP + j), s);
            f2 -= s;
        }
        double dtm = -f1 / f2;
        double tsum = 0.0;
        nseg = 1;
        bool skip_to_end = false;
        if (nbreak != 0) {
            int nleft = nbreak, it = 1;
            double tj = 0.0;
            for (;;) {
                const double tj0 = tj;
                int ibp;
                if (it == 1) {
                    tj = bkmin;
                    ibp = INDX2(ibkmin);
                } else {
                    if (it == 2) {
                        if (ibkmin != nbreak) { T(ibkmin) = T(nbreak); INDX2(ibkmin) = INDX2(nbreak); }
                    }
                    hpsolb(nleft, it - 2);
                    tj = T(nleft);
                    ibp = INDX2(nleft);
                }
                const double dt = tj - tj0;
                if (dtm < dt) break;
                tsum += dt;
                --nleft;
                ++it;
                const double dibp = D(ibp);
                D(ibp) = 0.0;
                double zibp;
                if (dibp > 0.0) { zibp = Up(ibp) - X(ibp); Z(ibp) = Up(ibp); IWHERE(ibp) = 2; }
                else { zibp = Lo(ibp) - X(ibp); Z(ibp) = Lo(ibp); IWHERE(ibp) = 1; }
                if (nleft == 0 && nbreak == N) { dtm = dt; skip_to_end = true; break; }
                ++nseg;
                const double dibp2 = dibp * dibp;
                f1 = f1 + dt * f2 + dibp2 - theta * dibp * zibp;
                f2 = f2 - theta * dibp2;
                if (col > 0) {
                    for (int j = 1; j <= col2; ++j) WA(OC + j) = fma(dt, WA(OP + j), WA(OC + j));
                    int pointr = head;
                    for (int j = 1; j <= col; ++j) {
                        WA(OW + j) = WY(ibp, pointr);
                        WA(OW + col + j) = theta * WS(ibp, pointr);
                        pointr = pointr % M + 1;
                    }
                    if (bmv(OW, OV) != 0) { info = 1; return; }
                    double wmc = 0.0, wmp = 0.0, wmw = 0.0;
                    for (int j = 1; j <= col2; ++j) {
                        const double vj = WA(OV + j);
                        wmc = fma(WA(OC + j), vj, wmc);
                        wmp = fma(WA(OP + j), vj, wmp);
                        wmw = fma(WA(OW + j), vj, wmw);
                    }
                    for (int j = 1; j <= col2; ++j) WA(OP + j) = fma(-dibp, WA(OW + j), WA(OP + j));
                    f1 += dibp * wmc;
                    f2 += 2.0 * dibp * wmp - dibp2 * wmw;
                }
                f2 = fmax(EPSMCH * f2_org, f2);
                if (nleft > 0) {
                    dtm = -f1 / f2;
                    continue;
                } else if (bnded) {
                    f1 = 0.0; f2 = 0.0; dtm = 0.0;
                } else {
                    dtm = -f1 / f2;
                }
                break;
            }
        }
        if (!skip_to_end) {
            if (dtm <= 0.0) dtm = 0.0;
            tsum += dtm;
            for (int i = 1; i <= N; ++i) Z(i) = fma(tsum, D(i), Z(i));     // (scipy: BLAS daxpy, fused on current CPUs)
        }
        if (col > 0)
            for (int j = 1; j <= col2; ++j) WA(OC + j) = fma(dtm, WA(OP + j), WA(OC + j));
    }

    // ---- free / active variables at the Cauchy point ------------------------------------------------------------
    LB_HD void freev() {
        nenter = 0;
        ileave = N + 1;
        if (iter > 0) {
            for (int i = 1; i <= nfree; ++i) {
                const int k = INDEX(i);
                if (IWHERE(k) > 0) { --ileave; INDX2(ileave) = k; }
            }
            for (int i = 1 + nfree; i <= N; ++i) {
                const int k = INDEX(i);
                if (IWHERE(k) <= 0) { ++nenter; INDX2(nenter) = k; }
            }
        }
        wrk = (ileave < N + 1) || (nenter > 0) || updatd;
        nfree = 0;
        int iact = N + 1;
        for (int i = 1; i <= N; ++i) {
            if (IWHERE(i) <= 0) { ++nfree; INDEX(nfree) = i; }
            else { --iact; INDEX(iact) = i; }
        }
    }

    // ---- LEL' factorization of the indefinite 2col x 2col matrix of the subspace problem -----------------------------
    LB_HD void formk() {
        if (updatd) {
            if (iupdat > M) {
                for (int jy = 1; jy <= M - 1; ++jy) {
                    const int js = M + jy;
                    for (int q = 0; q < M - jy; ++q) WN1(jy + q, jy) = WN1(jy + 1 + q, jy + 1);
                    for (int q = 0; q < M - jy; ++q) WN1(js + q, js) = WN1(js + 1 + q, js + 1);
                    for (int q = 0; q < M - 1; ++q) WN1(M + 1 + q, jy) = WN1(M + 2 + q, jy + 1);
                }
            }
            const int pbegin = 1, pend = nfree, dbegin = nfree + 1, dend = N;
            int iy = col, is = M + col;
            int ipntr = head + col - 1;
            if (ipntr > M) ipntr -= M;
            int jpntr = head;
            for (int jy = 1; jy <= col; ++jy) {
                const int js = M + jy;
                double temp1 = 0.0, temp2 = 0.0, temp3 = 0.0;
                for (int k = pbegin; k <= pend; ++k) {
                    const int k1 = INDEX(k);
                    temp1 += WY(k1, ipntr) * WY(k1, jpntr);
                }
                for (int k = dbegin; k <= dend; ++k) {
                    const int k1 = INDEX(k);
                    temp2 += WS(k1, ipntr) * WS(k1, jpntr);
                    temp3 += WS(k1, ipntr) * WY(k1, jpntr);
                }
                WN1(iy, jy) = temp1;
                WN1(is, js) = temp2;
                WN1(is, jy) = temp3;
                jpntr = jpntr % M + 1;
            }
            const int jy = col;
            jpntr = head + col - 1;
            if (jpntr > M) jpntr -= M;
            ipntr = head;
            for (int i = 1; i <= col; ++i) {
                const int is2 = M + i;
                double temp3 = 0.0;
                for (int k = pbegin; k <= pend; ++k) {
                    const int k1 = INDEX(k);
                    temp3 += WS(k1, ipntr) * WY(k1, jpntr);
                }
                ipntr = ipntr % M + 1;
                WN1(is2, jy) = temp3;
            }
        }
        const int upcl = updatd ? col - 1 : col;
        if (nenter > 0 || ileave <= N) {     // (no variable entered or left the free set: every correction below is zero)
            int ipntr = head;
            for (int iy = 1; iy <= upcl; ++iy) {
                const int is = M + iy;
                int jpntr = head;
                for (int jy = 1; jy <= iy; ++jy) {
                    const int js = M + jy;
                    double temp1 = 0.0, temp2 = 0.0, temp3 = 0.0, temp4 = 0.0;
                    for (int k = 1; k <= nenter; ++k) {
                        const int k1 = INDX2(k);
                        temp1 += WY(k1, ipntr) * WY(k1, jpntr);
                        temp2 += WS(k1, ipntr) * WS(k1, jpntr);
                    }
                    for (int k = ileave; k <= N; ++k) {
                        const int k1 = INDX2(k);
                        temp3 += WY(k1, ipntr) * WY(k1, jpntr);
                        temp4 += WS(k1, ipntr) * WS(k1, jpntr);
                    }
                    WN1(iy, jy) = WN1(iy, jy) + temp1 - temp3;
                    WN1(is, js) = WN1(is, js) - temp2 + temp4;
                    jpntr = jpntr % M + 1;
                }
                ipntr = ipntr % M + 1;
            }
            ipntr = head;
            for (int is = M + 1; is <= M + upcl; ++is) {
                int jpntr = head;
                for (int jy = 1; jy <= upcl; ++jy) {
                    double temp1 = 0.0, temp3 = 0.0;
                    for (int k = 1; k <= nenter; ++k) {
                        const int k1 = INDX2(k);
                        temp1 += WS(k1, ipntr) * WY(k1, jpntr);
                    }
                    for (int k = ileave; k <= N; ++k) {
                        const int k1 = INDX2(k);
                        temp3 += WS(k1, ipntr) * WY(k1, jpntr);
                    }
                    if (is <= jy + M) WN1(is, jy) = WN1(is, jy) + temp1 - temp3;
                    else WN1(is, jy) = WN1(is, jy) - temp1 + temp3;
                    jpntr = jpntr % M + 1;
                }
                ipntr = ipntr % M + 1;
            }
        }
        for (int iy = 1; iy <= col; ++iy) {
            const int is = col + iy, is1 = M + iy;
            for (int jy = 1; jy <= iy; ++jy) {
                const int js = col + jy, js1 = M + jy;
                WN(jy, iy) = WN1(iy, jy) / theta;
                WN(js, is) = WN1(is1, js1) * theta;
            }
            for (int jy = 1; jy <= iy - 1; ++jy) WN(jy, is) = -WN1(is1, jy);
            for (int jy = iy; jy <= col; ++jy) WN(jy, is) = WN1(is1, jy);
            WN(iy, iy) += SY(iy, iy);
        }
        auto w11 = [&](int i, int j) -> double& { return WN(i, j); };
        if (chol_upper(w11, col) != 0) { info = -1; return; }
        const int col2 = 2 * col;
        for (int js = col + 1; js <= col2; ++js) {
            auto bcol = [&](int i) -> double& { return WN(i, js); };
            if (tri_solve(w11, col, bcol, true) != 0) { info = -1; return; }   // cannot happen after a successful factorization
        }
        for (int is = col + 1; is <= col2; ++is)
            for (int js = is; js <= col2; ++js) {
                double s = 0.0;
                for (int q = 1; q <= col; ++q) s = fma(WN(q, is), WN(q, js), s);
                WN(is, js) += s;
            }
        const int c0 = col;
        auto w22 = [&, c0](int i, int j) -> double& { return WN(c0 + i, c0 + j); };
        if (chol_upper(w22, col) != 0) { info = -2; return; }
    }

    // ---- r = -Z'(B(xcp - x) + g) ----------------------------------------------------------------------------------
    LB_HD void cmprlb() {
        const int OP = 0, OC = 2 * M;
        for (int i = 1; i <= nfree; ++i) {
            const int k = INDEX(i);
            R(i) = -theta * (Z(k) - X(k)) - G(k);
        }
        if (bmv(OC, OP) != 0) { info = -8; return; }
        int pointr = head;
        for (int j = 1; j <= col; ++j) {
            const double a1 = WA(OP + j), a2 = theta * WA(OP + col + j);
            for (int i = 1; i <= nfree; ++i) {
                const int k = INDEX(i);
                R(i) += WY(k, pointr) * a1 + WS(k, pointr) * a2;
            }
            pointr = pointr % M + 1;
        }
    }

    // ---- subspace minimization: direct primal method + projection (version 3.0) ---------------------------------------
    LB_HD void subsm() {
        const int nsub = nfree;
        if (nsub <= 0) return;
        int pointr = head;
        for (int i = 1; i <= col; ++i) {
            double temp1 = 0.0, temp2 = 0.0;
            for (int j = 1; j <= nsub; ++j) {
                const int k = INDEX(j);
                temp1 += WY(k, pointr) * R(j);
                temp2 += WS(k, pointr) * R(j);
            }
            WA(i) = temp1;
            WA(col + i) = theta * temp2;
            pointr = pointr % M + 1;
        }
        const int col2 = 2 * col;
        auto wn = [&](int i, int j) -> double& { return WN(i, j); };
        auto wv = [&](int i) -> double& { return WA(i); };
        if (tri_solve(wn, col2, wv, true) != 0) { info = 1; return; }
        for (int i = 1; i <= col; ++i) WA(i) = -WA(i);
        if (tri_solve(wn, col2, wv, false) != 0) { info = 1; return; }
        pointr = head;
        for (int jy = 1; jy <= col; ++jy) {
            const int js = col + jy;
            for (int i = 1; i <= nsub; ++i) {
                const int k = INDEX(i);
                R(i) += WY(k, pointr) * WA(jy) / theta + WS(k, pointr) * WA(js);
            }
            pointr = pointr % M + 1;
        }
        for (int i = 1; i <= nsub; ++i) R(i) *= (1.0 / theta);
        // projection of the Newton step
        iword = 0;
        for (int i = 1; i <= N; ++i) XP(i) = Z(i);
        for (int i = 1; i <= nsub; ++i) {
            const int k = INDEX(i);
            const double dk = R(i);
            double xk = Z(k);
            xk = fmax(Lo(k), xk + dk);
            Z(k) = fmin(Up(k), xk);
            if (Z(k) == Lo(k) || Z(k) == Up(k)) iword = 1;
        }
        if (iword == 0) return;
        double dd_p = 0.0;
        for (int i = 1; i <= N; ++i) dd_p += (Z(i) - X(i)) * G(i);
        if (dd_p > 0.0) {
            for (int i = 1; i <= N; ++i) Z(i) = XP(i);
            double alpha = 1.0, temp1 = alpha;
            int ibd = 0;
            for (int i = 1; i <= nsub; ++i) {
                const int k = INDEX(i);
                const double dk = R(i);
                if (dk < 0.0) {
                    const double temp2 = Lo(k) - Z(k);
                    if (temp2 >= 0.0) temp1 = 0.0;
                    else if (dk * alpha < temp2) temp1 = temp2 / dk;
                } else if (dk > 0.0) {
                    const double temp2 = Up(k) - Z(k);
                    if (temp2 <= 0.0) temp1 = 0.0;
                    else if (dk * alpha > temp2) temp1 = temp2 / dk;
                }
                if (temp1 < alpha) { alpha = temp1; ibd = i; }
            }
            if (alpha < 1.0) {
                const double dk = R(ibd);
                const int k = INDEX(ibd);
                if (dk > 0.0) { Z(k) = Up(k); R(ibd) = 0.0; }
                else if (dk < 0.0) { Z(k) = Lo(k); R(ibd) = 0.0; }
            }
            for (int i = 1; i <= nsub; ++i) {
                const int k = INDEX(i);
                Z(k) += alpha * R(i);
            }
        }
    }

    // ---- safeguarded cubic/quadratic step of the More-Thuente line search ----------------------------------------------
    LB_HD static double max3(double p, double q, double r) { return fmax(fmax(p, q), r); }
    LB_HD void dcstep(double& stx_, double& fx_, double& dx, double& sty_, double& fy_, double& dy, double& stp_,
                      const double fp, const double dp, const double stpmin, const double stpmax) {
        const double sgnd = dp * (dx / fabs(dx));
        double stpf, stpc, stpq, th, s, gamma, p, q, r;
        if (fp > fx_) {
            th = 3.0 * (fx_ - fp) / (stp_ - stx_) + dx + dp;
            s = max3(fabs(th), fabs(dx), fabs(dp));
            gamma = s * sqrt((th / s) * (th / s) - (dx / s) * (dp / s));
            if (stp_ < stx_) gamma = -gamma;
            p = (gamma - dx) + th;
            q = ((gamma - dx) + gamma) + dp;
            r = p / q;
            stpc = stx_ + r * (stp_ - stx_);
            stpq = stx_ + ((dx / ((fx_ - fp) / (stp_ - stx_) + dx)) / 2.0) * (stp_ - stx_);
            if (fabs(stpc - stx_) < fabs(stpq - stx_)) stpf = stpc;
            else stpf = stpc + (stpq - stpc) / 2.0;
            brackt = 1;
        } else if (sgnd < 0.0) {
            th = 3.0 * (fx_ - fp) / (stp_ - stx_) + dx + dp;
            s = max3(fabs(th), fabs(dx), fabs(dp));
            gamma = s * sqrt((th / s) * (th / s) - (dx / s) * (dp / s));
            if (stp_ > stx_) gamma = -gamma;
            p = (gamma - dp) + th;
            q = ((gamma - dp) + gamma) + dx;
            r = p / q;
            stpc = stp_ + r * (stx_ - stp_);
            stpq = stp_ + (dp / (dp - dx)) * (stx_ - stp_);
            if (fabs(stpc - stp_) > fabs(stpq - stp_)) stpf = stpc;
            else stpf = stpq;
            brackt = 1;
        } else if (fabs(dp) < fabs(dx)) {
            th = 3.0 * (fx_ - fp) / (stp_ - stx_) + dx + dp;
            s = max3(fabs(th), fabs(dx), fabs(dp));
            gamma = s * sqrt(fmax(0.0, (th / s) * (th / s) - (dx / s) * (dp / s)));
            if (stp_ > stx_) gamma = -gamma;
            p = (gamma - dp) + th;
            q = (gamma + (dx - dp)) + gamma;
            r = p / q;
            if (r < 0.0 && gamma != 0.0) stpc = stp_ + r * (stx_ - stp_);
            else if (stp_ > stx_) stpc = stpmax;
            else stpc = stpmin;
            stpq = stp_ + (dp / (dp - dx)) * (stx_ - stp_);
            if (brackt) {
                if (fabs(stpc - stp_) < fabs(stpq - stp_)) stpf = stpc;
                else stpf = stpq;
                if (stp_ > stx_) stpf = fmin(stp_ + 0.66 * (sty_ - stp_), stpf);
                else stpf = fmax(stp_ + 0.66 * (sty_ - stp_), stpf);
            } else {
                if (fabs(stpc - stp_) > fabs(stpq - stp_)) stpf = stpc;
                else stpf = stpq;
                stpf = fmin(stpmax, stpf);
                stpf = fmax(stpmin, stpf);
            }
        } else {
            if (brackt) {
                th = 3.0 * (fp - fy_) / (sty_ - stp_) + dy + dp;
                s = max3(fabs(th), fabs(dy), fabs(dp));
                gamma = s * sqrt((th / s) * (th / s) - (dy / s) * (dp / s));
                if (stp_ > sty_) gamma = -gamma;
                p = (gamma - dp) + th;
                q = ((gamma - dp) + gamma) + dy;
                r = p / q;
                stpc = stp_ + r * (sty_ - stp_);
                stpf = stpc;
            } else if (stp_ > stx_) stpf = stpmax;
            else stpf = stpmin;
        }
        if (fp > fx_) {
            sty_ = stp_; fy_ = fp; dy = dp;
        } else {
            if (sgnd < 0.0) { sty_ = stx_; fy_ = fx_; dy = dx; }
            stx_ = stp_; fx_ = fp; dx = dp;
        }
        stp_ = stpf;
    }

    // ---- More-Thuente line search (MINPACK-2 dcsrch), ftol = 1e-3, gtol = 0.9, xtol = 0.1, stpmin = 0 ------------------
    LB_HD void dcsrch(const double f, const double g, const double stpmax_) {
        const double ftol = 1e-3, gtol = 0.9, xtol = 0.1, stpmin_ = 0.0, xtrapl = 1.1, xtrapu = 4.0;
        if (ls == LS_START) {
            // (argument errors - stp outside [stpmin, stpmax], g >= 0 - cannot be raised here: the caller has checked)
            brackt = 0;
            stage = 1;
            finit = f;
            ginit = g;
            gtest = ftol * ginit;
            width = stpmax_ - stpmin_;
            width1 = width / 0.5;
            stx = 0.0; fx = finit; gx = ginit;
            sty = 0.0; fy = finit; gy = ginit;
            stmin = 0.0;
            stmax = stp + xtrapu * stp;
            ls = LS_FG;
            return;
        }
        const double ftest = finit + stp * gtest;
        if (stage == 1 && f <= ftest && g >= 0.0) stage = 2;
        int res = LS_FG;
        if (brackt && (stp <= stmin || stp >= stmax)) res = LS_WARN;
        if (brackt && stmax - stmin <= xtol * stmax) res = LS_WARN;
        if (stp == stpmax_ && f <= ftest && g <= gtest) res = LS_WARN;
        if (stp == stpmin_ && (f > ftest || g >= gtest)) res = LS_WARN;
        if (f <= ftest && fabs(g) <= gtol * (-ginit)) res = LS_CONV;
        if (res != LS_FG) { ls = res; return; }
        if (stage == 1 && f <= fx && f > ftest) {
            const double fm = f - stp * gtest;
            double fxm = fx - stx * gtest, fym = fy - sty * gtest;
            const double gm = g - gtest;
            double gxm = gx - gtest, gym = gy - gtest;
            dcstep(stx, fxm, gxm, sty, fym, gym, stp, fm, gm, stmin, stmax);
            fx = fxm + stx * gtest;
            fy = fym + sty * gtest;
            gx = gxm + gtest;
            gy = gym + gtest;
        } else {
            dcstep(stx, fx, gx, sty, fy, gy, stp, f, g, stmin, stmax);
        }
        if (brackt) {
            if (fabs(sty - stx) >= 0.66 * width1) stp = stx + 0.5 * (sty - stx);
            width1 = width;
            width = fabs(sty - stx);
        }
        if (brackt) {
            stmin = fmin(stx, sty);
            stmax = fmax(stx, sty);
        } else {
            stmin = stp + xtrapl * (stp - stx);
            stmax = stp + xtrapu * (stp - stx);
        }
        stp = fmax(stp, stpmin_);
        stp = fmin(stp, stpmax_);
        if ((brackt && (stp <= stmin || stp >= stmax)) || (brackt && stmax - stmin <= xtol * stmax)) stp = stx;
        ls = LS_FG;
    }

    // ---- line search driver: returns with task = T_FG_LN (evaluate at X) or T_NEW_X -------------------------------------
    LB_HD void lnsrlb() {
        if (task != T_FG_LN) {
            dtd = 0.0;
            for (int i = 1; i <= N; ++i) dtd = fma(D(i), D(i), dtd);
            dnorm = sqrt(dtd);
            stpmx = 1e10;
            if (iter == 0) {
                stpmx = 1.0;
            } else {
                for (int i = 1; i <= N; ++i) {
                    const double a1 = D(i);
                    if (a1 < 0.0) {
                        const double a2 = Lo(i) - X(i);
                        if (a2 >= 0.0) stpmx = 0.0;
                        else if (a1 * stpmx < a2) stpmx = a2 / a1;
                    } else if (a1 > 0.0) {
                        const double a2 = Up(i) - X(i);
                        if (a2 <= 0.0) stpmx = 0.0;
                        else if (a1 * stpmx > a2) stpmx = a2 / a1;
                    }
                }
            }
            stp = 1.0;          // every variable is bounded ("boxed"), so the first step is not scaled by 1/|d|
            for (int i = 1; i <= N; ++i) { T(i) = X(i); R(i) = G(i); }
            fold = F();
            ifun = 0;
            iback = 0;
            ls = LS_START;
        }
        gd = 0.0;
        for (int i = 1; i <= N; ++i) gd = fma(G(i), D(i), gd);
        if (ifun == 0) {
            gdold = gd;
            if (gd >= 0.0) { info = -4; return; }     // not a descent direction
        }
        dcsrch(F(), gd, stpmx);
        if (ls != LS_CONV && ls != LS_WARN) {
            task = T_FG_LN;
            ++ifun;
            ++nfgv;
            iback = ifun - 1;
            if (stp == 1.0) {
                for (int i = 1; i <= N; ++i) X(i) = Z(i);
            } else {
                for (int i = 1; i <= N; ++i) X(i) = stp * D(i) + T(i);
            }
        } else {
            task = T_NEW_X;
        }
    }

    // ---- limited-memory update of WS, WY, SY, SS ----------------------------------------------------------------------
    LB_HD void matupd(const double rr, const double dr) {
        if (iupdat <= M) {
            col = iupdat;
            itail = (head + iupdat - 2) % M + 1;
        } else {
            itail = itail % M + 1;
            head = head % M + 1;
        }
        for (int i = 1; i <= N; ++i) { WS(i, itail) = D(i); WY(i, itail) = R(i); }
        theta = rr / dr;
        if (iupdat > M) {
            for (int j = 1; j <= col - 1; ++j) {
                for (int q = 0; q < j; ++q) SS(1 + q, j) = SS(2 + q, j + 1);
                for (int q = 0; q < col - j; ++q) SY(j + q, j) = SY(j + 1 + q, j + 1);
            }
        }
        int pointr = head;
        for (int j = 1; j <= col - 1; ++j) {
            double s1 = 0.0, s2 = 0.0;
            for (int i = 1; i <= N; ++i) { s1 = fma(D(i), WY(i, pointr), s1); s2 = fma(WS(i, pointr), D(i), s2); }
            SY(col, j) = s1;
            SS(j, col) = s2;
            pointr = pointr % M + 1;
        }
        if (stp == 1.0) SS(col, col) = dtd;
        else SS(col, col) = stp * stp * dtd;
        SY(col, col) = dr;
    }

    // ---- T = theta SS + L D^-1 L', Cholesky factor in the upper triangle of WT ----------------------------------------
    LB_HD void formt() {
        for (int j = 1; j <= col; ++j) WT(1, j) = theta * SS(1, j);
        for (int i = 2; i <= col; ++i)
            for (int j = i; j <= col; ++j) {
                const int k1 = (i < j ? i : j) - 1;
                double ddum = 0.0;
                for (int k = 1; k <= k1; ++k) ddum += SY(i, k) * SY(j, k) / SY(k, k);
                WT(i, j) = ddum + theta * SS(i, j);
            }
        auto wt = [&](int i, int j) -> double& { return WT(i, j); };
        if (chol_upper(wt, col) != 0) info = -3;
    }

    // ---- one call of the reverse-communication routine ------------------------------------------------------------------
    LB_HD void setulb() {
        if (task == T_START) {
            // x has been projected onto the box by init(); every variable is constrained and none is fixed
            for (int i = 1; i <= N; ++i) IWHERE(i) = (Up(i) - Lo(i) <= 0.0) ? 3 : 0;
            task = T_FG_START;
            return;
        }
        if (task == T_FG_START) {
            nfgv = 1;
            projgr();
            if (sbgnrm <= PGTOL) { task = T_CONV; return; }
        } else if (task == T_FG_LN) {
            goto line_search;
        } else if (task == T_NEW_X) {
            goto new_x;
        } else {
            return;
        }

    iteration:
        iword = -1;
        cauchy();
        if (info != 0) { refresh(); goto iteration; }
        freev();
        nact = N - nfree;
        if (nfree != 0 && col != 0) {
            if (wrk) formk();
            if (info != 0) { refresh(); goto iteration; }
            cmprlb();
            if (info == 0) subsm();
            if (info != 0) { refresh(); goto iteration; }
        }
        for (int i = 1; i <= N; ++i) D(i) = Z(i) - X(i);

    line_search:
        lnsrlb();
        if (info != 0 || iback >= MAXLS) {
            for (int i = 1; i <= N; ++i) { X(i) = T(i); G(i) = R(i); }
            F() = fold;
            if (col == 0) {
                if (info == 0) { info = -9; --nfgv; --ifun; --iback; }
                task = T_ABNORMAL;
                ++iter;
                return;
            }
            if (info == 0) --nfgv;
            refresh();
            task = T_FG_START;       // (value irrelevant below: the line search starts afresh because task != T_FG_LN)
            goto iteration;
        } else if (task == T_FG_LN) {
            return;
        } else {
            ++iter;
            projgr();
            return;                  // task == T_NEW_X
        }

    new_x:
        if (sbgnrm <= PGTOL) { task = T_CONV; return; }
        {
            const double ddum = max3(fabs(fold), fabs(F()), 1.0);
            if ((fold - F()) <= FTOL * ddum) {
                task = T_CONV;
                if (iback >= 10) info = -5;
                return;
            }
        }
        {
            double rr = 0.0, dr, ddum;
            for (int i = 1; i <= N; ++i) { R(i) = G(i) - R(i); }
            for (int i = 1; i <= N; ++i) rr = fma(R(i), R(i), rr);
            if (stp == 1.0) {
                dr = gd - gdold;
                ddum = -gdold;
            } else {
                dr = (gd - gdold) * stp;
                for (int i = 1; i <= N; ++i) D(i) *= stp;
                ddum = -gdold * stp;
            }
            if (dr <= EPSMCH * ddum) {
                ++nskip;
                updatd = 0;
            } else {
                updatd = 1;
                ++iupdat;
                matupd(rr, dr);
                formt();
                if (info != 0) refresh();
            }
        }
        goto iteration;
    }

    // ---- scipy's wrapper loop (_minimize_lbfgsb): run until the objective is needed (true) or the run has ended -------
    // Before the call (except the first one) the caller has written f and g at the point X.
    LB_HD bool advance() {
        for (;;) {
            setulb();
            if (task == T_FG_START || task == T_FG_LN) return true;
            if (task == T_NEW_X) {
                ++nit;
                if (nit >= MAXITER || nfev > MAXFUN) { task = T_STOP; return false; }
                continue;
            }
            return false;
        }
    }
};

}  // namespace gpet_lb
