"""The L-BFGS-B state machine of the final fit (csrc/gpet_lbfgsb.cuh) against scipy's own reverse-communication
routine, on the CPU: the host twins gpet_lbfgsb_host_init / _host_advance run the code the device kernels run.

Reference seam: sklearn_gpr.py:587-607 calls scipy.optimize.minimize(method='L-BFGS-B', jac=True, bounds=...); what
has to be reproduced is scipy's sequence of evaluation points for given objective values. Both sides are fed the SAME
(f, g) (evaluated at scipy's point), so the comparison isolates the algorithm: every request flag must agree and the
requested points must agree to rounding. A one-ulp difference can decide whether a step that runs to a bound lands on
it (after which the paths separate - true of scipy against itself across BLAS builds too), so a small fraction of
runs may leave the rounding-level envelope; their number is bounded and they must still terminate normally."""
import ctypes
import os
import sys

import numpy as np
import pytest
import scipy.optimize

import gpet_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gaussian_process_edge_trace_b200 import _cabi, _lbfgs_worker as W   # noqa: E402

LOW, UP = O.FINAL_FIT_BOUNDS[:, 0].copy(), O.FINAL_FIT_BOUNDS[:, 1].copy()


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_cabi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _cabi.load()


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class Ours:
    def __init__(self, lib, x0, fill=0.0):
        self.lib, self.E = lib, x0.shape[0]
        nd, ni = lib.gpet_lbfgsb_state_doubles(), lib.gpet_lbfgsb_state_ints()
        self.ds, self.ist = np.full((self.E, nd), fill), np.full((self.E, ni), int(fill != 0) * 77, dtype=np.int32)
        self.need, self.x = np.zeros(self.E, dtype=np.int32), np.zeros((self.E, 3))
        x0 = np.ascontiguousarray(x0)
        assert lib.gpet_lbfgsb_host_init(P(self.ds), P(self.ist), self.E, P(x0), P(LOW), P(UP)) == 0
        self.step(np.zeros(self.E, dtype=np.int32), np.zeros(self.E), np.zeros((self.E, 3)))

    def step(self, give, f, g):
        assert self.lib.gpet_lbfgsb_host_advance(P(self.ds), P(self.ist), self.E, P(give), P(f), P(g), P(self.need),
                                                 P(self.x)) == 0


def synthetic(seed, amp_max):
    r = np.random.RandomState(seed)
    Q = r.randn(3, 3)
    A = Q @ Q.T + 0.3 * np.eye(3)
    c = r.uniform(LOW - 5, UP + 5)          # the unconstrained optimum is often outside the box
    amp, sc = r.uniform(0, amp_max), np.exp(r.uniform(-3, 1, 3))

    def fg(x):
        d = (x - c) * sc
        q = A @ d
        return 0.5 * d @ q + amp * np.sum(np.cos(1.3 * x)), sc * q - 1.3 * amp * np.sin(1.3 * x)
    return fg


@pytest.mark.parametrize("amp_max", [2.0, 30.0])
def test_lockstep_against_scipy_setulb(lib, amp_max):
    E = 240
    rng = np.random.RandomState(int(amp_max))
    x0 = rng.uniform(LOW, UP, size=(E, 3))
    x0[:8] = np.clip(x0[:8] * 3, LOW, UP)            # some starts on the bounds
    fgs = [synthetic(1000 * int(amp_max) + e, amp_max) for e in range(E)]
    ours = Ours(lib, x0)
    ref = [W.Instance(x0[e], LOW, UP) for e in range(E)]
    ref_need = np.array([s.advance() for s in ref])
    separated, maxdev, rounds = set(), 0.0, 0
    give, f, g = np.zeros(E, dtype=np.int32), np.zeros(E), np.zeros((E, 3))
    while True:
        for e in range(E):
            if e in separated:
                continue
            assert bool(ours.need[e]) == bool(ref_need[e]) or e in separated, (e, rounds)
            if ref_need[e]:
                dev = np.abs(ours.x[e] - ref[e].x).max()
                if dev > 1e-8:
                    separated.add(e)
                else:
                    maxdev = max(maxdev, dev)
        act = [e for e in range(E) if ref_need[e] and e not in separated]
        if not act:
            break
        give[:] = 0
        for e in act:
            f[e], g[e] = fgs[e](ref[e].x.copy())
            give[e] = 1
            ref[e].give(f[e], g[e])
            ref_need[e] = ref[e].advance()
        for e in separated:
            ref_need[e] = False
        # separated runs keep going on their own objective values so that they can be checked at the end
        for e in separated:
            if ours.need[e]:
                f[e], g[e] = fgs[e](ours.x[e].copy())
                give[e] = 1
        ours.step(give, f, g)
        rounds += 1
        assert rounds < 3000
    while any(ours.need[e] for e in separated):
        give[:] = 0
        for e in separated:
            if ours.need[e]:
                f[e], g[e] = fgs[e](ours.x[e].copy())
                give[e] = 1
        ours.step(give, f, g)
    assert maxdev < 1e-8
    assert len(separated) <= 3, sorted(separated)
    task = ours.ist[:, 9 + 20]              # I_SC + 20 (gpet_lbfgsb.cuh): 4 converged, 5 abnormal line search, 6 limit
    for e in range(E):                      # same end point as scipy; a separated run must still have terminated normally
        if e not in separated:
            assert np.abs(ours.ds[e, 0:3] - ref[e].x).max() < 1e-8
            assert abs(ours.ds[e, 3] - float(ref[e].f)) <= 1e-12 * max(1.0, abs(float(ref[e].f)))
        assert task[e] in (4, 5), (e, task[e])


def test_final_fit_objective_best_of_13(lib):
    """The reference's use: 13 starts per trace on -(log marginal likelihood) (oracle restatement of
    sklearn_gpr.py:512-583), best run kept. Training sets: the README trace's final observations, thinned/perturbed."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "trace_cfg1.npz"))
    obs, init = g["final_obs"].reshape(-1, 2), g["init"]
    worst_best, bad_runs = 0.0, 0
    for t in range(4):
        r = np.random.RandomState(100 + t)
        keep = r.rand(len(obs)) < r.uniform(0.5, 1.0)
        pts = np.concatenate([init, obs[keep]])
        pts = pts[np.argsort(pts[:, 0], kind="stable")]
        X = pts[:, 0].astype(float)
        y = pts[:, 1].astype(float) + r.randn(len(pts)) * r.uniform(0, 3)
        w = np.ones(len(X))
        w[0] = w[-1] = 1e-7
        y = (y - y.mean()) / y.std()
        X = (X - X.mean()) / X.std()
        yt = (y - y.mean()) / y.std()

        def fun(th):
            return O.neg_lml_and_grad(th, X, yt, w, "RBF", None, 1e-6)
        rs = np.random.RandomState(7 + t)
        x0s = np.array([np.log(np.array([5.0, 5.0, 1.0]))] + [rs.uniform(LOW, UP) for _ in range(12)])
        ref = [scipy.optimize.minimize(fun, x0, method="L-BFGS-B", jac=True, bounds=O.FINAL_FIT_BOUNDS) for x0 in x0s]
        ours = Ours(lib, x0s)
        f, gg = np.zeros(13), np.zeros((13, 3))
        while ours.need.any():
            give = ours.need.copy()
            for e in np.nonzero(give)[0]:
                f[e], gg[e] = fun(ours.x[e].copy())
            ours.step(give, f, gg)
        xs, fs = ours.ds[:, 0:3], ours.ds[:, 3]
        rx, rf = np.array([q.x for q in ref]), np.array([q.fun for q in ref])
        bad_runs += int((np.abs(xs - rx).max(axis=1) > 1e-5).sum())
        worst_best = max(worst_best, np.abs(xs[np.argmin(fs)] - rx[np.argmin(rf)]).max())
    assert worst_best < 1e-5          # the bar the GPU parity tests put on the optimised theta
    assert bad_runs <= 4


def test_state_needs_no_initialisation_by_the_caller(lib):
    """init zeroes the workspace itself (scipy's iterates depend on never-written entries being zero): runs on
    caller memory full of large values reproduce the runs on zeroed memory bit for bit."""
    E = 300
    rng = np.random.RandomState(5)
    x0 = rng.uniform(LOW, UP, size=(E, 3))
    fgs = [synthetic(50 + e, 2.0) for e in range(E)]
    hist = []
    for fill in (0.0, 1e300):
        ours = Ours(lib, x0, fill)
        f, g, pts = np.zeros(E), np.zeros((E, 3)), []
        while ours.need.any():
            give = ours.need.copy()
            for e in np.nonzero(give)[0]:
                f[e], g[e] = fgs[e](ours.x[e].copy())
            pts.append(ours.x[give.astype(bool)].copy())
            ours.step(give, f, g)
        hist.append(pts)
    assert len(hist[0]) == len(hist[1])
    assert all(np.array_equal(a, b) for a, b in zip(*hist))
