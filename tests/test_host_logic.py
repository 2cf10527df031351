"""CPU tests of the host-side logic of the package (no CUDA needed): the pieces of the hot path that stay on the host
are checked against the oracle's restatement of the same reference lines, and the pipelined driver's scheduling is
exercised with stand-in sub-batches."""
import sys
import threading

import numpy as np
import pytest

import gpet_oracle as O
from conftest import ROOT

sys.path.insert(0, ROOT)
from gaussian_process_edge_trace_b200 import _gp_host  # noqa: E402


def test_threshold_loop_batch_matches_oracle():
    """Vectorised decay loop (gpet.py:589-609) == the oracle's scalar loop, trace by trace, incl. thresholds."""
    rng = np.random.default_rng(3)
    B, nb = 40, 37
    best = rng.uniform(0.0, 1.0, size=(B, nb))
    best[rng.uniform(size=(B, nb)) < 0.3] = -1.0                  # empty bins
    n_pre = rng.integers(0, 10, size=B)
    thr = rng.uniform(0.2, 1.0, size=B)
    active = rng.uniform(size=B) < 0.8
    thr_in = thr.copy()
    mask = _gp_host.threshold_loop_batch(best, n_pre.copy(), 5, 30, thr, active)
    for b in range(B):
        if not active[b]:
            assert thr[b] == thr_in[b] and not mask[b].any()
            continue
        m_o, t_o = O.threshold_loop(np.where(best[b] >= 0, best[b], -np.inf), int(n_pre[b]), 5, 30, float(thr_in[b]))
        assert np.array_equal(mask[b], m_o) and thr[b] == t_o


def test_threshold_loop_guard():
    best = np.full((1, 4), -1.0)
    with pytest.raises(RuntimeError):
        _gp_host.threshold_loop_batch(best, np.zeros(1, dtype=np.int64), 5, 30, np.ones(1), np.ones(1, dtype=bool))


@pytest.mark.parametrize("N,x_st,x_en,dx,fix", [(500, 0, 499, 5, True), (160, 3, 150, 8, True), (97, 0, 96, 2, False),
                                                 (1024, 10, 1000, 20, True)])
def test_column_bins(N, x_st, x_en, dx, fix):
    """Bins = np.round((x - x_st)/delta_x) (gpet.py:605-606); candidate columns per fix_endpoints (gpet.py:652-655);
    groups never split a bin and hold at most 64 columns."""
    col_bin, groups, nb, lo = _gp_host.column_bins(N, x_st, x_en, dx, fix)
    x = np.arange(N)
    want = np.round((x - x_st) / dx).astype(int)
    got = np.where(col_bin >= 0, col_bin, -(col_bin + 1)) + lo
    assert np.array_equal(got, want) and nb == want.max() - want.min() + 1
    cand = (x > x_st) & (x < x_en) if fix else np.ones(N, dtype=bool)
    assert np.array_equal(col_bin >= 0, cand)
    assert groups[0] == 0 and groups[-1] == N and np.all(np.diff(groups) > 0) and np.all(np.diff(groups) <= 64)
    for g in groups[1:-1]:
        assert want[g] != want[g - 1]


def test_training_set_assembly_matches_oracle():
    rng = np.random.default_rng(5)
    init = np.array([[0, 40], [99, 43]])
    obs = np.stack([rng.permutation(np.arange(1, 99))[:17], rng.integers(0, 80, size=17)], axis=1)
    alpha = np.array([1e-7, 1e-7])
    x, y, w = _gp_host.assemble_training_set(init, obs, alpha)
    X, yo, wo = O.assemble_training_set(init, obs, alpha)
    assert np.array_equal(x, X.astype(np.int64)) and np.array_equal(y, yo) and np.array_equal(w, wo)


@pytest.mark.parametrize("kernel", [("RBF", 2.5, 20.0), ("Matern", 2.5, 20.0), ("Matern", 1.5, 9.0)])
def test_grid_eigenbasis_reconstructs_kernel(kernel):
    """Low-rank basis of the unit kernel matrix on the pixel grid: U_r diag(lam) U_r^T == k(x_grid, x_grid) to the
    rank tolerance; rp is a multiple of 4; kd is the kernel by integer distance (the oracle's unit_kernel)."""
    ktype, nu, ls = kernel
    xg = np.arange(10, 210)
    kd, Ur, lam, r = _gp_host.grid_eigenbasis(ktype, nu, ls, xg, 128)
    Kref = O.unit_kernel(ktype, nu, ls, xg.astype(float))
    assert np.abs(kd - Kref[0]).max() < 1e-15
    if Ur is None:
        assert r > 128
        return
    assert Ur.shape[1] % 4 == 0 and Ur.shape[1] >= r
    assert np.abs((Ur * lam[None, :]) @ Ur.T - Kref).max() < 1e-12


def test_normal_draws_shared_and_exact():
    from gaussian_process_edge_trace_b200.engine import NormalDraws
    a = NormalDraws.shared(64, 30, 8, 11)
    b = NormalDraws.shared(64, 30, 8, 11)
    assert a is b and NormalDraws.shared(64, 30, 8, 12) is not a
    out = {}

    def worker(name, its):
        out[name] = [a.get(i).copy() for i in its]

    ts = [threading.Thread(target=worker, args=("x", [0, 1, 2, 3, 5])), threading.Thread(target=worker, args=("y", [1, 0, 4, 2]))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for name, its in (("x", [0, 1, 2, 3, 5]), ("y", [1, 0, 4, 2])):
        for z, it in zip(out[name], its):
            ref = np.random.RandomState(11 + it + 1).standard_normal((64, 30))[:, :8].T      # gpet.py:839
            assert np.array_equal(z, ref)


class _FakeBatch:
    """Stand-in for TraceBatch in the scheduling test: converges after `iters` iterations."""
    final_fit_mode = "device"

    def __init__(self, name, iters, log):
        self.name, self.left, self.log, self.B = name, iters, log, 1
        self.launched = False

    def use_own_stream(self):
        self.own_stream = True

    def release_loop_buffers(self):
        assert not self.launched and self.left == 0
        self.log.append(("release", self.name))

    def step_launch(self):
        assert not self.launched
        if self.left == 0:
            return False
        self.launched = True
        self.log.append(("launch", self.name))
        return True

    def step_finish(self):
        assert self.launched
        self.launched = False
        self.left -= 1
        self.log.append(("finish", self.name))


def test_pipelined_schedule(monkeypatch):
    """Window of two alternating sub-batches, merged fits of the ones that converged together, every sub-batch fitted
    exactly once, results in batch order; factories are only built when they enter the window."""
    torch = pytest.importorskip("torch")
    from gaussian_process_edge_trace_b200 import engine
    log, fits, built = [], [], []

    def fake_fit_group(tbs):
        fits.append([tb.name for tb in tbs])
        return [(np.full((1, 2, 2), ord(tb.name)), [(tb.name, tb.name)], {"rounds": 1}) for tb in tbs]

    class _Exec:
        def submit(self, fn, *a):
            fn(*a)
            f = type("F", (), {"result": lambda self: None, "done": lambda self: True})()
            return f

    class _Ctx:
        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    monkeypatch.setattr(engine, "final_fit_group", fake_fit_group)
    monkeypatch.setattr(engine, "_fit_resources", lambda: (_Exec(), None))
    monkeypatch.setattr(torch.cuda, "stream", lambda s: _Ctx())

    def factory(name, iters):
        def make():
            built.append((name, len(log)))
            return _FakeBatch(name, iters, log)
        return make

    batches = [factory("a", 3), factory("b", 3), factory("c", 2), _FakeBatch("d", 4, log), factory("e", 0)]
    edges, creds = engine.trace_pipelined(batches, window=2, fit_merge=2)
    assert [c[0] for c in creds] == list("abcde") and edges[:, 0, 0].tolist() == [ord(c) for c in "abcde"]
    assert sorted(sum(fits, [])) == list("abcde") and fits[0] == ["a", "b"]            # converged together -> one fit
    assert [b[0] for b in built] == ["a", "b", "c", "e"] and built[2][1] > 0              # c built only after a left
    # never more than two sub-batches with a launched, unfinished iteration
    open_ = set()
    for ev, name in log:
        if ev == "release":
            continue
        open_.add(name) if ev == "launch" else open_.discard(name)
        assert len(open_) <= 2
    # every sub-batch released its loop buffers exactly once, after its last iteration and before its fit
    assert sorted(n for ev, n in log if ev == "release") == list("abcde")
    handle = engine.trace_pipelined([_FakeBatch("z", 1, log)], wait=False)
    assert isinstance(handle, engine.PipelinedResult) and handle.result()[1][0][0] == "z"


def test_grid_eigenbasis_rank_precheck():
    """Long spans: the Lanczos look at the leading eigenvalues decides "more than max_rank" without the full host
    decomposition for full-rank kernels, and leaves low-rank ones to the exact path (same basis as without the check)."""
    import numpy as np
    from gaussian_process_edge_trace_b200 import _gp_host as H
    xg = np.arange(1100)
    kd, Ur, lam, r = H.grid_eigenbasis("Matern", 2.5, 20.0, xg, 160)
    assert Ur is None and lam is None and r > 160 and kd.shape == (1100,) and kd[0] == 1.0
    kd, Ur, lam, r = H.grid_eigenbasis("RBF", 2.5, 60.0, xg, 160)            # rank ~ 60: exact path
    assert Ur is not None and Ur.shape[0] == 1100 and r <= 160 and Ur.shape[1] % 4 == 0
    K = kd[np.abs(xg[:, None] - xg[None, :])]
    assert np.abs(Ur @ np.diag(lam) @ Ur.T - K).max() < 1e-12
    kd2, Ur2, lam2, r2 = H.grid_eigenbasis("RBF", 2.5, 8.0, xg, 160)         # short length scale: rank > 160
    assert Ur2 is None and r2 > 160
