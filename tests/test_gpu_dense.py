"""GPU tests of the HBM-resident blocked path for training sets beyond the shared-memory kernels (csrc/gpet_dense.cu:
m > GPET_MAX_TRAIN = 224, BASELINE config 3 sizes): the blocked primitives against LAPACK (numpy, as the checker
only), every `_big` entry point against the shared-memory kernel it extends on sizes both accept, and whole traces
against the oracle with no torch.linalg factorisation on the path."""
import numpy as np
import pytest

import gpet_oracle as O

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import __graft_entry__
    __graft_entry__.build()
    import gaussian_process_edge_trace_b200 as p
    return p


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda")


def _padded_spd(rng, ms, ld):
    """SPD matrices of different sizes, each padded with an identity block to ld x ld."""
    A = np.zeros((len(ms), ld, ld))
    for b, m in enumerate(ms):
        G = rng.standard_normal((m, m + 3))
        A[b, :m, :m] = G @ G.T + 0.5 * m * np.eye(m)
        A[b, m:, m:] = np.eye(ld - m)
    return A


def test_blocked_cholesky_and_forward_substitution(pkg):
    """gpet_dense_potrf_f64 / gpet_dense_trsm_f64 (blocked right-looking Cholesky, left-looking forward substitution, DMMA
    tiles) on a batch of matrices of different sizes against LAPACK."""
    from gaussian_process_edge_trace_b200._cabi import call, ptr
    rng = np.random.default_rng(5)
    st = torch.cuda.current_stream().cuda_stream
    ms = np.array([300, 65, 64, 129, 2, 257, 320], dtype=np.int32)
    B, ld, ldr = len(ms), 320, 192
    A = _padded_spd(rng, ms, ld)
    dA, dm = _t(A), _t(ms)
    stat = torch.full((B,), 7, dtype=torch.int32, device="cuda")
    call("gpet_dense_potrf_f64", ptr(dA), ld, ptr(dm), B, int(ms.max()), ptr(stat), st)
    L = np.tril(dA.cpu().numpy())
    assert np.all(stat.cpu().numpy() == 0)
    for b, m in enumerate(ms):
        Lref = np.linalg.cholesky(A[b, :m, :m])
        assert np.abs(L[b, :m, :m] - Lref).max() <= 1e-12 * np.abs(Lref).max(), (b, m)
        mp = (m + 63) // 64 * 64
        assert np.array_equal(L[b, m:mp, :mp], np.eye(mp)[m:mp])          # identity padding survives
    # R <- L^-1 R
    R = rng.standard_normal((B, ld, ldr))
    for b, m in enumerate(ms):
        R[b, m:] = 0.0
    dR = _t(R)
    call("gpet_dense_trsm_f64", ptr(dA), ld, ptr(dm), B, int(ms.max()), ptr(dR), ldr, 0, st)
    X = dR.cpu().numpy()
    for b, m in enumerate(ms):
        ref = np.linalg.solve(L[b, :m, :m], R[b, :m])
        assert np.abs(X[b, :m] - ref).max() <= 1e-11 * np.abs(ref).max(), (b, m)
    # the inverse of the factor (lower tiles only; the buffer is never read)
    dT = torch.full((B, ld, ld), np.nan, dtype=torch.float64, device="cuda")
    call("gpet_dense_trsm_f64", ptr(dA), ld, ptr(dm), B, int(ms.max()), ptr(dT), ld, 1, st)
    Tm = dT.cpu().numpy()
    for b, m in enumerate(ms):
        ref = np.linalg.inv(L[b, :m, :m])
        got = np.tril(Tm[b, :m, :m])
        assert np.all(np.isfinite(got))
        assert np.abs(got - ref).max() <= 1e-11 * np.abs(ref).max(), (b, m)
    # a matrix that is not positive definite is reported, the others are unaffected
    A2 = A.copy()
    A2[3, 70, 70] = -1.0
    dA2 = _t(A2)
    call("gpet_dense_potrf_f64", ptr(dA2), ld, ptr(dm), B, int(ms.max()), ptr(stat), st)
    s = stat.cpu().numpy()
    assert s[3] == 1 and np.all(np.delete(s, 3) == 0)
    assert np.array_equal(np.tril(dA2.cpu().numpy())[0], L[0])


def test_big_posterior_equals_shared_memory_posterior(pkg):
    """The same training sets through the shared-memory posterior kernels (mmax = 100) and, padded to mmax = 230 > 224,
    through the blocked HBM path: mean, y_s, reduced covariance, full covariance; and against the oracle's posterior."""
    from gaussian_process_edge_trace_b200 import _gp_host
    from gaussian_process_edge_trace_b200._cabi import call, ptr, query
    rng = np.random.default_rng(12)
    st = torch.cuda.current_stream().cuda_stream
    B, n = 5, 300
    xg = np.arange(n)
    kd, Ur, lam, r = _gp_host.grid_eigenbasis("RBF", 2.5, 18.0, xg, 160)
    rp = Ur.shape[1]
    ms = np.array([2, 17, 64, 65, 100], dtype=np.int32)
    out = {}
    sets = []
    for b in range(B):
        xs = np.sort(rng.choice(n, size=ms[b], replace=False))
        sets.append((xs, rng.integers(0, 200, size=ms[b]).astype(np.float64)))
    for mmax in (100, 230):
        xi = np.zeros((B, mmax), dtype=np.int32); y = np.zeros((B, mmax)); w = np.zeros((B, mmax))
        for b in range(B):
            xi[b, : ms[b]], y[b, : ms[b]] = sets[b]
            w[b, : ms[b]] = 1.0
            w[b, 0] = w[b, ms[b] - 1] = 1e-7
        d = dict(xi=_t(xi), y=_t(y), w=_t(w), m=_t(ms), sf=torch.full((B,), 40.0, dtype=torch.float64, device="cuda"),
                 kd=_t(kd), Ur=_t(Ur), lam=_t(lam))
        nbytes = query("gpet_posterior_lowrank_workspace_bytes", B, mmax, rp)
        work = torch.empty(max(nbytes, 8), dtype=torch.uint8, device="cuda")
        mean = torch.zeros((B, n), dtype=torch.float64, device="cuda")
        ys = torch.zeros(B, dtype=torch.float64, device="cuda")
        Mr = torch.zeros((B, rp, rp), dtype=torch.float64, device="cuda")
        stat = torch.full((B,), 9, dtype=torch.int32, device="cuda")
        call("gpet_posterior_lowrank_f64", ptr(d["xi"]), ptr(d["y"]), ptr(d["w"]), ptr(d["m"]), mmax, 0, B, n, ptr(d["sf"]), 1.0,
             1e-6, ptr(d["kd"]), ptr(d["Ur"]), ptr(d["lam"]), rp, ptr(mean), ptr(ys), ptr(Mr), ptr(stat), ptr(work), st)
        cov = torch.zeros((B, n, n), dtype=torch.float64, device="cuda")
        wf = torch.empty(query("gpet_posterior_full_workspace_bytes", B, mmax, n), dtype=torch.uint8, device="cuda")
        mean2 = torch.zeros((B, n), dtype=torch.float64, device="cuda")
        stat2 = torch.full((B,), 9, dtype=torch.int32, device="cuda")
        call("gpet_posterior_full_f64", ptr(d["xi"]), ptr(d["y"]), ptr(d["w"]), ptr(d["m"]), mmax, B, n, ptr(d["sf"]), 1.0, 1e-6,
             ptr(d["kd"]), ptr(mean2), ptr(ys), ptr(cov), ptr(stat2), ptr(wf), st)
        out[mmax] = [a.cpu().numpy() for a in (mean, ys, Mr, cov, mean2, stat, stat2)]
    small, big = out[100], out[230]
    assert np.all(big[5] == 0) and np.all(big[6] == 0)
    assert np.array_equal(small[1], big[1])                                  # y_s: numpy-exact scalars on both paths
    for b in range(B):
        xs, yy = sets[b]
        wv = np.ones(ms[b]); wv[0] = wv[-1] = 1e-7
        post = O.posterior(xs.astype(np.float64), yy.copy(), wv, xg, "RBF", 2.5, 18.0, 40.0, 1.0)
        c = post["cov"].max()
        for res in (small, big):
            assert np.abs(res[3][b] - post["cov"]).max() <= 1e-11 * c, b
            assert np.abs(res[0][b] - post["mean"]).max() <= 1e-10 * max(1.0, np.abs(post["mean"]).max()), b
            assert np.abs(res[4][b] - post["mean"]).max() <= 1e-10 * max(1.0, np.abs(post["mean"]).max()), b
            assert np.abs(Ur @ res[2][b] @ Ur.T - post["cov"]).max() <= 1e-9 * c, b
        assert np.abs(big[2][b] - small[2][b]).max() <= 1e-10 * np.abs(small[2][b]).max(), b
        assert np.array_equal(big[3][b], big[3][b].T)                        # exactly symmetric


@pytest.mark.parametrize("kind,ktype,nu", [(0, "RBF", 2.5), (3, "Matern", 2.5), (2, "Matern", 1.5), (1, "Matern", 0.5)])
def test_big_objective_and_prediction_equal_shared_memory_kernels(pkg, kind, ktype, nu):
    """gpet_lml_big_f64 / gpet_final_predict_big_f64 (blocked, HBM) against the numpy objective and against gpet_lml_f64 /
    gpet_final_predict_f64 on training sets both accept; unused slots (trace_of < 0) are left alone; a workspace that
    holds only two evaluations at a time gives the same numbers."""
    from gaussian_process_edge_trace_b200 import _gp_host as H
    from gaussian_process_edge_trace_b200._cabi import call, ptr, query
    rng = np.random.default_rng(21)
    st = torch.cuda.current_stream().cuda_stream
    ms = np.array([150, 64, 3, 97], dtype=np.int32)
    T, mmax, n = len(ms), 150, 210
    X = np.zeros((T, mmax)); y = np.zeros((T, mmax)); w = np.zeros((T, mmax))
    for t, m in enumerate(ms):
        cols = np.sort(rng.choice(400, size=m, replace=False)).astype(np.float64)
        X[t, :m] = (cols - cols.mean()) / cols.std()
        yy = np.cumsum(rng.standard_normal(m))
        y[t, :m] = (yy - yy.mean()) / yy.std()
        w[t, :m] = 1.0
        w[t, 0] = w[t, m - 1] = 1e-7
    E = 24
    trace_of = rng.integers(0, T, size=E).astype(np.int32)
    trace_of[[5, 11]] = -1
    th = rng.uniform(H.FINAL_BOUNDS[:, 0] * 0.5, H.FINAL_BOUNDS[:, 1], size=(E, 3))
    th[:, 2] = rng.uniform(-5, 0, size=E)
    th[7] = [2.0, -6.0, -40.0]          # K ~ c I (length scale far below the spacing), noise at its lower bound
    dX, dy, dw, dm, dtr, dth = _t(X), _t(y), _t(w), _t(ms), _t(trace_of), _t(th)
    res = {}
    for name, nslots in (("small", None), ("big", E), ("big_chunked", 2)):
        df = torch.full((E,), 123.0, dtype=torch.float64, device="cuda")
        dg = torch.full((E, 3), 123.0, dtype=torch.float64, device="cuda")
        if nslots is None:
            call("gpet_lml_f64", ptr(dX), ptr(dy), ptr(dw), None, ptr(dm), mmax, ptr(dtr), ptr(dth), E, kind, 1e-6, ptr(df),
                 ptr(dg), st)
        else:
            nbytes = int(query("gpet_lml_big_workspace_bytes", nslots, mmax))
            work = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
            call("gpet_lml_big_f64", ptr(dX), ptr(dy), ptr(dw), ptr(dm), mmax, ptr(dtr), ptr(dth), E, kind, 1e-6, ptr(df),
                 ptr(dg), ptr(work), nbytes, st)
        res[name] = (df.cpu().numpy(), dg.cpu().numpy())
    assert np.array_equal(res["big"][0], res["big_chunked"][0]) and np.array_equal(res["big"][1], res["big_chunked"][1])
    fs, gs = res["small"]
    fb, gb = res["big"]
    for e in range(E):
        if trace_of[e] < 0:
            assert fb[e] == 123.0 and np.all(gb[e] == 123.0)
            continue
        if not np.isfinite(fs[e]):
            assert fb[e] == fs[e] and np.all(gb[e] == 0.0), e
            continue
        t, m = trace_of[e], ms[trace_of[e]]
        fo, go = H.neg_lml(th[e], X[t, :m], y[t, :m], w[t, :m], ktype, nu)
        assert abs(fb[e] - fo) <= 1e-10 * max(1.0, abs(fo)), (e, fb[e], fo)
        assert np.abs(gb[e] - go).max() <= 1e-8 * max(1.0, np.abs(go).max()), (e, gb[e], go)
        assert abs(fb[e] - fs[e]) <= 1e-10 * max(1.0, abs(fs[e])), (e, fb[e], fs[e])
        assert np.abs(gb[e] - gs[e]).max() <= 1e-8 * max(1.0, np.abs(gs[e]).max()), (e, gb[e], gs[e])
    # final prediction
    thp = th[:T].copy()
    xq = np.sort(rng.uniform(-1.8, 1.8, size=(T, n)), axis=1)
    tmts = np.stack([rng.uniform(-1, 1, size=T), rng.uniform(0.5, 2, size=T)], axis=1)
    dthp, dxq, dtm = _t(thp), _t(xq), _t(tmts)
    outs = {}
    for name in ("small", "big"):
        mean = torch.zeros((T, n), dtype=torch.float64, device="cuda")
        sd = torch.zeros((T, n), dtype=torch.float64, device="cuda")
        stat = torch.full((T,), 5, dtype=torch.int32, device="cuda")
        if name == "small":
            call("gpet_final_predict_f64", ptr(dX), ptr(dy), ptr(dw), ptr(dm), mmax, T, ptr(dthp), kind, 1e-6, ptr(dxq), n,
                 ptr(dtm), ptr(mean), ptr(sd), ptr(stat), st)
        else:
            work = torch.empty(int(query("gpet_final_predict_big_workspace_bytes", T, mmax, n)), dtype=torch.uint8,
                               device="cuda")
            call("gpet_final_predict_big_f64", ptr(dX), ptr(dy), ptr(dw), ptr(dm), mmax, T, ptr(dthp), kind, 1e-6, ptr(dxq), n,
                 ptr(dtm), ptr(mean), ptr(sd), ptr(stat), ptr(work), st)
        outs[name] = (mean.cpu().numpy(), sd.cpu().numpy(), stat.cpu().numpy())
    assert np.all(outs["big"][2] == 0) and np.all(outs["small"][2] == 0)
    assert np.abs(outs["big"][0] - outs["small"][0]).max() <= 1e-10 * max(1.0, np.abs(outs["small"][0]).max())
    # the variance is a difference of nearly equal numbers next to the end points (noise weight 1e-7): compare it, not its root
    assert np.abs(outs["big"][1] ** 2 - outs["small"][1] ** 2).max() <= 1e-9 * max(1.0, (outs["small"][1] ** 2).max())


def _trace_case(shape, amp, curv, seed, kernel_options, pixel_thresh):
    kern = O.kernel_builder((11, 5))
    img, edge = O.construct_test_img(shape, amp, curv, 0.002, "sinusoidal", 0.5, noise_seed=seed)
    grad = O.comp_grad_img(img, kern)
    init = edge[[0, -1], :][:, [1, 0]]
    kw = dict(kernel_options=kernel_options, noise_y=1, N_samples=300, score_thresh=1, delta_x=2, keep_ratio=0.2,
              pixel_thresh=pixel_thresh, seed=2, return_std=True, fix_endpoints=True)
    return init, grad, kw


def _forbid(monkeypatch, names):
    def no_library(*a, **k):
        raise AssertionError("torch.linalg must not be on this path")
    for nm in names:
        monkeypatch.setattr(torch.linalg, nm, no_library)
    monkeypatch.setattr(torch, "cholesky_solve", no_library)
    monkeypatch.setattr(torch, "cholesky_inverse", no_library)


def test_trace_beyond_shared_memory_limits_is_native(pkg, monkeypatch):
    """m > GPET_MAX_TRAIN = 224 (delta_x = 2 on a 520-pixel span -> up to 263 training points), RBF: blocked posterior in HBM
    feeding the low-rank factor, device L-BFGS-B on the blocked objective, blocked final prediction - no torch.linalg call
    anywhere; same stagewise parity bars as every other trace."""
    from test_gpu_parity import check_pair, run_pair
    init, grad, kw = _trace_case((40, 520), 10, 3, 4, {"kernel": "RBF", "sigma_f": 10, "length_scale": 14}, 8)
    _forbid(monkeypatch, ["eigh", "cholesky_ex", "cholesky", "solve_triangular"])
    tr, rec, orc, out, out_o = run_pair(pkg, init, grad, kw, "device")
    assert tr._tb.large_m and tr._tb.lowrank and tr._tb.mmax > 224
    assert max(o["X"].shape[0] for o in orc.record) > 224
    check_pair(tr, rec, orc, out, out_o)
    # the all-host final fit (scipy on the numpy objective) and scipy's setulb driving the blocked device objective
    # (GPET_FIT_DRIVER=host): same edge, same credible interval
    tr_h = pkg.gpet.GP_Edge_Tracing(init, grad, final_fit="host", **kw)
    eh, ch = tr_h()
    assert np.array_equal(eh, out[0])
    assert np.abs(ch[0] - out[1][0]).max() <= 1e-6 * np.abs(out[1][0]).max()
    monkeypatch.setenv("GPET_FIT_DRIVER", "host")
    tr_s = pkg.gpet.GP_Edge_Tracing(init, grad, **kw)
    es, cs = tr_s()
    assert np.array_equal(es, out[0])
    assert np.abs(cs[0] - out[1][0]).max() <= 1e-6 * np.abs(out[1][0]).max()


def test_fullrank_trace_beyond_shared_memory_limits(pkg, monkeypatch):
    """The same sizes with a Matern kernel (full rank: full covariance from the blocked path, factored by the block Jacobi
    eigensolver of gpet_jacobi.cu): no torch.linalg call either."""
    from test_gpu_parity import check_pair, run_pair
    init, grad, kw = _trace_case((40, 480), 10, 3, 6, {"kernel": "Matern", "nu": 2.5, "sigma_f": 10, "length_scale": 14}, 8)
    _forbid(monkeypatch, ["eigh", "cholesky_ex", "cholesky", "solve_triangular"])
    tr, rec, orc, out, out_o = run_pair(pkg, init, grad, kw, "device")
    assert tr._tb.large_m and not tr._tb.lowrank and tr._tb.mmax > 224
    check_pair(tr, rec, orc, out, out_o)


@pytest.mark.parametrize("n", [70, 128, 300])
def test_block_jacobi_eigensolver(pkg, n):
    """gpet_block_jacobi_* (two-sided block Jacobi in HBM, 128 x 128 pivots through gpet_sym_eig_f64) on covariance-like
    matrices (kernel matrix minus a low-rank part, graded spectrum, exact null vectors) and a random dense one:
    eigenvalues against LAPACK, orthogonality, F^T F = Sigma, descending order and the canonical sign rule."""
    from gaussian_process_edge_trace_b200 import _gp_host
    from gaussian_process_edge_trace_b200._cabi import call, ptr, query
    rng = np.random.default_rng(n)
    st = torch.cuda.current_stream().cuda_stream
    x = np.arange(n)[:, None]
    d = np.abs(x - x.T) / 14.0
    Kss = 100.0 * (1 + np.sqrt(5) * d + 5 * d * d / 3) * np.exp(-np.sqrt(5) * d)      # Matern nu = 2.5
    mats = []
    for m in (2, 25, n // 2):
        idx = np.sort(rng.choice(n, size=m, replace=False))
        Kmm = Kss[np.ix_(idx, idx)] + 1e-2 * np.eye(m)
        mats.append(Kss - Kss[:, idx] @ np.linalg.solve(Kmm, Kss[idx, :]))
    G = rng.standard_normal((n, n // 3))
    mats.append(G @ G.T)                                        # rank n / 3: a large exact null space
    cov = np.stack([0.5 * (M + M.T) for M in mats])
    B = cov.shape[0]
    np_ = (n + 127) // 128 * 128
    rp = (n + 3) // 4 * 4
    dcov = _t(cov)
    A = torch.empty((B, np_, np_), dtype=torch.float64, device="cuda")
    V = torch.empty((B, np_, np_), dtype=torch.float64, device="cuda")
    off = torch.empty((B, 2), dtype=torch.float64, device="cuda")
    work = torch.empty(int(query("gpet_block_jacobi_workspace_bytes", B, np_)), dtype=torch.uint8, device="cuda")
    call("gpet_block_jacobi_init_f64", ptr(dcov), B, n, np_, ptr(A), ptr(V), st)
    rels = []
    for sweep in range(20):
        call("gpet_block_jacobi_sweep_f64", ptr(A), ptr(V), B, np_, ptr(off), ptr(work), st)
        o = off.cpu().numpy()
        rels.append(np.sqrt(o[:, 0] / o[:, 1]).max())
        if rels[-1] <= 3e-13:
            break
    assert rels[-1] <= 3e-13 and len(rels) <= 12, rels
    w = _gp_host.sign_weights(n)
    F = torch.full((B, rp, n), np.nan, dtype=torch.float64, device="cuda")
    call("gpet_block_jacobi_factor_f64", ptr(A), ptr(V), B, n, np_, rp, ptr(_t(w)), ptr(F), ptr(work), st)
    F, Vh, Ah = F.cpu().numpy(), V.cpu().numpy(), A.cpu().numpy()
    for b in range(B):
        c = np.abs(cov[b]).max()
        assert np.abs(Vh[b].T @ Vh[b] - np.eye(np_)).max() < 1e-12
        ev = np.sort(np.diagonal(Ah[b]))[::-1][:n]
        assert np.abs(ev - np.linalg.eigvalsh(cov[b])[::-1]).max() < 1e-12 * c * n
        assert np.all(F[b, n:] == 0.0)
        assert np.abs(F[b, :n].T @ F[b, :n] - cov[b]).max() <= 1e-11 * c
        nrm = np.linalg.norm(F[b, :n], axis=1)
        assert np.all(np.diff(nrm) <= 1e-9 * nrm[0])
        assert np.all(F[b, :n] @ w >= -1e-12 * nrm[0])
