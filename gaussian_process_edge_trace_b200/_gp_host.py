"""Host-side (numpy/scipy) pieces of the hot path that stay on the CPU by design:

* tables that are computed once per (kernel, edge span) and uploaded: unit kernel by integer distance,
  eigenbasis of the unit kernel matrix on the grid (shared by every trace of a batch);
* the sequential control flow around the CUDA stages: training-set assembly (gpet.py:209-224), the adaptive
  score-threshold loop (gpet.py:589-609) on the per-bin maxima the GPU returns;
* the final hyper-parameter fit (gpet.py:232-248, 263-266; sklearn_gpr.py:254-295, 475-607), 13 L-BFGS-B runs
  driven by scipy on the restated log-marginal likelihood (scipy on host keeps the integer
  edge_pred bit-exact).
"""
import math

import numpy as np
import scipy.linalg
import scipy.optimize
from scipy.spatial.distance import cdist, pdist, squareform

GP_ALPHA = 1e-6          # gpet.py:155
KDE_THRESH = 1e-3        # gpet.py:109
RANK_REL_TOL = 1e-15     # eigenvalues of the unit kernel matrix below tol * max are dropped


def parse_kernel_options(kernel_options, M, edge_length):
    """gpet.py:130-151 -> (kernel_type, nu, sigma_f, length_scale)."""
    if type(kernel_options) == dict:
        sigma_f = kernel_options["sigma_f"]
        sigma_l = kernel_options["length_scale"]
        ktype = kernel_options["kernel"]
        nu = kernel_options["nu"] if kernel_options["kernel"] == "Matern" else 2.5
    else:
        rbf_matern, sf_opt, sl_opt = kernel_options
        ktype = ["RBF", "Matern"][int(rbf_matern > 0)]
        nu = [2.5, 1.5][int(rbf_matern > 1)]
        sf_const = [10, 8, 6, 4, 2, 1][sf_opt - 1] if (sf_opt >= 0) and (sf_opt <= 5) else 1
        sigma_f = M // sf_const
        sl_const = [1, 4 / 3, 2, 4, 10][sl_opt - 1] if (sl_opt >= 0) and (sl_opt <= 4) else 10
        sigma_l = edge_length // sl_const
    return ktype, nu, sigma_f, sigma_l


def stationary_kernel(ktype, nu, d):
    """k as a function of the scaled distance d >= 0 (sklearn RBF / Matern formulas)."""
    if ktype == "RBF":
        return np.exp(-0.5 * d ** 2)
    if ktype != "Matern":
        raise ValueError(f"unknown kernel {ktype!r}")
    if nu == 0.5:
        return np.exp(-d)
    if nu == 1.5:
        t = d * math.sqrt(3)
        return (1.0 + t) * np.exp(-t)
    if nu == 2.5:
        t = d * math.sqrt(5)
        return (1.0 + t + t ** 2 / 3.0) * np.exp(-t)
    if nu == np.inf:
        return np.exp(-(d ** 2) / 2.0)
    raise NotImplementedError("Matern with general nu is not supported")


def kernel_by_distance(ktype, nu, length_scale, x_grid):
    """kd[d] = k(|x_i - x_j| = d) for integer d = 0..n-1 on the pixel grid; kd[0] = 1 exactly."""
    xs = np.asarray(x_grid, dtype=np.float64) / float(length_scale)
    kd = stationary_kernel(ktype, nu, np.abs(xs - xs[0]))
    kd[0] = 1.0
    return kd


_basis_cache = {}


def grid_eigenbasis(ktype, nu, length_scale, x_grid, max_rank):
    """Leading eigenpairs of the unit kernel matrix on the grid. Returns (kd, Ur[n, rp], lam[rp], r) with
    rp = r rounded up to a multiple of 4 (zero padded: the k-step of the fp64 DMMA; the Jacobi solver needs it even), or
    (kd, None, None, r) when r > max_rank."""
    n = len(x_grid)
    key = (ktype, float(nu), float(length_scale), int(x_grid[0]), n)
    if key not in _basis_cache:
        kd = kernel_by_distance(ktype, nu, length_scale, x_grid)
        K = scipy.linalg.toeplitz(kd)
        if n >= 1024 and n > 4 * (max_rank + 1):
            # long spans: a Lanczos look at the leading max_rank + 1 eigenvalues first - when even the last of them is far
            # above the rank threshold (Matern; RBF with a short length scale) the full decomposition (15 s of host
            # LAPACK at n = 4096) is not needed, the caller takes the full-covariance path
            from scipy.sparse.linalg import eigsh
            top = eigsh(K, k=max_rank + 1, which="LA", return_eigenvectors=False, tol=1e-3)
            if top.min() > 1e6 * RANK_REL_TOL * top.max():
                _basis_cache[key] = (kd, None, None)
        if key not in _basis_cache:
            lam, U = np.linalg.eigh(K)
            _basis_cache[key] = (kd, lam[::-1].copy(), U[:, ::-1].copy())
    kd, lam, U = _basis_cache[key]
    if lam is None:
        return kd, None, None, max_rank + 1        # "more than max_rank"
    r = int(np.sum(lam > RANK_REL_TOL * lam[0]))
    if r > max_rank:
        return kd, None, None, r
    rp = max(8, ((r + 3) // 4) * 4)
    Ur = np.zeros((n, rp))
    Ur[:, :r] = U[:, :r]
    lr = np.zeros(rp)
    lr[:r] = lam[:r]
    return kd, Ur, lr, r


def sign_weights(n):
    """Fixed generic weight vector of the canonical sign rule: <Vt[k], w> > 0."""
    return 1.0 + np.arange(n, dtype=np.float64) / n


def canonical_factor_host(cov):
    """Parity-mode factor provider: A = diag(sqrt(s)) Vt of numpy.linalg.svd(cov) - exactly what numpy's
    multivariate_normal uses (sklearn_gpr.py:464) - with the canonical signs."""
    _, s, vt = np.linalg.svd(cov)
    sg = np.sign(vt @ sign_weights(vt.shape[1]))
    sg[sg == 0] = 1.0
    return np.sqrt(s)[:, None] * (vt * sg[:, None])


def assemble_training_set(init_sorted, obs_xy, alpha_init):
    """gpet.py:209-214, 223-224. Returns (x int64[m], y float64[m], w float64[m]) sorted by x."""
    obs_xy = np.asarray(obs_xy).reshape(-1, 2)
    w = np.concatenate([alpha_init, np.ones(obs_xy.shape[0])], axis=0)
    pts = np.concatenate([init_sorted, obs_xy], axis=0)
    order = np.argsort(pts[:, 0], kind="stable")
    pts = pts[order]
    return pts[:, 0].astype(np.int64), pts[:, 1].astype(np.float64), w[order]


def column_bins(N, x_st, x_en, delta_x, fix_endpoints, max_group=48):
    """Bin of every image column (np.round((x - x_st)/delta_x), gpet.py:605-606), shifted to start at 0, in the
    encoding gpet_select_f64 expects, plus column groups that never split a bin.
    Returns (col_bin int32[N], group_cols int32[G+1], nb, bin_lo)."""
    x = np.arange(N)
    bins = np.round((x - x_st) / delta_x).astype(np.int64)
    bin_lo = int(bins.min())
    bins = bins - bin_lo
    nb = int(bins.max()) + 1
    cand = (x > x_st) & (x < x_en) if fix_endpoints else np.ones(N, dtype=bool)
    col_bin = np.where(cand, bins, -(bins + 1)).astype(np.int32)
    bin_starts = np.concatenate([[0], np.flatnonzero(np.diff(bins)) + 1, [N]])
    starts = [0]
    for k in range(len(bin_starts) - 1):
        bs, be = int(bin_starts[k]), int(bin_starts[k + 1])
        if be - starts[-1] > max_group and bs > starts[-1]:
            starts.append(bs)
        if be - starts[-1] > 64:
            raise ValueError("a single bin spans more than 64 columns (delta_x too large for the select kernel)")
    group_cols = np.array(starts + [N], dtype=np.int32)
    if np.any(np.diff(group_cols) > 64) or np.any(np.diff(group_cols) <= 0):
        raise ValueError("column grouping failed")
    return col_bin, group_cols, nb, bin_lo


def threshold_loop_batch(best, n_pre, pixel_thresh, algo_thresh, thr, active, max_decays=4000):
    """Decay loop of gpet.py:589-609 on the per-bin maxima, for all traces at once and without iterating:
    best[B, nb] (-1 = empty bin), n_pre[B], thr[B] (updated in place for `active` traces). Returns mask[B, nb] of
    accepted bins.

    The reference multiplies the threshold by 0.95 (by 1.0 on the first pass) until the number of bins at or above it
    reaches T = min(n_pre + pixel_thresh, algo_thresh). That count is monotone in the threshold, so the loop stops at
    the first element of the sequence thr, thr*0.95, (thr*0.95)*0.95, ... that is <= the T-th largest bin maximum. The
    sequence is produced by a left fold (np.multiply.accumulate), i.e. with the reference's own roundings."""
    B, nb = best.shape
    mask = np.zeros(best.shape, dtype=bool)
    rows = np.flatnonzero(active & (pixel_thresh > 0) & (n_pre < algo_thresh))
    if rows.shape[0] == 0:
        return mask
    bs = best[rows]
    target = np.minimum(n_pre[rows] + pixel_thresh, algo_thresh).astype(np.int64)      # bins needed to stop
    srt = -np.sort(-np.where(bs >= 0, bs, -np.inf), axis=1)                              # descending
    enough = target <= nb
    v_t = np.where(enough, srt[np.arange(rows.shape[0]), np.minimum(target, nb) - 1], -np.inf)
    if np.any(~np.isfinite(v_t)) or np.any(v_t <= 0.0):
        # fewer non-empty bins than needed (or only zero scores): the reference's loop never ends
        raise RuntimeError("compute_new_obs: score threshold decayed to zero without enough new pixels "
                           "(the reference loops forever here, gpet.py:591-609)")
    t0 = thr[rows]
    L = 64
    while True:
        seq = np.multiply.accumulate(np.concatenate([t0[:, None], np.full((rows.shape[0], L), 0.95)], axis=1), axis=1)
        hit = seq <= v_t[:, None]
        done = hit.any(axis=1)
        if done.all():
            break
        if L >= max_decays:
            raise RuntimeError("compute_new_obs: score threshold decayed to zero without enough new pixels "
                               "(the reference loops forever here, gpet.py:591-609)")
        L = min(max_decays, 4 * L)
    first = hit.argmax(axis=1)
    t_fin = seq[np.arange(rows.shape[0]), first]
    thr[rows] = t_fin
    mask[rows] = (bs >= t_fin[:, None]) & (bs >= 0)
    return mask


# ---------------------------------------------------------------------------------------------------
# final hyper-parameter fit
# ---------------------------------------------------------------------------------------------------
FINAL_BOUNDS = np.log(np.array([[0.01, 1e3], [0.1, 100.0], [1e-18, 1.0]]))   # gpet.py:246-248


def _k_and_grad(theta, X, w, ktype, nu):
    """K and dK/dtheta of Constant*(RBF|Matern) + WeightedWhite, theta = log[const, length_scale, noise]
    (sklearn kernels.py Product/Sum composition; sklearn_gpr.py:684-688 for the noise term)."""
    const, ls, noise = np.exp(theta)
    Xc = X.reshape(-1, 1)
    m = Xc.shape[0]
    if ktype == "RBF":
        d2 = pdist(Xc / ls, metric="sqeuclidean")
        k = squareform(np.exp(-0.5 * d2))
        np.fill_diagonal(k, 1)
        dk = k * squareform(d2)
    else:
        d = pdist(Xc / ls, metric="euclidean")
        D = squareform(d ** 2)
        k = squareform(stationary_kernel(ktype, nu, d))
        np.fill_diagonal(k, 1)
        if nu == 0.5:
            den = np.sqrt(D)
            q = np.zeros_like(D)
            np.divide(D, den, out=q, where=den != 0)
            dk = k * q
        elif nu == 1.5:
            dk = 3 * D * np.exp(-np.sqrt(3 * D))
        elif nu == 2.5:
            tmp = np.sqrt(5 * D)
            dk = 5.0 / 3.0 * D * (tmp + 1) * np.exp(-tmp)
        else:
            raise NotImplementedError("final fit supports Matern nu in {0.5, 1.5, 2.5}")
    K1 = np.full((m, m), const)
    Kww = noise * np.diag(w)
    K = K1 * k + Kww
    dK = np.dstack((K1[:, :, None] * k[:, :, None], dk[:, :, None] * K1[:, :, None], Kww[:, :, None]))
    return K, dK


def neg_lml(theta, X, y, w, ktype, nu):
    """-(log marginal likelihood, gradient): sklearn_gpr.py:512-583 with the sign flip of obj_func (:257-262)."""
    K, dK = _k_and_grad(theta, X, w, ktype, nu)
    K[np.diag_indices_from(K)] += GP_ALPHA
    try:
        L = scipy.linalg.cholesky(K, lower=True, check_finite=False)
    except np.linalg.LinAlgError:
        return np.inf, np.zeros_like(theta)
    yt = y[:, None]
    a = scipy.linalg.cho_solve((L, True), yt, check_finite=False)
    lml = -0.5 * np.einsum("ik,ik->k", yt, a)
    lml -= np.log(np.diag(L)).sum()
    lml -= K.shape[0] / 2 * np.log(2 * np.pi)
    lml = lml.sum(axis=-1)
    inner = np.einsum("ik,jk->ijk", a, a)
    Kinv = scipy.linalg.cho_solve((L, True), np.eye(K.shape[0]), check_finite=False)
    inner -= Kinv[..., None]
    grad = (0.5 * np.einsum("ijl,jik->kl", inner, dK)).sum(axis=-1)
    return -lml, -grad


def final_fit(X, y, w, x_grid, ktype, nu, noise_y, seed, n_restarts=12):
    """Converged branch of fit_predict_GP. Returns (y_mean[n], y_std[n], theta)."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    y_m, y_s = np.mean(y), np.std(y)
    y = (y - y_m) / y_s
    X_m, X_s = np.mean(X), np.std(X)
    X = (X - X_m) / X_s
    tm, ts = np.mean(y), np.std(y)           # GPR centres and scales once more (sklearn_gpr.py:229-234)
    if ts < 10 * np.finfo(np.float64).eps:
        ts = 1.0
    yt = (y - tm) / ts
    rng = np.random.RandomState(seed)         # sklearn_gpr.py:205
    starts = [np.log(np.array([5.0, 5.0, float(noise_y)]))]
    best_x, best_f = None, np.inf
    for i in range(n_restarts + 1):
        t0 = starts[0] if i == 0 else rng.uniform(FINAL_BOUNDS[:, 0], FINAL_BOUNDS[:, 1])
        res = scipy.optimize.minimize(neg_lml, t0, args=(X, yt, w, ktype, nu), method="L-BFGS-B", jac=True,
                                      bounds=FINAL_BOUNDS)
        if res.fun < best_f:                  # np.argmin keeps the first minimum
            best_x, best_f = res.x, res.fun
    theta = best_x
    K, _ = _k_and_grad(theta, X, w, ktype, nu)
    K[np.diag_indices_from(K)] += GP_ALPHA
    L = scipy.linalg.cholesky(K, lower=True, check_finite=False)
    a = scipy.linalg.cho_solve((L, True), yt, check_finite=False)
    const, ls, _ = np.exp(theta)
    xs = ((np.asarray(x_grid) - X_m) / X_s).reshape(-1, 1)
    if ktype == "RBF":
        kx = np.exp(-0.5 * cdist(xs / ls, X.reshape(-1, 1) / ls, metric="sqeuclidean"))
    else:
        kx = stationary_kernel(ktype, nu, cdist(xs / ls, X.reshape(-1, 1) / ls, metric="euclidean"))
    Ks = const * kx
    mu = ts * (Ks @ a) + tm
    V = scipy.linalg.solve_triangular(L, Ks.T, lower=True, check_finite=False)
    var = np.full(xs.shape[0], const) * np.ones(xs.shape[0])
    var -= np.einsum("ij,ji->i", V.T, V)
    var[var < 0] = 0.0
    return y_s * mu + y_m, np.sqrt(var * ts ** 2), theta
