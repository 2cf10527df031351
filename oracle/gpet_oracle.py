"""TEST INFRASTRUCTURE ONLY - CPU restatement (numpy/scipy) of the reference hot path.

This module is the parity ORACLE of gaussian_process_edge_trace_b200. It is imported only by
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py`; the product package never imports it and has no CPU fallback.

Every function cites the reference lines (relative to /root/reference/gp_edge_tracing/) it
restates. Pinning: `oracle/make_golden.py` runs the UNMODIFIED reference in the build
container (through `oracle/ref_harness.py`) and stores per-stage golden vectors under
`tests/golden/`; `tests/test_oracle_golden.py` checks this module against them. Two boundaries
remain "parity unpinned" because the third-party package is absent from the image and not
vendored by the reference: KDEpy.FFTKDE (restated in `kde_blur`/`linear_bin_columns`) and
skimage.util.random_noise (bench-input noise only).

Conventions: curves are float64[n, S] (column = one posterior curve, exactly the
reference's `y_samples`), observations `fobs` are int64[k, 2] in xy order, images are [row, col].
"""
import math

import numpy as np
import scipy.integrate
import scipy.interpolate
import scipy.linalg
import scipy.ndimage
import scipy.optimize
import scipy.signal
from scipy.spatial.distance import cdist, pdist, squareform

KDE_THRESH = 1e-3  # gpet.py:109


# --------------------------------------------------------------------------------------
# gpet_utils.py
# --------------------------------------------------------------------------------------
def kernel_builder(size, b2d=False, normalize=False, vertical_edges=False, unit=False):
    """gpet_utils.py:10-61. Sobel-like filter; (11,5) default rows documented in SURVEY A.6."""
    rows, cols = size
    mid_r, mid_c = rows // 2, cols // 2
    kernel = np.zeros(size)
    if unit:
        kernel[:mid_r, :] = 1
    else:
        i = np.arange(mid_r)[:, None]
        j = np.arange(cols)[None, :]
        kernel[:mid_r, :] = 1 + np.maximum(0, mid_r + 1 - np.abs(i - mid_r) - np.abs(j - mid_c))
    kernel[mid_r + 1:, :] = -np.flip(kernel[0:mid_r, :], axis=0)
    if b2d:
        kernel = np.flipud(kernel)
    if vertical_edges:
        kernel = kernel.T
    if normalize:
        kernel = kernel / kernel.max()
    return kernel


def normalise(img, minmax_val=(0, 1), astyp=np.float32):
    """gpet_utils.py:65-91. float32 min-max; divides by the max taken AFTER subtracting the min."""
    lo, hi = minmax_val
    img = np.asarray(img).astype(np.float32)
    img -= img.min()
    img /= img.max()
    img *= (hi - lo)
    img += lo
    return img.astype(astyp)


def comp_grad_img(img, kernel, norm=True, astyp=np.float32):
    """gpet_utils.py:95-119. `norm` has no effect in the reference (`if normalise:` tests the
    function object, :114), so the output is always the float32 min-max normalised map."""
    grad = scipy.ndimage.convolve(img, kernel, mode="nearest")
    grad[np.where(grad < 0)] = 0
    return normalise(grad, (0, 1), astyp)


def comp_grad_img_explicit(img, kernel):
    """Same result as `comp_grad_img` written as an explicit edge-padded correlation with the
    flipped kernel; used to cross-check the ndimage semantics."""
    kh, kw = kernel.shape
    P = np.pad(np.asarray(img, dtype=np.float64), ((kh // 2, kh // 2), (kw // 2, kw // 2)), mode="edge")
    Kf = kernel[::-1, ::-1]
    M, N = img.shape
    G = np.zeros((M, N))
    for a in range(kh):
        for b in range(kw):
            if Kf[a, b] != 0:
                G += P[a:a + M, b:b + N] * Kf[a, b]
    G[G < 0] = 0
    return normalise(G, (0, 1), np.float32)


def gaussian_noise(image, seed, mean=0.0, var=0.01):
    """Stand-in for skimage.util.random_noise(mode='gaussian') used at gpet_utils.py:251
    (same arithmetic as oracle/shims/skimage/util.py)."""
    rng = np.random.default_rng(seed)
    return np.clip(image + rng.normal(mean, var ** 0.5, image.shape), 0.0, 1.0)


def construct_test_img(size, amplitude, curvature, noise_level, ltype="sinusoidal", intensity=0.3,
                       gaps=False, noise_seed=1):
    """gpet_utils.py:163-253 for ltype in {'sinusoidal','co-sinusoidal','straight','diag'}.
    The reference hard-codes the noise seed to 1 (:251); `noise_seed` generalises that for
    multi-image benches."""
    M, N = size
    img = np.zeros((M, N))
    x = np.linspace(-np.pi, np.pi, N)
    A = M // 2 if amplitude > M else amplitude // 2
    cols = np.arange(N)
    if ltype == "sinusoidal":
        rows = (np.rint(A * np.sin(N * curvature * x)) + M // 2).astype(int)
    elif ltype == "co-sinusoidal":
        rows = (np.rint(A * np.cos(N * curvature * x)) + M // 2).astype(int)
    elif ltype == "straight":
        rows = np.full(N, M // 2, dtype=int)
    elif ltype == "diag":
        rows = cols.copy()
    else:
        raise NotImplementedError(ltype)
    for j in range(N):
        img[rows[j]:M, j] = intensity
    edge_idx = np.stack([rows, cols], axis=1)
    if gaps:
        img[:, 20:30] = 0
        img[:, N // 2:(N // 2 + 10)] = 0
        img[:, N - 100:N - 90] = 0
        img[:, N // 4:(N // 4 + 20)] = 0
    img = gaussian_noise(img, noise_seed, 0.0, noise_level)
    return img, edge_idx


# --------------------------------------------------------------------------------------
# KDEpy.FFTKDE restatement (gpet.py:455-529)
# --------------------------------------------------------------------------------------
def gaussian_taps_2d():
    """9x9 kernel exp(-(dx^2+dy^2)/2)/(2 pi), offsets -4..4 (KDEpy practical support L=4)."""
    o = np.arange(-4.0, 5.0)
    r2 = o[:, None] ** 2 + o[None, :] ** 2
    return np.exp(-0.5 * r2) / (2 * np.pi)


def kde_blur(binned_xy):
    """scipy.signal.convolve(binned, taps, mode='same') on the (N+2, M+2) x-major grid."""
    return scipy.signal.convolve(binned_xy, gaussian_taps_2d(), mode="same")


def linear_bin_columns(xs, ys, weights, N, M):
    """KDEpy linear binning on the integer lattice x in [-1..N], y in [-1..M] for points whose
    x is an integer (fx = 0): two taps down the column. Accumulates sequentially in point
    order. Returns the x-major grid (N+2, M+2)."""
    w = np.asarray(weights, dtype=np.float64)
    w = w / np.sum(w)
    ty = np.asarray(ys, dtype=np.float64) + 1.0
    iy = np.floor(ty)
    fy = ty - iy
    iy = iy.astype(np.int64)
    ix = np.asarray(xs, dtype=np.int64) + 1
    out = np.zeros((N + 2) * (M + 2))
    base = ix * (M + 2) + iy
    idx = np.stack([base, base + 1], axis=1).ravel()
    val = np.stack([(1.0 - fy) * w, fy * w], axis=1).ravel()
    ok = (idx >= 0) & (idx < out.size)
    np.add.at(out, idx[ok], val[ok])
    return out.reshape(N + 2, M + 2)


def kde_of_curves(Y_keep, costs, x_grid, M, N):
    """gpet.py:485-500, 514-527. Y_keep float64[n, Kp] = kept curves in ascending-cost order.
    Returns float64[M, N] (float32-representable values)."""
    n, Kp = Y_keep.shape
    inv = 1 / costs
    weights = inv / np.sum(inv)
    xs = np.repeat(np.asarray(x_grid), Kp)        # point order: column-major j, then curve c
    ys = Y_keep.reshape(-1)
    ws = np.tile(weights, (n, 1)).reshape(-1)
    keep = ~((ys < 0) | (ys > M - 1))
    binned = linear_bin_columns(xs[keep], ys[keep], ws[keep], N, M)
    dens = kde_blur(binned).T[1:-1, 1:-1]
    return normalise(dens, (0, 1), np.float64)


def kde_of_gradient(G):
    """gpet.py:503-509, 514-527. KDE of the gradient image: points = pixels with G > 1e-3,
    weights = G."""
    M, N = G.shape
    pts = np.argwhere(G > KDE_THRESH)
    w = G[pts[:, 0], pts[:, 1]]
    binned = linear_bin_columns(pts[:, 1], pts[:, 0].astype(np.float64), w, N, M)
    dens = kde_blur(binned).T[1:-1, 1:-1]
    return normalise(dens, (0, 1), np.float64)


# --------------------------------------------------------------------------------------
# GP kernels (sklearn RBF / Matern semantics) and the posterior (gpet.py:182-268,
# sklearn_gpr.py:183-321, 379-473)
# --------------------------------------------------------------------------------------
def unit_kernel(kind, nu, length_scale, X, Y=None):
    """sklearn.gaussian_process.kernels.RBF / Matern __call__ for 1-D inputs (column vectors),
    without the constant factor. k(X) has an exact unit diagonal."""
    X = np.asarray(X, dtype=np.float64).reshape(-1, 1)
    ls = float(length_scale)
    if kind == "RBF":
        if Y is None:
            K = squareform(np.exp(-0.5 * pdist(X / ls, metric="sqeuclidean")))
            np.fill_diagonal(K, 1)
            return K
        Y = np.asarray(Y, dtype=np.float64).reshape(-1, 1)
        return np.exp(-0.5 * cdist(X / ls, Y / ls, metric="sqeuclidean"))
    if kind != "Matern":
        raise ValueError(kind)
    if Y is None:
        d = pdist(X / ls, metric="euclidean")
    else:
        Y = np.asarray(Y, dtype=np.float64).reshape(-1, 1)
        d = cdist(X / ls, Y / ls, metric="euclidean")
    if nu == 0.5:
        K = np.exp(-d)
    elif nu == 1.5:
        K = d * math.sqrt(3)
        K = (1.0 + K) * np.exp(-K)
    elif nu == 2.5:
        K = d * math.sqrt(5)
        K = (1.0 + K + K ** 2 / 3.0) * np.exp(-K)
    elif nu == np.inf:
        K = np.exp(-(d ** 2) / 2.0)
    else:
        raise NotImplementedError("general-nu Matern (Bessel) is not restated")
    if Y is None:
        K = squareform(K)
        np.fill_diagonal(K, 1)
    return K


def assemble_training_set(init_sorted, obs_xy, alpha_init):
    """gpet.py:209-214, 223-224: concat(init, obs), sort by x; per-point noise weights."""
    obs_xy = np.asarray(obs_xy).reshape(-1, 2)
    alpha = np.concatenate([alpha_init, np.ones(obs_xy.shape[0])], axis=0)
    pts = np.concatenate([init_sorted, obs_xy], axis=0)
    order = np.argsort(pts[:, 0], kind="stable")
    alpha, pts = alpha[order], pts[order]
    return pts[:, 0].astype(np.float64), pts[:, 1].astype(np.float64), alpha


def posterior(X, y, w, x_grid, kind, nu, length_scale, sigma_f, noise_y, gp_alpha=1e-6):
    """Non-converged branch. gpet.py:227-230, 253-261 + sklearn_gpr.py:221-227 (mean removed,
    NOT scaled, but std kept), :304-320 (K, chol, alpha), :381-407 (mean uses the kept std -
    reference quirk; cov = (K** - V^T V) * std^2). Returns dict."""
    y = np.array(y, dtype=np.float64)
    y_s = np.std(y) + 1
    y /= y_s
    c = sigma_f ** 2 / y_s ** 2
    ybar = np.mean(y)
    sy = np.std(y)
    if sy < 10 * np.finfo(np.float64).eps:      # sklearn _handle_zeros_in_scale
        sy = 1.0
    y = y - ybar
    K = c * unit_kernel(kind, nu, length_scale, X)
    if X.shape[0] != x_grid.shape[0]:            # WeightedWhiteKernel edge_length quirk, sklearn_gpr.py:672-677
        K = K + noise_y * np.diag(w)
    K[np.diag_indices_from(K)] += gp_alpha
    L = scipy.linalg.cholesky(K, lower=True, check_finite=False)
    a = scipy.linalg.cho_solve((L, True), y, check_finite=False)
    Ks = c * unit_kernel(kind, nu, length_scale, x_grid, X)
    mu = sy * (Ks @ a) + ybar
    V = scipy.linalg.solve_triangular(L, Ks.T, lower=True, check_finite=False)
    cov = (c * unit_kernel(kind, nu, length_scale, x_grid) - V.T @ V) * sy ** 2
    return dict(mean=mu, cov=cov, y_s=y_s, c=c, ybar=ybar, sy=sy, L=L, alpha=a)


def sign_weights(n):
    """Fixed generic weights of the canonical sign rule."""
    return 1.0 + np.arange(n, dtype=np.float64) / n


def canonical_factor(cov):
    """A = diag(sqrt(s)) Vt of numpy.linalg.svd(cov) with signs pinned so <Vt[k], w> > 0.
    numpy's legacy multivariate_normal (called at sklearn_gpr.py:464) computes
    Z @ (sqrt(s)[:,None]*Vt) + mean with this SVD."""
    _, s, vt = np.linalg.svd(cov)
    sg = np.sign(vt @ sign_weights(vt.shape[1]))
    sg[sg == 0] = 1.0
    return np.sqrt(s)[:, None] * (vt * sg[:, None])


def standard_normals(seed, S, n):
    """RandomState(seed).standard_normal((S, n)) - the draw made inside multivariate_normal."""
    return np.random.RandomState(seed).standard_normal((S, n))


def sample_curves(Z, A, mean, y_s):
    """sklearn_gpr.py:460-464 + gpet.py:261 -> float64[n, S]. A may have r <= n rows (then only
    Z[:, :r] is used - a truncated factor)."""
    r = A.shape[0]
    return (Z[:, :r] @ A + mean).T * y_s


# --------------------------------------------------------------------------------------
# Curve cost (gpet.py:336-410) and top-N_keep (gpet.py:414-451)
# --------------------------------------------------------------------------------------
def make_grad_interp(G):
    """gpet.py:122-125. Same scipy object the reference builds (bilinear, coefficients == data)."""
    M, N = G.shape
    return scipy.interpolate.RectBivariateSpline(np.arange(M), np.arange(N), G, kx=1, ky=1)


def cost_funct(grad_interp, edge):
    """gpet.py:371-410 verbatim in behaviour (one curve, xy rows). This per-curve form is what
    the reference loops over in Python (gpet.py:437-440) and is the CPU baseline."""
    edge = edge[edge[:, 0].argsort(), :]
    grad_score = grad_interp(edge[:, 1], edge[:, 0], grid=False) + KDE_THRESH
    pixel_diff = np.cumsum(np.sqrt(np.sum(np.diff(edge, axis=0) ** 2, axis=1)))
    yy = edge[:, 1]
    pixel_deriv = yy[1:] - yy[:-1]                     # finite_diff(typ=0, h=1), gpet.py:360-365
    integrand = np.sqrt(1 + pixel_deriv ** 2)
    line_integral = scipy.integrate.simpson(grad_score[:-1], x=pixel_diff)
    arc_length = scipy.integrate.simpson(integrand, x=edge[:-1, 0])
    return arc_length / line_integral


def costs_loop(G, Y, x_grid):
    """gpet.py:434-440: the reference's per-curve Python loop."""
    gi = make_grad_interp(G)
    xg = np.asarray(x_grid, dtype=np.float64)
    return np.asarray([cost_funct(gi, np.stack([xg, Y[:, i]], axis=1)) for i in range(Y.shape[1])])


def _simpson_nonuniform(y, x):
    """scipy.integrate.simpson along axis 0 for any sample count K >= 3: `_basic_simpson` (non-uniform branch) on an
    odd K; on an even K the first K - 1 samples plus Cartwright's correction of the last interval, as scipy >= 1.11 does
    (scipy/integrate/_quadrature.py `simpson`; the reference calls the `simps` alias, gpet.py:404-405)."""
    K = y.shape[0]
    if K % 2 == 0:
        h0, h1 = x[K - 2] - x[K - 3], x[K - 1] - x[K - 2]
        alpha = (2 * h1 ** 2 + 3 * h0 * h1) / (6 * (h1 + h0))
        beta = (h1 ** 2 + 3.0 * h0 * h1) / (6 * h0)
        eta = (1 * h1 ** 3) / (6 * h0 * (h0 + h1))
        return _simpson_nonuniform(y[:-1], x[:-1]) + (alpha * y[K - 1] + beta * y[K - 2] - eta * y[K - 3])
    h = np.diff(x, axis=0)
    h0, h1 = h[0::2], h[1::2]
    hs, hp, r = h0 + h1, h0 * h1, h0 / h1
    return np.sum(hs / 6.0 * (y[0:-2:2] * (2.0 - 1.0 / r) + y[1:-1:2] * (hs * (hs / hp)) + y[2::2] * (2.0 - r)), axis=0)


def costs_vectorised(G, Y, x_grid):
    """all curves at once (validated against `costs_loop` to ~1e-15 relative), any edge_length >= 5."""
    M = G.shape[0]
    xg = np.asarray(x_grid, dtype=np.int64)
    yc = np.clip(Y, 0, M - 1)
    i0 = np.minimum(np.floor(yc), M - 2).astype(np.int64)
    f = yc - i0
    cols = xg[:, None]
    g = G[i0, cols] * ((i0 + 1) - yc) + G[i0 + 1, cols] * f + KDE_THRESH
    dy = Y[1:] - Y[:-1]
    seg = np.sqrt(1 + dy ** 2)
    t = np.cumsum(seg, axis=0)
    LI = _simpson_nonuniform(g[:-1], t)
    AL = _simpson_nonuniform(seg, xg[:-1, None].astype(np.float64) * np.ones((1, Y.shape[1])))
    return AL / LI


def top_keep(costs, N_keep):
    """gpet.py:443-445: ascending-cost order matters (weights and 'optimal curve')."""
    idx = np.argsort(costs)[:N_keep]
    return idx, costs[idx]


# --------------------------------------------------------------------------------------
# Candidate scoring / selection (gpet.py:532-662)
# --------------------------------------------------------------------------------------
def compute_new_obs(kde, grad_kde, pre_fobs_yx, x_st, x_en, delta_x, pixel_thresh, algo_thresh,
                    score_thresh, fix_endpoints=True, max_decays=4000):
    """gpet.py:622-662 + :532-618, thresholding first and arg-maxing per bin afterwards exactly
    like the reference. Returns (fobs int64[k,2] xy, new score_thresh). `max_decays` guards
    the reference's latent infinite loop (documented deviation: raises instead of hanging)."""
    pixel_idx = np.argwhere(kde > KDE_THRESH)
    if fix_endpoints:
        pixel_idx = pixel_idx[(pixel_idx[:, 1] > x_st) & (pixel_idx[:, 1] < x_en)]
    pre = np.asarray(pre_fobs_yx, dtype=np.int64).reshape(-1, 2)
    N_pre = pre.shape[0]
    new_grad = grad_kde[pixel_idx[:, 0], pixel_idx[:, 1]]
    new_int = kde[pixel_idx[:, 0], pixel_idx[:, 1]]
    old_int = kde[pre[:, 0], pre[:, 1]]
    keep_old = old_int > KDE_THRESH
    old = pre[keep_old]
    old_int = old_int[keep_old]
    old_grad = grad_kde[old[:, 0], old[:, 1]]
    cand = np.concatenate([old, pixel_idx], axis=0)
    iv = np.concatenate([old_int, new_int], axis=0)
    gv = np.concatenate([old_grad, new_grad], axis=0)
    scores = 1 / 3 * (iv * gv + iv + gv)
    Np, i = N_pre, 0
    best_xy_s, bin_idx, uniq = np.zeros((0, 3)), np.zeros(0, dtype=int), np.zeros(0, dtype=int)
    while (Np - N_pre < pixel_thresh) and (Np < algo_thresh):
        score_thresh *= [0.95, 1.0][int(i == 0)]
        mask = scores >= score_thresh
        bp = cand[mask].reshape(-1, 2)
        bs = scores[mask].reshape(-1, 1)
        best_xy_s = np.concatenate((bp[:, [1, 0]], bs), axis=1)
        bin_idx = np.round((best_xy_s[:, 0] - x_st) / delta_x).astype(int)
        uniq = np.unique(bin_idx)
        Np = uniq.shape[0]
        i += 1
        if i > max_decays:
            raise RuntimeError("compute_new_obs: score threshold decayed to zero without enough bins")
    fobs = np.zeros((Np, 2), dtype=np.int64)
    for k, b in enumerate(uniq):
        rows = best_xy_s[bin_idx == b].reshape(-1, 3)
        fobs[k] = rows[np.argmax(rows[:, -1]), :2]
    return fobs, score_thresh


def bin_best(kde, grad_kde, pre_fobs_yx, x_st, x_en, delta_x, fix_endpoints=True):
    """Collapsed form: per bin the max score and the FIRST candidate (old
    observations first, then row-major new pixels) attaining it. Returns dict bin -> (score,
    x, y). This is the quantity the CUDA select kernel produces."""
    pixel_idx = np.argwhere(kde > KDE_THRESH)
    if fix_endpoints:
        pixel_idx = pixel_idx[(pixel_idx[:, 1] > x_st) & (pixel_idx[:, 1] < x_en)]
    pre = np.asarray(pre_fobs_yx, dtype=np.int64).reshape(-1, 2)
    old = pre[kde[pre[:, 0], pre[:, 1]] > KDE_THRESH]
    cand = np.concatenate([old, pixel_idx], axis=0)
    iv = kde[cand[:, 0], cand[:, 1]]
    gv = grad_kde[cand[:, 0], cand[:, 1]]
    scores = 1 / 3 * (iv * gv + iv + gv)
    bins = np.round((cand[:, 1] - x_st) / delta_x).astype(int)
    out = {}
    for k in range(cand.shape[0]):
        b = int(bins[k])
        if b not in out or scores[k] > out[b][0]:
            out[b] = (scores[k], int(cand[k, 1]), int(cand[k, 0]))
    return out


def threshold_loop(best_scores, N_pre, pixel_thresh, algo_thresh, score_thresh, max_decays=4000):
    """The decay loop of gpet.py:589-609 on the per-bin maxima. Returns (mask over bins, thr)."""
    best_scores = np.asarray(best_scores, dtype=np.float64)
    Np, i = N_pre, 0
    mask = np.zeros(best_scores.shape, dtype=bool)
    while (Np - N_pre < pixel_thresh) and (Np < algo_thresh):
        score_thresh *= [0.95, 1.0][int(i == 0)]
        mask = best_scores >= score_thresh
        Np = int(mask.sum())
        i += 1
        if i > max_decays:
            raise RuntimeError("threshold loop: score threshold decayed to zero without enough bins")
    return mask, score_thresh


# --------------------------------------------------------------------------------------
# Final hyper-parameter fit (gpet.py:232-248, 263-266; sklearn_gpr.py:229-234, 254-295, 410-436,
# 475-607, 647-718)
# --------------------------------------------------------------------------------------
def _kernel_and_grad(theta, X, w, kind, nu):
    """K(theta) and dK/dtheta for (Constant * RBF|Matern) + WeightedWhite, theta = log[const,
    length_scale, noise] (sklearn Sum/Product composition order)."""
    const, ls, noise = np.exp(theta)
    X = X.reshape(-1, 1)
    m = X.shape[0]
    if kind == "RBF":
        d2 = pdist(X / ls, metric="sqeuclidean")
        k = squareform(np.exp(-0.5 * d2))
        np.fill_diagonal(k, 1)
        dk = k * squareform(d2)
    else:
        d = pdist(X / ls, metric="euclidean")
        D = squareform(d ** 2)
        if nu == 0.5:
            k = np.exp(-d)
        elif nu == 1.5:
            t = d * math.sqrt(3)
            k = (1.0 + t) * np.exp(-t)
        elif nu == 2.5:
            t = d * math.sqrt(5)
            k = (1.0 + t + t ** 2 / 3.0) * np.exp(-t)
        else:
            raise NotImplementedError(nu)
        k = squareform(k)
        np.fill_diagonal(k, 1)
        if nu == 0.5:
            den = np.sqrt(D)
            q = np.zeros_like(D)
            np.divide(D, den, out=q, where=den != 0)
            dk = k * q
        elif nu == 1.5:
            dk = 3 * D * np.exp(-np.sqrt(3 * D))
        else:
            tmp = np.sqrt(5 * D)
            dk = 5.0 / 3.0 * D * (tmp + 1) * np.exp(-tmp)
    K1 = np.full((m, m), const)
    Kww = noise * np.diag(w)
    K = K1 * k + Kww
    dK = np.dstack((K1[:, :, None] * k[:, :, None], dk[:, :, None] * K1[:, :, None], Kww[:, :, None]))
    return K, dK


def neg_lml_and_grad(theta, X, y, w, kind, nu, gp_alpha=1e-6):
    """sklearn_gpr.py:512-583 (sign flipped like obj_func :257-262)."""
    K, dK = _kernel_and_grad(theta, X, w, kind, nu)
    K[np.diag_indices_from(K)] += gp_alpha
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


FINAL_FIT_BOUNDS = np.log(np.array([[0.01, 1e3], [0.1, 100.0], [1e-18, 1.0]]))


def final_fit(X, y, w, x_grid, kind, nu, noise_y, seed, n_restarts=12, gp_alpha=1e-6):
    """Converged branch. Returns (y_mean float64[n], y_std float64[n], theta_opt)."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    y_m, y_s = np.mean(y), np.std(y)
    y = (y - y_m) / y_s
    X_m, X_s = np.mean(X), np.std(X)
    X = (X - X_m) / X_s
    # GPR normalize_y=False branch still centres AND scales (sklearn_gpr.py:229-234)
    tm, ts = np.mean(y), np.std(y)
    if ts < 10 * np.finfo(np.float64).eps:
        ts = 1.0
    yt = (y - tm) / ts
    bounds = FINAL_FIT_BOUNDS
    rng = np.random.RandomState(seed)
    theta0 = np.log(np.array([5.0, 5.0, float(noise_y)]))

    def run(t0):
        res = scipy.optimize.minimize(neg_lml_and_grad, t0, args=(X, yt, w, kind, nu, gp_alpha),
                                      method="L-BFGS-B", jac=True, bounds=bounds)
        return res.x, res.fun

    optima = [run(theta0)]
    for _ in range(n_restarts):
        optima.append(run(rng.uniform(bounds[:, 0], bounds[:, 1])))
    vals = [o[1] for o in optima]
    theta = optima[int(np.argmin(vals))][0]
    K, _ = _kernel_and_grad(theta, X, w, kind, nu)
    K[np.diag_indices_from(K)] += gp_alpha
    L = scipy.linalg.cholesky(K, lower=True, check_finite=False)
    a = scipy.linalg.cho_solve((L, True), yt, check_finite=False)
    const, ls, _ = np.exp(theta)
    xs = (np.asarray(x_grid) - X_m) / X_s
    Ks = const * unit_kernel(kind, nu, ls, xs, X)
    mu = ts * (Ks @ a) + tm
    V = scipy.linalg.solve_triangular(L, Ks.T, lower=True, check_finite=False)
    var = np.full(xs.shape[0], const) * np.ones(xs.shape[0])
    var -= np.einsum("ij,ji->i", V.T, V)
    var[var < 0] = 0.0
    sd = np.sqrt(var * ts ** 2)
    return y_s * mu + y_m, sd, theta


# --------------------------------------------------------------------------------------
# Whole trace (gpet.py:22-178, 768-908)
# --------------------------------------------------------------------------------------
class OracleTracer:
    """Stage-split restatement of GP_Edge_Tracing. `factor_fn(cov, it) -> A` lets a test inject
    the factor the GPU produced (default: canonical host SVD). `record` collects per-iteration
    intermediates."""

    def __init__(self, init, grad_img, kernel_options=(1, 3, 3), noise_y=1, obs=np.array([], dtype=np.int8),
                 N_samples=500, score_thresh=1, delta_x=20, keep_ratio=0.1, pixel_thresh=5, seed=42,
                 return_std=False, fix_endpoints=True, factor_fn=None, loop_costs=False):
        self.init = init[np.argsort(init[:, 0])].astype(int)
        self.x_st, self.x_en = int(init[0, 0]), int(init[-1, 0])
        self.grad_img = normalise(grad_img, (0, 1), np.float64)
        self.noise_y = noise_y
        self.N_samples = int(N_samples) if N_samples > 100 else 1000
        self.obs = np.asarray(obs).reshape(-1, 2).astype(np.int64)
        self.seed = seed
        self.keep_ratio = float(keep_ratio) if 0 < keep_ratio <= 1 else 0.1
        self.pixel_thresh = int(pixel_thresh) if pixel_thresh >= 2 else 2
        self.score_thresh = float(score_thresh) if 0 < score_thresh <= 1 else 1
        self.delta_x = int(delta_x) if delta_x > 3 else 2
        self.return_std = return_std
        self.fix_endpoints = fix_endpoints
        self.M, self.N = grad_img.shape
        self.x_grid = self.x_st + np.arange(self.x_en - self.x_st + 1).astype(int)
        self.edge_length = self.x_grid.shape[0]
        self.N_subints = int(self.edge_length // self.delta_x)
        self.N_keep = int(keep_ratio * N_samples)
        self.algo_thresh = self.N_subints - (self.pixel_thresh - 1)
        self.grad_kde = kde_of_gradient(self.grad_img)
        if type(kernel_options) == dict:
            self.sigma_f = kernel_options["sigma_f"]
            self.sigma_l = kernel_options["length_scale"]
            self.kernel_type = kernel_options["kernel"]
            self.kernel_nu = kernel_options["nu"] if kernel_options["kernel"] == "Matern" else 2.5
        else:
            k, so, lo = kernel_options
            self.kernel_type = ["RBF", "Matern"][int(k > 0)]
            self.kernel_nu = [2.5, 1.5][int(k > 1)]
            sf = [10, 8, 6, 4, 2, 1][so - 1] if (so >= 0) and (so <= 5) else 1
            self.sigma_f = self.M // sf
            sl = [1, 4 / 3, 2, 4, 10][lo - 1] if (lo >= 0) and (lo <= 4) else 10
            self.sigma_l = self.edge_length // sl
        self.alpha_init = np.array(self.init.shape[0] * [[0.5, 1e-7][int(fix_endpoints)]])
        self.factor_fn = factor_fn
        self.loop_costs = loop_costs
        self.record = []

    def iteration(self, pre_fobs, it):
        """One pass of the while-loop body (gpet.py:839-861). Returns new fobs (xy)."""
        X, y, w = assemble_training_set(self.init, pre_fobs, self.alpha_init)
        post = posterior(X, y, w, self.x_grid, self.kernel_type, self.kernel_nu, self.sigma_l,
                         self.sigma_f, self.noise_y)
        A = self.factor_fn(post["cov"], it) if self.factor_fn is not None else canonical_factor(post["cov"])
        Z = standard_normals(self.seed + it + 1, self.N_samples, self.edge_length)
        Y = sample_curves(Z, A, post["mean"], post["y_s"])
        costs = costs_loop(self.grad_img, Y, self.x_grid) if self.loop_costs else \
            costs_vectorised(self.grad_img, Y, self.x_grid)
        idx, best_costs = top_keep(costs, self.N_keep)
        kde = kde_of_curves(Y[:, idx], best_costs, self.x_grid, self.M, self.N)
        thr_in = self.score_thresh
        fobs, self.score_thresh = compute_new_obs(
            kde, self.grad_kde, pre_fobs[:, [1, 0]], self.x_st, self.x_en, self.delta_x,
            self.pixel_thresh, self.algo_thresh, self.score_thresh, self.fix_endpoints)
        self.record.append(dict(it=it, X=X, y=y, w=w, mean=post["mean"], y_s=post["y_s"], cov=post["cov"], A=A,
                                costs=costs, keep_idx=idx, kde=kde, thr_in=thr_in,
                                thr_out=self.score_thresh, fobs=fobs, samples=Y))
        return fobs

    def __call__(self):
        pre_fobs = self.obs
        it = 0
        while pre_fobs.shape[0] < self.algo_thresh:
            pre_fobs = self.iteration(pre_fobs, it)
            it += 1
        X, y, w = assemble_training_set(self.init, pre_fobs, self.alpha_init)
        y_mean, y_std, theta = final_fit(X, y, w, self.x_grid, self.kernel_type, self.kernel_nu,
                                         self.noise_y, self.seed + it)
        self.final = dict(X=X, y=y, w=w, theta=theta, y_mean=y_mean, y_std=y_std, n_iter=it, fobs=pre_fobs)
        cred = (y_mean - 1.96 * y_std, y_mean + 1.96 * y_std)
        edge = np.rint(np.stack([y_mean, self.x_grid.astype(np.float64)], axis=1)).astype(int)
        return (edge, cred) if self.return_std else edge
