"""Large training sets (m > GPET_MAX_TRAIN = 160, e.g. BASELINE configs 3/4): the posterior, the final-fit objective
and the final prediction as batched dense fp64 linear algebra on the device through torch (cuSOLVER / cuBLAS
library calls: Cholesky, triangular solves, DGEMM).  The hand-written shared-memory kernels (gpet_posterior.cu,
gpet_finalfit.cu) keep one training matrix per CTA and stop at m = 160; beyond that the matrices live in HBM/L2 and a
plain library factorisation is the right tool - documented in DESIGN.md as a library path, off the headline workload.

Same arithmetic as the kernels and the oracle: SURVEY.md A.4 / A.5, sklearn_gpr.py:221-227, 304-320, 379-407
(posterior), :475-585 (objective), :410-436 (final prediction).  Variable m per trace is handled by padding every
matrix to mmax with an identity block (Cholesky of the padding is 1, its log is 0, padded y / K* entries are 0).
"""
import numpy as np
import torch

from . import _gp_host

_SQ3, _SQ5 = 1.7320508075688772, 2.23606797749979


def posterior_full(x, y, w, m, x_st, n, sigma_f, noise_y, kd, dev):
    """x int64[B, mm] (sorted training columns, padded), y, w float64[B, mm], m int[B].
    Returns (mean [B, n], y_s [B], cov [B, n, n]) on `dev` (gpet.py:209-230, 253-261)."""
    B, mm = x.shape
    valid = np.arange(mm)[None, :] < m[:, None]
    ys = np.empty(B); ybar = np.empty(B); sy = np.empty(B)
    yn = np.zeros((B, mm))
    for k in np.unique(m):                                   # numpy-exact scalars, one group per training-set size
        rows = np.flatnonzero(m == k)
        yk = y[rows, :k]
        s = np.std(yk, axis=1) + 1.0                         # gpet.py:226
        yk = yk / s[:, None]
        mu = np.mean(yk, axis=1)                             # sklearn_gpr.py:221-227: the mean is removed, the std is
        sd = np.std(yk, axis=1)                              # kept but y is NOT divided by it (reference quirk)
        sd = np.where(sd < 10 * np.finfo(np.float64).eps, 1.0, sd)
        yn[rows, :k] = yk - mu[:, None]
        ys[rows], ybar[rows], sy[rows] = s, mu, sd
    c = (np.asarray(sigma_f, dtype=np.float64) ** 2 / ys ** 2)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    kd_t = t(kd)
    xi = t(x - x_st)
    vt = t(valid)
    ct = t(c)[:, None, None]
    dist = (xi[:, :, None] - xi[:, None, :]).abs().clamp_(max=kd_t.shape[0] - 1)
    pair = vt[:, :, None] & vt[:, None, :]
    K = torch.where(pair, ct * kd_t[dist], torch.zeros((), dtype=torch.float64, device=dev))
    # WeightedWhiteKernel drops the noise when the training set has exactly edge_length rows (sklearn_gpr.py:672-677)
    nz = t(np.where((m == n)[:, None], 0.0, noise_y * w))
    diag = torch.where(vt, (ct[:, :, 0] + nz) + _gp_host.GP_ALPHA, torch.ones((), dtype=torch.float64, device=dev))
    K.diagonal(dim1=1, dim2=2).copy_(diag)
    L, info = torch.linalg.cholesky_ex(K)
    if int(info.max()) != 0:
        bad = torch.nonzero(info).flatten().tolist()
        raise np.linalg.LinAlgError(f"Cholesky of the training kernel matrix failed for traces {bad} "
                                    "(sklearn_gpr.py:306-314)")
    ynt = t(yn)[:, :, None]
    alpha = torch.cholesky_solve(ynt, L)
    grid = torch.arange(n, device=dev)
    dq = (grid[None, :, None] - xi[:, None, :]).abs().clamp_(max=kd_t.shape[0] - 1)          # [B, n, mm]
    Ks = torch.where(vt[:, None, :], ct * kd_t[dq], torch.zeros((), dtype=torch.float64, device=dev))
    mean = t(sy)[:, None] * (Ks @ alpha)[:, :, 0] + t(ybar)[:, None]
    V = torch.linalg.solve_triangular(L, Ks.transpose(1, 2), upper=False)                   # [B, mm, n]
    dg = (grid[:, None] - grid[None, :]).abs()
    cov = (ct * kd_t[dg][None] - V.transpose(1, 2) @ V) * (t(sy) ** 2)[:, None, None]
    return mean, t(ys), cov


def _kernel_and_dlogl(kind, D):
    """Unit kernel and d k / d log(length_scale) from the squared scaled distance (sklearn kernels.py)."""
    if kind == 0:
        k = torch.exp(-0.5 * D)
        return k, k * D
    d = torch.sqrt(D)
    if kind == 1:
        k = torch.exp(-d)
        return k, k * d
    if kind == 2:
        tt = d * _SQ3
        e = torch.exp(-tt)
        return (1.0 + tt) * e, 3.0 * D * e
    tt = d * _SQ5
    e = torch.exp(-tt)
    return (1.0 + tt + tt * tt / 3.0) * e, 5.0 / 3.0 * D * (tt + 1.0) * e


class LargeFit:
    """Final-fit objective and prediction for padded training sets [B, mm] held on the device."""

    def __init__(self, Xs, yt, ws, ms, kind, dev, chunk_bytes=2 << 30):
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        self.X, self.y, self.w = t(Xs), t(yt), t(ws)
        self.m = t(ms.astype(np.int64))
        self.mm = Xs.shape[1]
        self.valid = torch.arange(self.mm, device=dev)[None, :] < self.m[:, None]
        self.kind, self.dev = kind, dev
        self.chunk = max(1, int(chunk_bytes // (6 * self.mm * self.mm * 8)))

    def _K(self, tr, theta):
        c, ls, noise = (torch.exp(theta[:, i]) for i in range(3))
        xs = self.X[tr] / ls[:, None]
        v = self.valid[tr]
        pair = v[:, :, None] & v[:, None, :]
        D = (xs[:, :, None] - xs[:, None, :]) ** 2
        k, dk = _kernel_and_dlogl(self.kind, D)
        zero = torch.zeros((), dtype=torch.float64, device=self.dev)
        eye = torch.eye(self.mm, dtype=torch.bool, device=self.dev)[None]
        k = torch.where(pair & ~eye, k, zero)
        dk = torch.where(pair & ~eye, dk, zero)
        K = c[:, None, None] * k
        dn = noise[:, None] * self.w[tr]
        diag = torch.where(v, (c[:, None] + dn) + _gp_host.GP_ALPHA, torch.ones((), dtype=torch.float64, device=self.dev))
        K.diagonal(dim1=1, dim2=2).copy_(diag)
        return K, k, dk, c, dn, v

    def objective(self, trace_of, theta):
        """-(log marginal likelihood), -(gradient) for E evaluations (sklearn_gpr.py:512-583, :257-262)."""
        E = theta.shape[0]
        f = np.empty(E); g = np.empty((E, 3))
        tr_all = torch.from_numpy(np.ascontiguousarray(trace_of.astype(np.int64))).to(self.dev)
        th_all = torch.from_numpy(np.ascontiguousarray(theta)).to(self.dev)
        for a in range(0, E, self.chunk):
            tr, th = tr_all[a:a + self.chunk], th_all[a:a + self.chunk]
            K, k, dk, c, dn, v = self._K(tr, th)
            L, info = torch.linalg.cholesky_ex(K)
            ok = info == 0
            L = torch.where(ok[:, None, None], L, torch.eye(self.mm, dtype=torch.float64, device=self.dev)[None])
            yv = self.y[tr][:, :, None]
            al = torch.cholesky_solve(yv, L)
            Kinv = torch.cholesky_inverse(L)
            mf = self.m[tr].to(torch.float64)
            nl = 0.5 * (yv * al).sum(dim=(1, 2)) + torch.log(L.diagonal(dim1=1, dim2=2)).sum(dim=1) \
                + 0.5 * mf * 1.8378770664093453
            Q = al @ al.transpose(1, 2) - Kinv
            qd = Q.diagonal(dim1=1, dim2=2)
            vd = v.to(torch.float64)
            g0 = (Q * (c[:, None, None] * k)).sum(dim=(1, 2)) + (qd * vd).sum(dim=1) * c
            g1 = (Q * (c[:, None, None] * dk)).sum(dim=(1, 2))
            g2 = (qd * dn * vd).sum(dim=1)
            gg = -0.5 * torch.stack([g0, g1, g2], dim=1)
            nl = torch.where(ok, nl, torch.full_like(nl, float("inf")))
            gg = torch.where(ok[:, None], gg, torch.zeros_like(gg))
            f[a:a + self.chunk] = nl.cpu().numpy()
            g[a:a + self.chunk] = gg.cpu().numpy()
        return f, g

    def predict(self, theta, xq, tm_ts):
        """Mean / std on the standardised grid at the optimum (sklearn_gpr.py:379-436). Returns (mean, sd, status)."""
        B = theta.shape[0]
        n = xq.shape[1]
        mean = np.empty((B, n)); sd = np.empty((B, n)); status = np.zeros(B, dtype=np.int32)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)
        for a in range(0, B, self.chunk):
            b = min(B, a + self.chunk)
            tr = torch.arange(a, b, device=self.dev)
            th = t(theta[a:b])
            K, _, _, c, _, v = self._K(tr, th)
            L, info = torch.linalg.cholesky_ex(K)
            status[a:b] = (info != 0).cpu().numpy()
            L = torch.where((info == 0)[:, None, None], L, torch.eye(self.mm, dtype=torch.float64, device=self.dev)[None])
            al = torch.cholesky_solve(self.y[a:b][:, :, None], L)
            ls = torch.exp(th[:, 1])
            D = (t(xq[a:b])[:, :, None] / ls[:, None, None] - (self.X[a:b] / ls[:, None])[:, None, :]) ** 2
            kx, _ = _kernel_and_dlogl(self.kind, D)
            Ks = torch.where(v[:, None, :], c[:, None, None] * kx, torch.zeros((), dtype=torch.float64, device=self.dev))
            tmts = t(tm_ts[a:b])
            mean[a:b] = (tmts[:, 1:2] * (Ks @ al)[:, :, 0] + tmts[:, 0:1]).cpu().numpy()
            V = torch.linalg.solve_triangular(L, Ks.transpose(1, 2), upper=False)
            var = (c[:, None] - (V * V).sum(dim=1)).clamp_min(0.0)
            sd[a:b] = torch.sqrt(var * tmts[:, 1:2] ** 2).cpu().numpy()
        return mean, sd, status
