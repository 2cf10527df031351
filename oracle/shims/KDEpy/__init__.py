"""TEST INFRASTRUCTURE ONLY. Stand-in for KDEpy.FFTKDE (package absent in this image;
"parity unpinned" at this one boundary - see oracle/README and DESIGN.md).

Restates the published KDEpy 1.1.x algorithm for `FFTKDE(kernel='gaussian', bw).fit(data,
weights).evaluate(grid)` as used at gpet.py:514-521:
  1. weights normalised to sum 1 (KDEpy `linear_binning`),
  2. linear binning of the points on the equidistant grid (multilinear "splat"; points are
     visited in input order, corners in binary-counter order),
  3. Gaussian kernel exp(-|x|^2/(2 bw^2)) / (2 pi bw^2)^(d/2) sampled on the grid out to
     L = floor(practical_support(bw)/dx) steps, practical_support = root of the 1-D pdf minus
     1e-4 (brentq, xtol=1e-3) + 1e-3  (= 4.07 for bw=1 -> L=4 -> 9x9 taps),
  4. scipy.signal.convolve(binned, kernel, mode='same').
"""
import numpy as np
from scipy.optimize import brentq
from scipy.signal import convolve


def _gauss_pdf_1d(x, bw):
    return np.exp(-0.5 * (x / bw) ** 2) / (np.sqrt(2 * np.pi) * bw)


def practical_support(bw, atol=10e-5):
    xtol = 1e-3
    return brentq(lambda x: _gauss_pdf_1d(x, bw) - atol, a=0, b=8 * bw, xtol=xtol, full_output=False) + xtol


def linear_binning(data, grid_axes, weights):
    """data (P,d); grid_axes: list of d sorted equidistant 1-D axes; returns array of shape
    (len(ax0), ..., len(ax_{d-1}))."""
    data = np.asarray(data, dtype=np.float64)
    P, d = data.shape
    weights = np.asarray(weights, dtype=np.float64)
    weights = weights / np.sum(weights)
    shape = tuple(len(a) for a in grid_axes)
    mins = np.array([a[0] for a in grid_axes], dtype=np.float64)
    dxs = np.array([(a[-1] - a[0]) / (len(a) - 1) for a in grid_axes], dtype=np.float64)
    t = (data - mins) / dxs
    integral = np.floor(t)
    frac = t - integral
    integral = integral.astype(np.int64)
    out = np.zeros(int(np.prod(shape)), dtype=np.float64)
    strides = np.array([int(np.prod(shape[i + 1:])) for i in range(d)], dtype=np.int64)
    idx_all = np.zeros((P, 2 ** d), dtype=np.int64)
    w_all = np.zeros((P, 2 ** d), dtype=np.float64)
    for corner in range(2 ** d):
        bits = [(corner >> (d - 1 - i)) & 1 for i in range(d)]
        fr = np.ones(P, dtype=np.float64)
        idx = np.zeros(P, dtype=np.int64)
        ok = np.ones(P, dtype=bool)
        for i, b in enumerate(bits):
            fr = fr * (frac[:, i] if b else (1.0 - frac[:, i]))
            ii = integral[:, i] + b
            ok &= (ii >= 0) & (ii < shape[i])
            idx += np.clip(ii, 0, shape[i] - 1) * strides[i]
        idx_all[:, corner] = idx
        w_all[:, corner] = np.where(ok, fr * weights, 0.0)
    # np.add.at is unbuffered: accumulation is sequential in (point, corner) order, like the
    # Cython loop of KDEpy (points outer, corners inner)
    np.add.at(out, idx_all.ravel(), w_all.ravel())
    return out.reshape(shape)


class FFTKDE:
    def __init__(self, kernel="gaussian", bw=1, norm=2):
        if kernel != "gaussian":
            raise NotImplementedError("stand-in implements the gaussian kernel only")
        self.bw = float(bw)

    def fit(self, data, weights=None):
        self.data = np.asarray(data, dtype=np.float64)
        if self.data.ndim == 1:
            self.data = self.data.reshape(-1, 1)
        self.weights = np.ones(self.data.shape[0]) if weights is None else np.asarray(weights, dtype=np.float64)
        return self

    def evaluate(self, grid_points):
        grid_points = np.asarray(grid_points, dtype=np.float64)
        d = grid_points.shape[1]
        axes = [np.unique(grid_points[:, i]) for i in range(d)]
        mn, mx = grid_points.min(axis=0), grid_points.max(axis=0)
        if not ((mn < self.data.min(axis=0)).all() and (mx > self.data.max(axis=0)).all()):
            raise ValueError("Every data point must be inside of the grid.")
        binned = linear_binning(self.data, axes, self.weights)
        num_intervals = np.array([len(a) - 1 for a in axes])
        dx = (mx - mn) / num_intervals
        real_bw = practical_support(self.bw)
        L = np.minimum(np.floor(real_bw / dx), num_intervals + 1)
        grids = [np.linspace(-dxi * Li, dxi * Li, int(Li * 2 + 1)) for dxi, Li in zip(dx, L)]
        mesh = np.stack(np.meshgrid(*grids, indexing="ij"), axis=-1)
        r2 = np.sum(mesh ** 2, axis=-1)
        kernel_weights = np.exp(-0.5 * r2 / self.bw ** 2) / ((2 * np.pi) ** (d / 2) * self.bw ** d)
        ans = convolve(binned, kernel_weights, mode="same")
        return ans.reshape(-1)
