"""TEST INFRASTRUCTURE ONLY - never imported by the product package.

Imports the UNMODIFIED reference (`/root/reference/gp_edge_tracing`) in this container so
that golden vectors can be generated from the reference's own code.
`/root/reference` does not exist on the GPU box: `oracle/make_golden.py` (run by hand in the build
container, output committed) uses it there; the reference arm of `bench.py` (`--impl reference`) imports
the copy that `__graft_entry__.build()` staged under `oracle/_ref/` (git-ignored, shipped by gpurun).

Compatibility shims (none of them touches reference files):
  * `oracle/shims/{matplotlib,skimage,KDEpy}` - packages absent from this image.
    matplotlib: stub. skimage.util.random_noise and KDEpy.FFTKDE: stand-ins that restate the
    third-party arithmetic ("parity unpinned" at those two boundaries).
  * scipy >= 1.14 dropped `scipy.integrate.simps` (gpet.py:404-405) -> alias to `simpson`
    (identical arithmetic for an odd number of samples, i.e. even edge_length).
  * sklearn >= 1.6 dropped `BaseEstimator._validate_data` (sklearn_gpr.py:211,361) -> adapter
    onto `sklearn.utils.validation.validate_data`.
  * numpy's `multivariate_normal` resolves `numpy.linalg.svd` at call time; LAPACK's singular
    vector signs depend on the BLAS thread count, so the factor is pinned
    with the project's canonical sign rule (`canonical_svd`), the same rule the CUDA factor
    provider applies. A caller may instead inject any factor through `set_factor_hook`.
"""
import os
import sys

import numpy as np

# the unmodified reference package: /root/reference in the build container; on the GPU box the copy that
# __graft_entry__.build() staged under oracle/_ref/ (git-ignored, shipped by gpurun)
_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.path.join(_HERE, "_ref") if os.path.isdir(os.path.join(_HERE, "_ref", "gp_edge_tracing")) \
    else "/root/reference"
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")

_raw_svd = np.linalg.svd
_factor_hook = None


def sign_weights(n):
    """Fixed generic weight vector of the canonical sign rule."""
    return 1.0 + np.arange(n, dtype=np.float64) / n


def canonical_svd(a, *args, **kw):
    """numpy.linalg.svd with every right singular vector flipped so that <Vt[k], w> > 0."""
    u, s, vt = _raw_svd(a, *args, **kw)
    sg = np.sign(vt @ sign_weights(vt.shape[1]))
    sg[sg == 0] = 1.0
    return u * sg, s, vt * sg[:, None]


def _hooked_svd(a, *args, **kw):
    if _factor_hook is not None:
        out = _factor_hook(np.asarray(a))
        if out is not None:
            return out
    return canonical_svd(a, *args, **kw)


def set_factor_hook(fn):
    """fn(cov) -> (u, s, vt) or None. Lets a test inject the factor the GPU produced."""
    global _factor_hook
    _factor_hook = fn


def load_reference():
    """Returns (gpet, gpet_utils, sklearn_gpr) modules of the unmodified reference."""
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f"{REFERENCE_ROOT} is not present (build container: /root/reference; elsewhere oracle/_ref)")
    for p in (_SHIMS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    import scipy.integrate

    if not hasattr(scipy.integrate, "simps"):
        scipy.integrate.simps = lambda y, x=None, dx=1.0, axis=-1, even=None: scipy.integrate.simpson(
            y, x=x, dx=dx, axis=axis
        )
    from sklearn.base import BaseEstimator

    if not hasattr(BaseEstimator, "_validate_data"):
        from sklearn.utils.validation import validate_data

        def _validate_data(self, X="no_validation", y="no_validation", reset=True, **kw):
            return validate_data(self, X=X, y=y, reset=reset, **kw)

        BaseEstimator._validate_data = _validate_data
    np.linalg.svd = _hooked_svd
    from gp_edge_tracing import gpet, gpet_utils, sklearn_gpr

    return gpet, gpet_utils, sklearn_gpr
