import os
import sys

# Pin the host BLAS thread count before numpy loads: LAPACK's singular-vector signs (and the
# numerically-null subspace) depend on it.
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
os.environ.setdefault("OMP_NUM_THREADS", "1")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# A pytest plugin may have imported numpy before this file ran (the env vars above are then too
# late), so also pin the already-loaded BLAS at run time.
try:
    import threadpoolctl

    _BLAS_LIMIT = threadpoolctl.threadpool_limits(limits=1)
except Exception:  # pragma: no cover
    _BLAS_LIMIT = None


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
