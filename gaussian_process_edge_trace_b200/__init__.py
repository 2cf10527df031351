"""gaussian_process_edge_trace_b200 - B200 (sm_100a) implementation of the hot path of
jaburke166/gaussian_process_edge_trace behind the reference's own API.

    from gaussian_process_edge_trace_b200 import gpet, gpet_utils
    kernel = gpet_utils.kernel_builder(size=(11, 5))
    grad = gpet_utils.comp_grad_img(img, kernel)
    edge_pred, credint = gpet.GP_Edge_Tracing(init, grad, kernel_options, ..., return_std=True)()

mirrors `from gp_edge_tracing import gpet, gpet_utils` (reference gp_edge_tracing/__init__.py).
Submodules are imported lazily so that the L-BFGS-B worker processes (which need numpy/scipy only) do not pay
for importing torch.
"""
import importlib

__all__ = ["gpet", "gpet_utils", "engine", "sequence", "GP_Edge_Tracing", "TraceBatch"]
_LAZY = {"GP_Edge_Tracing": ("gpet", "GP_Edge_Tracing"), "TraceBatch": ("engine", "TraceBatch")}


def __getattr__(name):
    if name in ("gpet", "gpet_utils", "engine", "sequence", "dist", "_cabi", "_gp_host", "_lbfgs_worker"):
        return importlib.import_module(f"{__name__}.{name}")
    if name in _LAZY:
        mod, attr = _LAZY[name]
        return getattr(importlib.import_module(f"{__name__}.{mod}"), attr)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
