"""gaussian_process_edge_trace_b200 - B200 (sm_100a) implementation of the hot path of
jaburke166/gaussian_process_edge_trace behind the reference's own API.

    from gaussian_process_edge_trace_b200 import gpet, gpet_utils
    kernel = gpet_utils.kernel_builder(size=(11, 5))
    grad = gpet_utils.comp_grad_img(img, kernel)
    edge_pred, credint = gpet.GP_Edge_Tracing(init, grad, kernel_options, ..., return_std=True)()

mirrors `from gp_edge_tracing import gpet, gpet_utils` (reference gp_edge_tracing/__init__.py).
"""
from . import gpet, gpet_utils  # noqa: F401
from .gpet import GP_Edge_Tracing  # noqa: F401
from .engine import TraceBatch  # noqa: F401

__all__ = ["gpet", "gpet_utils", "GP_Edge_Tracing", "TraceBatch"]
