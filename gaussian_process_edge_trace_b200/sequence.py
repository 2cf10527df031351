"""Image sequences (BASELINE.json config 4): every frame is traced with the previous frame's trace as its prior.

The reference has no helper for this; it offers the mechanism: the `obs` argument of GP_Edge_Tracing
(gpet.py:57-61, 100) seeds the observation set of the first iteration (gpet.py:820), so a caller feeds a thinned
copy of the previous edge_pred. This module is that caller, batched: B independent sequences advance frame by frame in
lock step (frames are sequential by construction - frame t needs the trace of frame t-1 - the parallelism is across
sequences and, inside a frame, across posterior samples).
"""
import numpy as np

from .engine import TraceBatch


def prior_from_trace(edge_pred, stride):
    """Observations (xy) handed to the next frame: every `stride`-th pixel of an edge_pred int[n, 2] (y, x), end points
    excluded (they are the next frame's `init`).  SURVEY.md 8(d) cfg 4: edge_pred[::4*delta_x][1:-1][:, [1, 0]]."""
    return np.asarray(edge_pred)[::stride][1:-1][:, [1, 0]].astype(np.int64)


def trace_sequence(frames, init, prior_stride=None, comp_grad=None, on_frame=None, **kw):
    """frames: iterable of gradient images float[B, M, N] (or [M, N] for one sequence), or of raw images when `comp_grad`
    (a callable image batch -> gradient batch, e.g. lambda x: gpet_utils.comp_grad_img(x, kernel, return_tensor=True)) is
    given.  init: int[B, 2, 2] (or [2, 2]) end points (x, y) of frame 0; later frames start from the end points of the
    previous trace.  kw: the GP_Edge_Tracing / TraceBatch options (kernel_options, N_samples, delta_x, ..., seed).
    Returns (edges int[T, B, n, 2], creds list over frames of per-trace (lo, hi), iterations int[T, B])."""
    init = np.asarray(init)
    single = init.ndim == 2
    if single:
        init = init[None]
    delta_x = int(kw.get("delta_x", 20))
    delta_x = delta_x if delta_x > 3 else 2                                   # gpet.py:105
    stride = int(prior_stride) if prior_stride is not None else 4 * delta_x
    edges_all, creds_all, iters_all = [], [], []
    obs = None
    for t, fr in enumerate(frames):
        g = comp_grad(fr) if comp_grad is not None else fr
        if not hasattr(g, "shape") or len(g.shape) == 2:
            g = g[None]
        tb = TraceBatch(init, g, obs=obs, **kw)
        edges, creds = tb.trace()
        edges_all.append(edges)
        creds_all.append(creds)
        iters_all.append(tb.n_iter.copy())
        if on_frame is not None:
            on_frame(t, tb, edges, creds)
        obs = [prior_from_trace(e, stride) for e in edges]
        init = np.stack([e[[0, -1]][:, [1, 0]] for e in edges]).astype(np.int64)
    return np.stack(edges_all), creds_all, np.stack(iters_all)
