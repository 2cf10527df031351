"""Developer tool (GPU box): runs the BASELINE.json configurations that are not the bench workload once each and
prints wall time and trace quality (mean |edge_pred - true edge|).  cfg 1: README trace; cfg 2: N_samples = 100 000;
cfg 4 (shortened): frames of 1024 x 1024 with the previous trace as prior."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
import numpy as np, torch
import __graft_entry__
__graft_entry__.build()
from gaussian_process_edge_trace_b200 import gpet, gpet_utils

which = sys.argv[1:] or ["cfg1", "cfg2", "cfg4"]
kern = gpet_utils.kernel_builder((11, 5))
out = {}


def run(name, img, true_edge, obs=np.array([]), **kw):
    grad = gpet_utils.comp_grad_img(img, kern)
    init = true_edge[[0, -1], :][:, [1, 0]]
    torch.cuda.synchronize()
    t0 = time.time()
    tr = gpet.GP_Edge_Tracing(init, grad, obs=obs, return_std=True, **kw)
    edge, cred = tr()
    torch.cuda.synchronize()
    dt = time.time() - t0
    tb = tr._tb
    err = float(np.abs(edge[:, 0] - true_edge[:, 0]).mean())
    out[name] = dict(seconds=round(dt, 2), iterations=int(tb.n_iter[0]), observations=int(tb.n_obs[0]),
                     mean_abs_err_px=round(err, 2), large_m=bool(tb.large_m), lowrank=bool(tb.lowrank),
                     device_rng=bool(tb.device_rng), mmax=int(tb.mmax))
    print(name, out[name], flush=True)
    return edge


README = dict(kernel_options={"kernel": "RBF", "sigma_f": 75, "length_scale": 20}, noise_y=1, score_thresh=1, delta_x=5,
              keep_ratio=0.1, pixel_thresh=5, seed=1, fix_endpoints=True)
if "cfg1" in which or "cfg2" in which:
    img, edge = gpet_utils.construct_test_img((500, 500), 200, 4, 0.05, "sinusoidal", 0.3, gaps=True)
if "cfg1" in which:
    run("cfg1_S1000", img, edge, N_samples=1000, **README)
if "cfg2" in which:
    run("cfg2_S100000", img, edge, N_samples=100000, **README)
if "cfg4" in which:
    prev = None
    for t in range(int(os.environ.get("GPET_CFG4_FRAMES", "3"))):
        im, ed = gpet_utils.construct_test_img((1024, 1024), 300, 3, 0.05, "sinusoidal", 0.3, gaps=True)
        # frame t: phase-shifted copy (roll the columns) so the previous trace is a useful but imperfect prior
        sh = 16 * t
        im, ed2 = np.roll(im, sh, axis=1), ed.copy()
        ed2[:, 0] = np.roll(ed[:, 0], sh)
        obs = np.array([]) if prev is None else prev[::20][1:-1][:, [1, 0]]       # gpet.py:57-61
        prev = run(f"cfg4_frame{t}", im, ed2, obs=obs, N_samples=1000, **README)
if "cfg3" in which:
    # BASELINE config 3, scaled by GPET_CFG3_SIZE (4096 = the named size): E stacked dark->bright steps in
    # one image, every edge traced as one item of a TraceBatch, Matern nu = 2.5, delta_x = 2 (m up to size/2 + 2)
    from gaussian_process_edge_trace_b200 import TraceBatch
    size = int(os.environ.get("GPET_CFG3_SIZE", "2048"))
    E = int(os.environ.get("GPET_CFG3_EDGES", str(max(2, size // 256))))
    x = np.arange(size)
    img = np.zeros((size, size))
    rows = np.arange(size)[:, None]
    edges = []
    for e in range(E):
        ye = np.rint(0.23 * (size / E) * np.sin(4 * 2 * np.pi * x / (size - 1) + 0.7 * e)).astype(int) + (size // E) // 2 + (size // E) * e
        edges.append(ye)
        img += (rows >= ye[None, :]) / (E + 1.0)
    rng = np.random.default_rng(1)
    img = np.clip(img + rng.normal(0.0, np.sqrt(0.05) / (E + 1.0) / 0.3, img.shape), 0, 1)
    t0 = time.time()
    grad = gpet_utils.comp_grad_img(img, kern)
    inits = np.stack([np.array([[0, ye[0]], [size - 1, ye[-1]]]) for ye in edges])
    grads = np.broadcast_to(grad[None], (E, size, size))
    tb = TraceBatch(inits, np.ascontiguousarray(grads), kernel_options={"kernel": "Matern", "nu": 2.5, "sigma_f": 75,
                                                                        "length_scale": 20},
                    noise_y=1, N_samples=1000, score_thresh=1, delta_x=2, keep_ratio=0.1, pixel_thresh=5, seed=1,
                    fix_endpoints=True)
    e_pred, creds = tb.trace()
    torch.cuda.synchronize()
    err = [float(np.abs(e_pred[k][:, 0] - edges[k]).mean()) for k in range(E)]
    out["cfg3_scaled"] = dict(size=size, edges=E, seconds=round(time.time() - t0, 1), iterations=int(tb.n_iter.max()),
                              mmax=int(tb.mmax), observations=tb.n_obs.tolist(), large_m=bool(tb.large_m),
                              mean_abs_err_px=[round(v, 2) for v in err])
    print("cfg3_scaled", out["cfg3_scaled"], flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)
