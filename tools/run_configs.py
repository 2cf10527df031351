"""Developer tool (GPU box): runs the BASELINE.json configurations that are not the bench workload once each and
prints wall time and trace quality (mean |edge_pred - true edge|).  cfg 1: README trace; cfg 2: N_samples = 100 000;
cfg 4: frames of 1024 x 1024 with the previous trace as prior; cfg 3 (on request: `python tools/run_configs.py cfg3`,
GPET_CFG3_SIZE = 4096 for the named size): stacked edges, Matern nu = 2.5, delta_x = 2 (training sets up to size / 2 + 2).
JAC_BLOCK / JAC_INNER / JAC_EIG set the block Jacobi knobs for experiments."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
import numpy as np, torch
import __graft_entry__
__graft_entry__.build()
from gaussian_process_edge_trace_b200 import gpet, gpet_utils

if os.environ.get("JAC_BLOCK"):        # block Jacobi eigensolver: 32 (64 x 64 pivots) or 64 (128 x 128 pivots)
    from gaussian_process_edge_trace_b200._cabi import load as _load
    _load().gpet_set_tuning(10, int(os.environ["JAC_BLOCK"]))
if os.environ.get("JAC_INNER"):       # cap on the inner sweeps of a pivot solve of the block Jacobi eigensolver
    from gaussian_process_edge_trace_b200._cabi import load as _load
    _load().gpet_set_tuning(12, int(os.environ["JAC_INNER"]))
if os.environ.get("JAC_EIG"):          # pivots by the in-CTA parallel Jacobi kernel (experiment: also switches the low-rank path)
    from gaussian_process_edge_trace_b200._cabi import load as _load
    _load().gpet_set_tuning(2, int(os.environ["JAC_EIG"]))
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
    # BASELINE config 4: 64 frames of 1024 x 1024, frame t = sinusoid with phase 2 pi t / 640 and noise seed
    # t; frame 0 traced from its end points, frame t > 0 gets every 4*delta_x-th pixel of the previous edge_pred as obs
    from gaussian_process_edge_trace_b200 import sequence
    T = int(os.environ.get("GPET_CFG4_FRAMES", "64"))
    Nn = 1024
    xx = np.arange(Nn)

    def frame(t):
        wave = np.rint(150 * np.sin(3 * 2 * np.pi * xx / (Nn - 1) + 2 * np.pi * t / 640)).astype(int) + Nn // 2   # README-like: amplitude 300 // 2
        im = (np.arange(Nn)[:, None] >= wave[None, :]) * 0.3
        for a, b in ((20, 30), (Nn // 2, Nn // 2 + 10), (Nn - 100, Nn - 90), (Nn // 4, Nn // 4 + 20)):   # gaps
            im[:, a:b] = 0.0
        return gpet_utils.gaussian_noise(im, t + 1, 0.0, 0.05), wave

    frames, waves = zip(*[frame(t) for t in range(T)])
    init0 = np.array([[0, waves[0][0]], [Nn - 1, waves[0][-1]]])
    stamps = []
    torch.cuda.synchronize()
    t0 = time.time()
    edges, creds, iters = sequence.trace_sequence(
        frames, init0, comp_grad=lambda im: gpet_utils.comp_grad_img(im, kern, return_tensor=True), N_samples=1000,
        on_frame=lambda t, tb, e, c: stamps.append((time.time(), bool(tb.large_m), bool(tb.lowrank), int(tb.mmax), int(tb.rp))),
        **README)
    torch.cuda.synchronize()
    dt = time.time() - t0
    per = np.diff([t0] + [s_[0] for s_ in stamps])
    err = [float(np.abs(edges[t, 0, :, 0] - waves[t]).mean()) for t in range(T)]
    out["cfg4_sequence"] = dict(frames=T, size=Nn, seconds=round(dt, 2), first_frame_s=round(float(per[0]), 2),
                                later_frames_s_mean=round(float(per[1:].mean()), 3) if T > 1 else None,
                                iterations=iters[:, 0].tolist(), mean_abs_err_px=[round(v, 2) for v in err],
                                large_m=stamps[-1][1], lowrank=stamps[-1][2], mmax=stamps[-1][3], rp=stamps[-1][4])
    print("cfg4_sequence", out["cfg4_sequence"], flush=True)
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
                              jacobi_sweeps=getattr(tb, "jacobi_sweeps", None),
                              mean_abs_err_px=[round(v, 2) for v in err])
    print("cfg3_scaled", out["cfg3_scaled"], flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)
