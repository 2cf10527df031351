"""Developer tool (GPU box): times individual C-ABI stages on realistic cfg5 data, across tuning-knob variants."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
import numpy as np, torch
import bench
import __graft_entry__
__graft_entry__.build()
from gaussian_process_edge_trace_b200 import TraceBatch, gpet_utils, _gp_host
from gaussian_process_edge_trace_b200._cabi import call, ptr, load, query

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1250
ONLY = sys.argv[2] if len(sys.argv) > 2 else ""
lib = load()
imgs = np.empty((B, 500, 500)); inits = np.empty((B, 2, 2), dtype=np.int64)
for i in range(B):
    imgs[i], inits[i] = bench.make_image(i)
grad = gpet_utils.comp_grad_img(torch.from_numpy(imgs).cuda(), gpet_utils.kernel_builder((11, 5)), return_tensor=True)
tb = TraceBatch(inits, grad, **bench.TRACE_KW)
for _ in range(8):
    tb.step()
st = torch.cuda.current_stream().cuda_stream
n, S, M, N, Kp = tb.n, tb.N_samples, tb.M, tb.N, tb.N_keep
nb = min(B, tb.Bc)


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


res = {}
cost_ref = None
if ONLY == "score1":     # the default scoring kernel only (target of the ncu --set full capture)
    f = lambda: call("gpet_score_f64", ptr(tb.curve_buffer()), ptr(tb.gradT), None, nb, n, S, M, N, tb.x_st, ptr(tb.d_cost), st)
    ms = timeit(f, reps=5)
    print(f"score default: {ms:.3f} ms  {nb*S*(8*n+8)/ms/1e6:.0f} GB/s")
    sys.exit(0)
for stages, mb in (() if ONLY == "density" else ((0, 6), (4, 4), (4, 5), (4, 6), (8, 4), (8, 5), (8, 6))):
    for scan in (1, 0):
        lib.gpet_set_tuning(0, 128); lib.gpet_set_tuning(1, scan); lib.gpet_set_tuning(4, stages); lib.gpet_set_tuning(5, mb); lib.gpet_set_tuning(7, 1)
        f = lambda: call("gpet_score_f64", ptr(tb.curve_buffer()), ptr(tb.gradT), None, nb, n, S, M, N, tb.x_st, ptr(tb.d_cost), st)
        ms = timeit(f, reps=5)
        c = tb.d_cost[:nb].cpu().numpy()
        if cost_ref is None:
            cost_ref = c
        res[f"score stages={stages} minb={mb} scan={scan}"] = (round(ms, 3), f"{nb*S*(8*n+8)/ms/1e6:.0f} GB/s", f"maxrel {np.abs(c/cost_ref-1).max():.1e}")
for cpt, mb in (() if ONLY == "density" else ((2, 3), (2, 4), (3, 3))):
    for scan in (1, 0):
        lib.gpet_set_tuning(1, scan); lib.gpet_set_tuning(4, 4); lib.gpet_set_tuning(5, mb); lib.gpet_set_tuning(7, cpt)
        tb.d_cost.zero_()
        f = lambda: call("gpet_score_f64", ptr(tb.curve_buffer()), ptr(tb.gradT), None, nb, n, S, M, N, tb.x_st, ptr(tb.d_cost), st)
        ms = timeit(f, reps=5)
        c = tb.d_cost[:nb].cpu().numpy()
        res[f"score cpt={cpt} minb={mb} scan={scan}"] = (round(ms, 3), f"{nb*S*(8*n+8)/ms/1e6:.0f} GB/s", f"maxrel {np.abs(c/cost_ref-1).max():.1e}")
lib.gpet_set_tuning(7, 0)
lib.gpet_set_tuning(5, 4)
lib.gpet_set_tuning(4, 4)
lib.gpet_set_tuning(0, 128); lib.gpet_set_tuning(1, 1)
if ONLY == "lml":
    res = {}
if ONLY == "score":
    Yv = tb.curve_buffer()[:64].reshape(64, n, S)
    spread = (Yv.amax(dim=2) - Yv.amin(dim=2))          # per trace, per column: rows spanned by the S curves
    print("curve spread (rows) per column: median", float(spread.median()), "p90", float(spread.flatten().kthvalue(int(0.9 * spread.numel())).values), "max", float(spread.max()))
    for k, v in res.items():
        print(f"{k:28s} {v}")
    sys.exit(0)
for th in (512, 0):
    lib.gpet_set_tuning(2, th)
    f = lambda: call("gpet_sym_eig_f64", ptr(tb.d_Mr), B, tb.rp, ptr(tb.d_d), ptr(tb.d_Q), ptr(tb.d_sweeps), ptr(tb.d_eig_work), st)
    res[f"eig th={th}"] = (round(timeit(f), 3), f"sweeps {int(tb.d_sweeps.max())}")
lib.gpet_set_tuning(2, 0)
lib.gpet_set_tuning(6, 0)
res["sample tiles"] = (round(timeit(lambda: call("gpet_sample_f64", ptr(tb.d_Zt), ptr(tb.d_A), ptr(tb.d_mean), ptr(tb.d_ys), nb, tb.rp, n, S, ptr(tb.curve_buffer()), st)), 3),)
y_ref = tb.curve_buffer()[:2].clone()
lib.gpet_set_tuning(6, 1)
res["sample"] = (round(timeit(lambda: call("gpet_sample_f64", ptr(tb.d_Zt), ptr(tb.d_A), ptr(tb.d_mean), ptr(tb.d_ys), nb, tb.rp, n, S, ptr(tb.curve_buffer()), st)), 3),
                 f"{2.0*nb*S*n*tb.rp/1e9:.1f} GFLOP")
res["sample"] = res["sample"] + (f"max diff vs tile kernel {float((tb.curve_buffer()[:2] - y_ref).abs().max()):.1e}",)
res["posterior"] = (round(timeit(lambda: call("gpet_posterior_lowrank_f64", ptr(tb.d_xi), ptr(tb.d_y), ptr(tb.d_w), ptr(tb.d_m), tb.mmax, int(tb.d_m.max()), B, n, ptr(tb.d_sigma_f), float(tb.noise_y), 1e-6, ptr(tb.kd), ptr(tb.Ur), ptr(tb.lam), tb.rp, ptr(tb.d_mean), ptr(tb.d_ys), ptr(tb.d_Mr), ptr(tb.d_status), ptr(tb.d_post_work), st)), 3), f"m max {int(tb.d_m.max())}")
res["assemble"] = (round(timeit(lambda: call("gpet_factor_assemble_f64", ptr(tb.d_d), ptr(tb.d_Q), ptr(tb.Ur), ptr(tb.uw), B, tb.rp, n, ptr(tb.d_A), st)), 3),)
if tb.bands_width:
    res["density bands"] = (round(timeit(lambda: tb._density_bands(tb.curve_buffer(), tb.d_idx, tb.d_wts, nb, S, st)), 3),)
    bh = tb.d_bands[:nb].cpu().numpy()
    res["density bands"] += (f"mean band {float((bh[:, :, 1] - bh[:, :, 0]).mean()):.1f} of {M} rows",)
    res["select bands"] = (round(timeit(lambda: call("gpet_select_bands_f64", ptr(tb.d_dens), ptr(tb.d_dmm), ptr(tb.grad_kde), None, ptr(tb.d_bands), nb, M, N, ptr(tb.col_bin), ptr(tb.group_cols), tb.n_groups, ptr(tb.d_old), ptr(tb.d_nold), tb.max_old, tb.nb, ptr(tb.d_bscore), ptr(tb.d_bpos), st)), 3),)
gwork = torch.empty(query("gpet_density_workspace_bytes", nb, M, N, Kp), dtype=torch.uint8, device="cuda")
res["density"] = (round(timeit(lambda: call("gpet_density_f64", ptr(tb.curve_buffer()), ptr(tb.d_idx), ptr(tb.d_wts), nb, n, S, Kp, M, N, tb.x_st, ptr(tb.d_dens), ptr(tb.d_dmm), ptr(gwork), st)), 3),)
res["select"] = (round(timeit(lambda: call("gpet_select_f64", ptr(tb.d_dens), ptr(tb.d_dmm), ptr(tb.grad_kde), None, nb, M, N, ptr(tb.col_bin), ptr(tb.group_cols), tb.n_groups, ptr(tb.d_old), ptr(tb.d_nold), tb.max_old, tb.nb, ptr(tb.d_bscore), ptr(tb.d_bpos), st)), 3),)
del gwork
res["topk"] = (round(timeit(lambda: call("gpet_topk_f64", ptr(tb.d_cost), nb, S, Kp, ptr(tb.d_idx), ptr(tb.d_best), ptr(tb.d_wts), st)), 3),)
if ONLY == "density":
    for k, v in res.items():
        if "density" in k or "select" in k:
            print(f"{k:28s} {v}")
    sys.exit(0)
# LML on the converged training sets
tb.run_loop()
tx, ty, tw, tm = tb._training_sets()
mm = tb.mmax
Xs = np.zeros((B, mm)); yt = np.zeros((B, mm))
for b in range(B):
    k = int(tm[b]); X = tx[b, :k].astype(float); y = ty[b, :k]
    y = (y - y.mean()) / y.std(); X = (X - X.mean()) / X.std()
    Xs[b, :k] = X; yt[b, :k] = (y - y.mean()) / y.std()
dX, dy, dw = (torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (Xs, yt, tw))
dm = torch.from_numpy(tm).cuda()
dxc = torch.from_numpy(np.ascontiguousarray(tx.astype(np.int32))).cuda()
XC = True
R = 13; E = B * R
rng = np.random.RandomState(5)
th0 = np.stack([np.log([5.0, 5.0, 1.0])] + [rng.uniform(_gp_host.FINAL_BOUNDS[:, 0], _gp_host.FINAL_BOUNDS[:, 1]) for _ in range(R - 1)])
theta = torch.from_numpy(np.tile(th0, (B, 1))).cuda()
tr = torch.from_numpy(np.repeat(np.arange(B, dtype=np.int32), R)).cuda()
df = torch.empty(E, dtype=torch.float64, device="cuda"); dg = torch.empty((E, 3), dtype=torch.float64, device="cuda")
fref = None
print("host cores", os.cpu_count())
for th, XC in ((256, False), (0, False), (0, True)):
    lib.gpet_set_tuning(3, th)
    f = lambda: call("gpet_lml_f64", ptr(dX), ptr(dy), ptr(dw), ptr(dxc) if XC else None, ptr(dm), mm, ptr(tr), ptr(theta), E, 0, 1e-6, ptr(df), ptr(dg), st)
    ms = timeit(f, reps=2)
    fv = df.cpu().numpy()
    if fref is None:
        fref = fv
    gv = dg.cpu().numpy()
    if th == 256:
        gref = gv
    else:
        ef = np.abs(fv / fref - 1); eg = np.abs(gv - gref) / (np.abs(gref).max(axis=1, keepdims=True) + 1e-300)
        print("f err quantiles 50/90/99/100:", np.nanquantile(ef, [0.5, 0.9, 0.99, 1.0]))
        print("g err (rel to max comp) quantiles:", np.nanquantile(eg, [0.5, 0.9, 0.99, 1.0]))
        print("first start only: f", np.nanmax(ef[::R]), "g", np.nanmax(eg[::R]))
        worst = int(np.nanargmax(eg.max(axis=1)))
        print("worst g at eval", worst, "theta", theta[worst].cpu().numpy(), "f", fv[worst], fref[worst], "g", gv[worst], gref[worst])
    res[f"lml th={th} table={int(XC)}"] = (round(ms, 3), f"{E} evals, {ms*1e3/E:.2f} us/eval amortised, m~{int(tm.max())}", f"maxrel f {np.nanmax(np.abs(fv/fref-1)):.1e} g {np.nanmax(np.abs(gv-gref)/(np.abs(gref)+1e-6)):.1e}")
for k, v in res.items():
    print(f"{k:28s} {v}")
json.dump({k: list(map(str, v)) for k, v in res.items()}, open(os.path.join(ROOT, "gpurun_out", "kernels.json"), "w"), indent=1)
