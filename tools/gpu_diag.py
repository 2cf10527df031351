"""Developer diagnostic (GPU box): per-stage error metrics of the CUDA path vs the oracle. Not a test."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
import numpy as np
import torch
import gpet_oracle as O
import __graft_entry__
__graft_entry__.build()
import gaussian_process_edge_trace_b200 as pkg

GOLDEN = os.path.join(ROOT, "tests", "golden")


def diag(name, init, grad, kw, factor="device"):
    t0 = time.time()
    tr = pkg.gpet.GP_Edge_Tracing(init, grad, record=True, factor=factor, **kw)
    t1 = time.time()
    edge, cred = tr()
    t2 = time.time()
    rec = tr.record
    orc = O.OracleTracer(init, grad, factor_fn=lambda cov, it: rec[it]["A"][0], **kw)
    gk, gko = tr.grad_kde, orc.grad_kde
    big = gko > 1e-6
    print(f"[{name}/{factor}] init {t1-t0:.2f}s call {t2-t1:.2f}s rank {tr._tb.rank} rp {tr._tb.rp} lowrank {tr._tb.lowrank}")
    print(f"  grad_img equal {np.array_equal(tr.grad_img, orc.grad_img)}; grad_kde: mismatches(>1e-6) {(gk[big] != gko[big]).sum()}/{big.sum()} maxabs {np.abs(gk-gko).max():.3e}")
    edge_o, cred_o = orc()
    print(f"  iterations gpu {len(rec)} oracle {len(orc.record)}")
    for r, o in zip(rec, orc.record):
        A = r["A"][0]; c = o["cov"].max()
        k_g = r["kde"][0].astype(np.float64); k_o = o["kde"]; big = k_o > 1e-6
        nm = (k_g[big] != k_o[big]).sum()
        rel = np.abs(k_g[big] / k_o[big] - 1).max() if big.any() else 0
        print(f"  it{r['it']:2d} m={o['X'].shape[0]:3d} AtA {np.abs(A.T@A-o['cov']).max()/c:.1e} mean {np.abs(r['mean'][0]-o['mean']).max():.1e} "
              f"Y {np.abs(r['samples'][0]-o['samples']).max():.1e} cost {np.abs(r['costs'][0]/o['costs']-1).max():.1e} "
              f"keep {np.array_equal(r['keep_idx'][0], o['keep_idx'])} kde_mis {nm}/{big.sum()} rel {rel:.1e} small {np.abs(k_g-k_o)[~big].max():.1e} "
              f"fobs {np.array_equal(r['fobs'][0], o['fobs'])} thr {r['thr_out'][0]==o['thr_out']}"
              + (f" sweeps {r['sweeps'][0]}" if 'sweeps' in r else ""))
        if nm:
            idx = np.argwhere(big & (k_g != k_o))[:4]
            for y, x in idx:
                print(f"      kde[{y},{x}] gpu {k_g[y,x]!r} oracle {k_o[y,x]!r}")
    print(f"  edge equal {np.array_equal(edge, edge_o)} cred maxdiff {np.abs(cred[0]-cred_o[0]).max():.2e}")
    if hasattr(tr, "final_info"):
        fi = tr.final_info
        print(f"  final fit: theta gpu {fi['theta'][0]} oracle {orc.final['theta']} rounds {fi['rounds']} lml_evals {fi['lml_evals']} "
              f"mean maxdiff {np.abs(fi['y_mean'][0]-orc.final['y_mean']).max():.2e}")
    return tr, orc


def small_case(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    kopt = {"kernel": str(g["kernel"]), "sigma_f": float(g["sigma_f"]), "length_scale": float(g["length_scale"]), "nu": float(g["nu"])}
    kw = dict(kernel_options=kopt, noise_y=1, N_samples=int(g["S"]), score_thresh=1, delta_x=int(g["delta_x"]),
              keep_ratio=0.25, pixel_thresh=3, seed=5, return_std=True, fix_endpoints=bool(g["fix_endpoints"]))
    return g, kw


if __name__ == "__main__":
    which = sys.argv[1:] or ["small", "cfg1"]
    if "small" in which:
        for nm in ("trace_small_rbf", "trace_small_matern", "trace_small_tuple_free"):
            g, kw = small_case(nm)
            for f in ("device", "host_svd"):
                try:
                    diag(nm, g["init"], g["grad"], kw, f)
                except Exception as e:
                    import traceback; traceback.print_exc()
    if "cfg1" in which:
        img, edge = O.construct_test_img((500, 500), 200, 4, 0.05, "sinusoidal", 0.3, gaps=True)
        grad = O.comp_grad_img(img, O.kernel_builder((11, 5)))
        init = edge[[0, -1], :][:, [1, 0]]
        kw = dict(kernel_options={"kernel": "RBF", "sigma_f": 75, "length_scale": 20}, noise_y=1, N_samples=1000,
                  score_thresh=1, delta_x=5, keep_ratio=0.1, pixel_thresh=5, seed=1, return_std=True, fix_endpoints=True)
        try:
            tr, orc = diag("cfg1", init, grad, kw, "device")
            g = np.load(os.path.join(GOLDEN, "trace_cfg1.npz"))
            same = [np.array_equal(r["fobs"][0], g[f"it{i}_fobs"]) for i, r in enumerate(tr.record) if i < int(g["n_iter"])]
            print("  vs golden (reference, pinned host SVD): fobs identical per iteration:", same)
        except Exception as e:
            import traceback; traceback.print_exc()
