"""TEST INFRASTRUCTURE ONLY. Generates tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    OPENBLAS_NUM_THREADS=1 OMP_NUM_THREADS=1 python oracle/make_golden.py

The reference is imported through oracle/ref_harness.py (shims + pinned SVD signs); its
instance methods are wrapped (not edited) to record per-iteration intermediates. Dependency
versions used are stored in each file (`versions`). See tests/test_oracle_golden.py for how the
vectors are consumed.
"""
import hashlib
import os
import sys

os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
os.environ.setdefault("OMP_NUM_THREADS", "1")

import numpy as np  # noqa: E402
import scipy  # noqa: E402
import sklearn  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
VERSIONS = f"numpy {np.__version__}; scipy {scipy.__version__}; sklearn {sklearn.__version__}; " \
           f"OPENBLAS_NUM_THREADS={os.environ['OPENBLAS_NUM_THREADS']}"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def record_trace(gpet, tracer, full):
    """Wrap the stage methods of one reference instance and run it. Returns dict of arrays."""
    rec = {"iters": []}
    cur = {}
    factors = []

    orig_fit = tracer.fit_predict_GP
    orig_best = tracer.get_best_curves
    orig_kde = tracer.kernel_density_estimate
    orig_new = tracer.compute_new_obs

    def hook(cov):
        out = ref_harness.canonical_svd(cov)
        factors.append((np.array(cov), np.sqrt(out[1])[:, None] * out[2]))
        return out

    ref_harness.set_factor_hook(hook)

    def fit(obs, converged=False, seed=0):
        out = orig_fit(obs, converged=converged, seed=seed)
        if not converged:
            cur.clear()
            cur["obs_in"] = np.array(obs).reshape(-1, 2)
            cur["seed"] = seed
            cur["samples"] = np.array(out)
            cur["cov"], cur["A"] = factors[-1]
        else:
            rec["final_obs"] = np.array(obs).reshape(-1, 2)
            rec["final_seed"] = seed
            rec["final_mean"], rec["final_std"] = np.array(out[0]), np.array(out[1])
        return out

    def best(y_samples):
        # costs are recomputed through the reference's own cost_funct for recording
        out = orig_best(y_samples)
        curves = np.stack((tracer.X, y_samples), axis=-1)
        cur["costs"] = np.asarray([tracer.cost_funct(curves[:, i, :]) for i in range(tracer.N_samples)])
        cur["best_costs"] = np.array(out[1])
        cur["keep_idx"] = np.argsort(cur["costs"])[: tracer.N_keep]
        return out

    def kde(best_curves, costs, bw=1):
        out = orig_kde(best_curves, costs, bw)
        if costs is not None:
            cur["kde"] = np.array(out)
        return out

    def new_obs(pixel_idx, kde_arr, pre_fobs):
        cur["thr_in"] = tracer.score_thresh
        out = orig_new(pixel_idx, kde_arr, pre_fobs)
        cur["thr_out"] = tracer.score_thresh
        cur["fobs"] = np.array(out)
        cur["n_cand"] = pixel_idx.shape[0]
        rec["iters"].append(dict(cur))
        return out

    tracer.fit_predict_GP = fit
    tracer.get_best_curves = best
    tracer.kernel_density_estimate = kde
    tracer.compute_new_obs = new_obs
    edge, cred = tracer()
    ref_harness.set_factor_hook(None)
    rec["edge"] = edge
    rec["cred_lo"], rec["cred_hi"] = cred
    rec["grad_kde"] = np.array(tracer.grad_kde)
    rec["theta"] = None
    out = dict(versions=VERSIONS, n_iter=len(rec["iters"]), edge=rec["edge"], cred_lo=rec["cred_lo"],
               cred_hi=rec["cred_hi"], final_obs=rec["final_obs"], final_seed=rec["final_seed"],
               final_mean=rec["final_mean"], final_std=rec["final_std"],
               grad_kde_sha=sha(rec["grad_kde"]), grad_kde_probe=rec["grad_kde"][::7, ::7].copy())
    if full:
        out["grad_kde"] = rec["grad_kde"].astype(np.float32)
    for i, it in enumerate(rec["iters"]):
        p = f"it{i}_"
        out[p + "obs_in"] = it["obs_in"]
        out[p + "seed"] = it["seed"]
        out[p + "fobs"] = it["fobs"]
        out[p + "thr_in"] = it["thr_in"]
        out[p + "thr_out"] = it["thr_out"]
        out[p + "costs"] = it["costs"]
        out[p + "keep_idx"] = it["keep_idx"]
        out[p + "n_cand"] = it["n_cand"]
        out[p + "kde_sha"] = sha(it["kde"])
        out[p + "samples_sha"] = sha(it["samples"])
        out[p + "mean_est"] = it["samples"].mean(axis=1)
        if full:
            out[p + "cov"] = it["cov"]
            out[p + "A"] = it["A"]
            out[p + "samples"] = it["samples"]
            out[p + "kde"] = it["kde"].astype(np.float32)
            assert np.array_equal(out[p + "kde"].astype(np.float64), it["kde"])
    return out


def small_image(M, N, seed, amp, period, noise):
    """Deterministic small synthetic edge image (dark above / bright below a sinusoid)."""
    rng = np.random.default_rng(seed)
    cols = np.arange(N)
    rows = np.rint(M / 2 + amp * np.sin(2 * np.pi * cols / period)).astype(int)
    img = np.zeros((M, N))
    for j in range(N):
        img[rows[j]:, j] = 0.6
    img = np.clip(img + rng.normal(0, noise, img.shape), 0, 1)
    return img, rows


def main():
    gpet, gu, _ = ref_harness.load_reference()
    os.makedirs(OUT, exist_ok=True)

    # ---- kernel_builder / comp_grad_img / normalise --------------------------------------
    kb = {}
    for name, kw in {
        "k11x5": dict(size=(11, 5)),
        "k11x5_unit": dict(size=(11, 5), unit=True),
        "k7x3_b2d": dict(size=(7, 3), b2d=True),
        "k5x5_norm": dict(size=(5, 5), normalize=True),
        "k11x5_vert": dict(size=(11, 5), vertical_edges=True),
    }.items():
        kb[name] = gu.kernel_builder(**kw)
    rng = np.random.default_rng(7)
    img_a = rng.random((80, 96))
    img_b, _ = small_image(72, 64, 3, 14, 40, 0.05)
    st = dict(versions=VERSIONS, img_a=img_a, img_b=img_b)
    st["grad_a"] = gu.comp_grad_img(img_a, kb["k11x5"])
    st["grad_b"] = gu.comp_grad_img(img_b, kb["k11x5"])
    st["grad_a_k7x3"] = gu.comp_grad_img(img_a, kb["k7x3_b2d"])
    st["grad_b_unit"] = gu.comp_grad_img(img_b, kb["k11x5_unit"])
    st["norm_a_f64"] = gu.normalise(img_a * 3 - 1, (0, 1), np.float64)
    np.savez_compressed(os.path.join(OUT, "utils.npz"), **kb, **{"st_" + k: v for k, v in st.items()})

    # ---- README test image (cfg 1 input): checksums + the float32 gradient is regenerated by
    # the oracle at test time, only hashes are stored ------------------------------------
    N = 500
    img, edge = gu.construct_test_img(size=(N, N), amplitude=200, curvature=4, noise_level=0.05,
                                      ltype="sinusoidal", intensity=0.3, gaps=True)
    k = gu.kernel_builder(size=(11, 5), unit=False)
    g = gu.comp_grad_img(img, k)

    # ---- small traces, all intermediates stored -------------------------------------------
    for name, (M_, N_, kopt, dx, S, fix) in {
        "trace_small_rbf": (48, 64, {"kernel": "RBF", "sigma_f": 8, "length_scale": 8}, 4, 128, True),
        "trace_small_matern": (48, 64, {"kernel": "Matern", "nu": 2.5, "sigma_f": 8, "length_scale": 10}, 4, 128, True),
        "trace_small_tuple_free": (48, 64, (2, 2, 3), 6, 160, False),
    }.items():
        im, rows = small_image(M_, N_, 11, 9, 50, 0.04)
        gi = gu.comp_grad_img(im, k)
        init = np.array([[0, rows[0]], [N_ - 1, rows[-1]]])
        tr = gpet.GP_Edge_Tracing(init, gi, kernel_options=kopt, noise_y=1, obs=np.array([]), N_samples=S,
                                  score_thresh=1, delta_x=dx, keep_ratio=0.25, pixel_thresh=3, seed=5,
                                  return_std=True, fix_endpoints=fix)
        out = record_trace(gpet, tr, full=True)
        out.update(img=im, grad=gi, init=init, S=S, delta_x=dx, fix_endpoints=fix, keep_ratio=0.25,
                   pixel_thresh=3, seed=5, noise_y=1,
                   kernel=tr.kernel_type, nu=tr.kernel_nu, sigma_f=tr.sigma_f, length_scale=tr.sigma_l)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(name, "iters", out["n_iter"], "n_obs", out["final_obs"].shape[0])

    # ---- cfg 1: README recipe by keyword ----------------------------------------------------
    init = edge[[0, -1], :][:, [1, 0]]
    tr = gpet.GP_Edge_Tracing(init, g, kernel_options={"kernel": "RBF", "sigma_f": 75, "length_scale": 20},
                              noise_y=1, obs=np.array([]), N_samples=1000, score_thresh=1, delta_x=5,
                              keep_ratio=0.1, pixel_thresh=5, seed=1, return_std=True, fix_endpoints=True)
    out = record_trace(gpet, tr, full=False)
    out.update(img_sha=sha(img), grad_sha=sha(g), edge_true=edge, init=init,
               grad_probe=g[::50, ::50].copy(), img_probe=img[::50, ::50].copy())
    np.savez_compressed(os.path.join(OUT, "trace_cfg1.npz"), **out)
    print("cfg1 iters", out["n_iter"], "n_obs", out["final_obs"].shape[0])
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
