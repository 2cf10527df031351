"""Drop-in for reference gp_edge_tracing/gpet.py: class GP_Edge_Tracing with the reference's 13-parameter
constructor (gpet.py:22-35) and __call__ (gpet.py:768-773). The per-iteration work runs on the B200 through
`engine.TraceBatch` (a batch of one); the stage methods of the reference (`fit_predict_GP`, `get_best_curves`,
`cost_funct`, `kernel_density_estimate`, `get_best_pixels`) are kept as thin views onto the same CUDA stages
so code written against the reference's internal seams keeps working.

Not provided (out of scope): matplotlib plots behind `print_final_diagnostics`,
`show_init_post`, `show_post_iter` (accepted, raise NotImplementedError when set) and the dead
`grad_interpolation`.
"""
import time as t

import numpy as np
import torch

from . import _gp_host
from ._cabi import GpetError, call, ptr, query
from .engine import TraceBatch, _stream, final_fit_group


class GP_Edge_Tracing(object):
    """Traces an individual edge in an image using Gaussian process regression (B200 hot path)."""

    def __init__(self, init, grad_img, kernel_options=(1, 3, 3), noise_y=1, obs=np.array([], dtype=np.int8),
                 N_samples=500, score_thresh=1, delta_x=20, keep_ratio=0.1, pixel_thresh=5, seed=42,
                 return_std=False, fix_endpoints=True, factor="device", device=None, record=False,
                 final_fit="device", device_rng="auto"):
        init = np.asarray(init)
        obs = np.asarray(obs).reshape(-1, 2).astype(np.int64)                 # gpet.py:100
        self._tb = TraceBatch(init[None], np.asarray(grad_img)[None], kernel_options=kernel_options, noise_y=noise_y,
                              obs=[obs], N_samples=N_samples, score_thresh=score_thresh, delta_x=delta_x,
                              keep_ratio=keep_ratio, pixel_thresh=pixel_thresh, seed=seed,
                              fix_endpoints=fix_endpoints, factor=factor, device=device, record=record,
                              final_fit=final_fit, device_rng=device_rng)
        tb = self._tb
        # public attributes of the reference object (gpet.py:95-119, 130-151, 161-162)
        self.init = tb.init[0]
        self.x_st, self.x_en = tb.x_st, tb.x_en
        self.noise_y = noise_y
        self.N_samples = tb.N_samples
        self.obs = obs
        self.seed = seed
        self.keep_ratio = tb.keep_ratio
        self.pixel_thresh = tb.pixel_thresh
        self.delta_x = tb.delta_x
        self.half_delta = self.delta_x // 2
        self.return_std = return_std
        self.fix_endpoints = fix_endpoints
        self.kde_thresh = _gp_host.KDE_THRESH
        self.N_inits = tb.N_inits
        self.M, self.N = tb.M, tb.N
        self.x_grid = tb.x_grid
        self.X = np.repeat(self.x_grid.reshape(-1, 1), self.N_samples, axis=-1)
        self.edge_length = tb.n
        self.N_subints = tb.N_subints
        self.N_keep = tb.N_keep
        self.algo_thresh = tb.algo_thresh
        self.sigma_f, self.sigma_l = tb.sigma_f, tb.sigma_l
        self.kernel_type, self.kernel_nu = tb.ktype, tb.nu
        self.alpha_init = tb.alpha_init

    # ---- views of device state -------------------------------------------------------------------------
    @property
    def score_thresh(self):
        return float(self._tb.score_thresh[0])

    @score_thresh.setter
    def score_thresh(self, v):
        self._tb.set_score_thresh(0, v)

    @property
    def grad_img(self):
        """float64 view of the float32-normalised gradient image (gpet.py:97)."""
        return self._tb.grad[0].cpu().numpy().astype(np.float64)

    @property
    def grad_kde(self):
        """KDE of the gradient image (gpet.py:127)."""
        return self._tb.grad_kde[0].cpu().numpy().astype(np.float64)

    @property
    def record(self):
        return self._tb.record

    # ---- stage seams of the reference ------------------------------------------------------------------------
    def fit_predict_GP(self, obs, converged=False, seed=0):
        """gpet.py:182-268. converged=False: float64[n, N_samples] posterior curves drawn with RandomState(seed);
        converged=True: (y_mean, y_std) after the hyper-parameter fit."""
        tb = self._tb
        obs = np.asarray(obs).reshape(-1, 2).astype(np.int64)
        saved = tb.fobs[0]
        if converged:
            # the converged branch on the device as well (L-BFGS-B state machines + objective kernel); final_fit="host"
            # keeps scipy's own optimiser as a parity instrument
            if tb.final_fit_mode != "device":
                X, y, w = _gp_host.assemble_training_set(tb.init[0], obs, tb.alpha_init)
                y_mean, y_std, _ = _gp_host.final_fit(X.astype(np.float64), y, w, tb.x_grid, tb.ktype, tb.nu, tb.noise_y, seed)
                return y_mean, y_std
            tb.set_obs(0, obs)
            try:
                (_, _, info), = final_fit_group([tb], seed=seed)
                return info["y_mean"][0], info["y_std"][0]
            finally:
                tb.set_obs(0, saved)
        tb.set_obs(0, obs)
        try:
            A = self._posterior_and_factor()
            z = np.random.RandomState(seed).standard_normal((tb.N_samples, tb.n))
            zt = np.zeros((tb.rp, tb.N_samples))
            k = min(tb.rp, tb.n)
            zt[:k] = z[:, :k].T
            d_zt = torch.from_numpy(zt).to(tb.dev)
            call("gpet_sample_f64", ptr(d_zt), ptr(A), ptr(tb.d_mean), ptr(tb.d_ys), 1, tb.rp, tb.n, tb.N_samples,
                 ptr(tb.curve_buffer()), _stream())
            return tb.curve_buffer()[0].cpu().numpy()
        finally:
            tb.set_obs(0, saved)

    def _posterior_and_factor(self):
        tb = self._tb
        tb._ensure_device_state(all_traces=True)
        if tb.lowrank:
            st = _stream()
            call("gpet_posterior_lowrank_f64", ptr(tb.d_xi), ptr(tb.d_y), ptr(tb.d_w), ptr(tb.d_m), tb.mmax, 0, 1, tb.n,
                 ptr(tb.d_sigma_f), float(tb.noise_y), _gp_host.GP_ALPHA, ptr(tb.kd), ptr(tb.Ur), ptr(tb.lam), tb.rp,
                 ptr(tb.d_mean), ptr(tb.d_ys), ptr(tb.d_Mr), ptr(tb.d_status), ptr(tb.d_post_work), st)
            call("gpet_sym_eig_f64", ptr(tb.d_Mr), 1, tb.rp, ptr(tb.d_d), ptr(tb.d_Q), ptr(tb.d_sweeps), ptr(tb.d_eig_work), st)
            call("gpet_factor_assemble_f64", ptr(tb.d_d), ptr(tb.d_Q), ptr(tb.Ur), ptr(tb.uw), 1, tb.rp, tb.n,
                 ptr(tb.d_A), st)
            return tb.d_A
        return tb._factor_full(0, tb.B)

    def _upload_curves(self, y_samples):
        tb = self._tb
        y = np.ascontiguousarray(np.asarray(y_samples, dtype=np.float64))
        if y.shape != (tb.n, tb.N_samples):
            raise GpetError(f"y_samples must have shape {(tb.n, tb.N_samples)}")
        tb.curve_buffer()[0].copy_(torch.from_numpy(y))

    def cost_funct(self, edge):
        """gpet.py:371-410 for one curve given as xy rows (x must be the pixel grid x_st..x_en)."""
        tb = self._tb
        edge = np.asarray(edge, dtype=np.float64)
        edge = edge[edge[:, 0].argsort(), :]
        if edge.shape[0] != tb.n or not np.array_equal(edge[:, 0], tb.x_grid):
            raise GpetError("cost_funct: the curve must be sampled on the pixel grid x_st..x_en")
        Y = torch.from_numpy(np.ascontiguousarray(edge[:, 1:2])).to(tb.dev)
        cost = torch.empty((1, 1), dtype=torch.float64, device=tb.dev)
        call("gpet_score_f64", ptr(Y), ptr(tb.gradT), None, 1, tb.n, 1, tb.M, tb.N, tb.x_st, ptr(cost), _stream())
        return np.float64(cost.item())

    def get_best_curves(self, y_samples):
        """gpet.py:414-451 -> (best_curves[n, N_keep, 2], best_costs[N_keep], (optimal_curve, optimal_cost))."""
        tb = self._tb
        self._upload_curves(y_samples)
        st = _stream()
        call("gpet_score_f64", ptr(tb.curve_buffer()), ptr(tb.gradT), None, 1, tb.n, tb.N_samples, tb.M, tb.N, tb.x_st, ptr(tb.d_cost), st)
        call("gpet_topk_f64", ptr(tb.d_cost), 1, tb.N_samples, tb.N_keep, ptr(tb.d_idx), ptr(tb.d_best), ptr(tb.d_wts), st)
        idx = tb.d_idx[0].cpu().numpy()
        best_costs = tb.d_best[0].cpu().numpy()
        curves = np.stack((self.X, np.asarray(y_samples)), axis=-1)
        best_curves = curves[:, idx, :]
        return best_curves, best_costs, (best_curves[:, 0, :], best_costs[0])

    def kernel_density_estimate(self, best_curves, costs, bw=1):
        """gpet.py:455-529. (None, None) returns the KDE of the gradient image."""
        tb = self._tb
        if costs is None:
            return self.grad_kde
        if bw != 1:
            raise NotImplementedError("only bw=1 (the value the reference uses) is implemented")
        best_curves = np.asarray(best_curves, dtype=np.float64)
        Kp = best_curves.shape[1]
        Y = torch.from_numpy(np.ascontiguousarray(best_curves[:, :, 1])).to(tb.dev)          # [n, Kp]
        idx = torch.arange(Kp, dtype=torch.int32, device=tb.dev)
        inv = 1 / np.asarray(costs, dtype=np.float64)
        wts = torch.from_numpy(inv / np.sum(inv)).to(tb.dev)
        st = _stream()
        kde = torch.empty((1, tb.M, tb.N), dtype=torch.float32, device=tb.dev)
        if tb.bands_width:      # band-limited form (what the loop runs): workspace for this call's number of curves
            work = torch.empty(query("gpet_density_bands_workspace_bytes", 1, tb.n, Kp), dtype=torch.uint8, device=tb.dev)
            call("gpet_density_bands_f64", ptr(Y), ptr(idx), ptr(wts), 1, tb.n, Kp, Kp, tb.M, tb.N, tb.x_st,
                 ptr(tb.group_cols), tb.n_groups, tb.bands_width, ptr(tb.d_dens), ptr(tb.d_dmm), ptr(tb.d_bands), ptr(work), st)
            call("gpet_kde_bands_f32", ptr(tb.d_dens), ptr(tb.d_dmm), ptr(tb.d_bands), ptr(tb.group_cols), tb.n_groups, 1,
                 tb.M, tb.N, ptr(kde), st)
        else:
            work = torch.empty(query("gpet_density_workspace_bytes", 1, tb.M, tb.N, Kp), dtype=torch.uint8, device=tb.dev)
            call("gpet_density_f64", ptr(Y), ptr(idx), ptr(wts), 1, tb.n, Kp, Kp, tb.M, tb.N, tb.x_st, ptr(tb.d_dens),
                 ptr(tb.d_dmm), ptr(work), st)
            call("gpet_kde_normalised_f32", ptr(tb.d_dens), ptr(tb.d_dmm), 1, tb.M, tb.N, ptr(kde), st)
        return kde[0].cpu().numpy().astype(np.float64)

    def get_best_pixels(self, best_curves, costs, pre_fobs):
        """gpet.py:622-662 (`pre_fobs` in yx order, as the reference passes it). Returns fobs int64[k, 2] in xy
        order and updates `score_thresh` like compute_new_obs does."""
        tb = self._tb
        self.kernel_density_estimate(best_curves, costs)          # leaves dens/minmax of trace 0 on the device
        pre = np.asarray(pre_fobs).reshape(-1, 2).astype(np.int64)
        if pre.shape[0] > tb.max_old:
            raise GpetError("too many previous observations")
        old = torch.zeros((1, tb.max_old, 2), dtype=torch.int32)
        old[0, : pre.shape[0]] = torch.from_numpy(pre.astype(np.int32))
        nold = torch.tensor([pre.shape[0]], dtype=torch.int32)
        tb.d_old[:1].copy_(old)
        tb.d_nold[:1].copy_(nold)
        if tb.bands_width:
            call("gpet_select_bands_f64", ptr(tb.d_dens), ptr(tb.d_dmm), ptr(tb.grad_kde), None, ptr(tb.d_bands), 1, tb.M,
                 tb.N, ptr(tb.col_bin), ptr(tb.group_cols), tb.n_groups, ptr(tb.d_old), ptr(tb.d_nold), tb.max_old, tb.nb,
                 ptr(tb.d_bscore), ptr(tb.d_bpos), _stream())
        else:
            call("gpet_select_f64", ptr(tb.d_dens), ptr(tb.d_dmm), ptr(tb.grad_kde), None, 1, tb.M, tb.N, ptr(tb.col_bin),
                 ptr(tb.group_cols), tb.n_groups, ptr(tb.d_old), ptr(tb.d_nold), tb.max_old, tb.nb, ptr(tb.d_bscore),
                 ptr(tb.d_bpos), _stream())
        best = tb.d_bscore[:1].cpu().numpy()
        pos = tb.d_bpos[0].cpu().numpy().astype(np.int64)
        thr = tb.score_thresh[:1].copy()
        mask = _gp_host.threshold_loop_batch(best, np.array([pre.shape[0]]), tb.pixel_thresh, tb.algo_thresh, thr,
                                             np.array([True]))
        tb.set_score_thresh(0, thr[0])                                          # gpet.py:595 decays it in place
        sel = np.flatnonzero(mask[0])
        p = pos[sel]
        fobs = np.empty((sel.shape[0], 2), dtype=np.int64)
        is_old = p < tb.max_old
        fobs[is_old] = pre[p[is_old]][:, [1, 0]]
        q = p[~is_old] - tb.max_old
        fobs[~is_old, 0] = q % tb.N
        fobs[~is_old, 1] = q // tb.N
        return fobs

    # ---- the algorithm ---------------------------------------------------------------------------------------
    def __call__(self, print_final_diagnostics=False, show_init_post=False, show_post_iter=False, verbose=False,
                 return_lines=False):
        """gpet.py:768-908. Returns edge_trace int[n, 2] (y, x) [, (lo, hi) 95% credible interval]."""
        if print_final_diagnostics or show_init_post or show_post_iter:
            raise NotImplementedError("matplotlib diagnostics are outside the B200 hot path")
        tb = self._tb
        alg_st = t.time()
        all_obs = [self.obs]
        all_samples = []
        iter_optimal_curves = []
        keep = tb.record is not None or return_lines
        if return_lines and tb.record is None:
            tb.record = []
        while True:
            st = t.time()
            if verbose:
                print('Fitting Gaussian process and computing next set of observations...')
            if not tb.step():
                break
            all_obs.append(tb.fobs[0])
            if keep:
                rec = tb.record[-1]
                all_samples.append(rec["samples"][0])
                opt = rec["samples"][0][:, rec["keep_idx"][0, 0]]
                iter_optimal_curves.append(np.stack([self.x_grid.astype(np.float64), opt], axis=1))
            if verbose:
                print(f'Number of observations: {tb.fobs[0].shape[0]}')
                print(f'Iteration {int(tb.n_iter[0]) + 1} - Time Elapsed: {round(t.time() - st, 4)}\n\n')
        if tb.final_fit_mode == "device":
            edges, creds, info = tb.final_fit_all()
            edge_trace, cred_interval, y_mean = edges[0], creds[0], info["y_mean"][0]
            self.final_info = info
        else:
            edge_trace, cred_interval, (y_mean, _, _) = tb.final_fit(0)
        all_samples.append(y_mean)
        all_obs.append(tb.fobs[0])
        iter_optimal_curves.append(edge_trace[:, [1, 0]])
        if verbose:
            print(f'Time elapsed before algorithm converged: {round(t.time() - alg_st, 3)}')
        if self.return_std:
            return edge_trace, cred_interval
        if not return_lines:
            return edge_trace
        return edge_trace, (all_samples, all_obs, iter_optimal_curves)
