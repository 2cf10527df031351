"""Lock-step batched edge tracing on one B200: the host control flow of GP_Edge_Tracing.__call__
(gpet.py:768-908) around the CUDA stages of libgpet_b200.so.

A `TraceBatch` holds B independent traces that share the image shape, the edge span [x_st, x_en], the GP
kernel and every scalar option (they may differ in image, endpoints' rows and prior observations). Per
iteration of the reference's while-loop (gpet.py:829-870) every unfinished trace goes through

    posterior (Cholesky, mean, reduced covariance)   gpet_posterior_lowrank_f64 / gpet_posterior_full_f64
    factor of the covariance                          gpet_sym_eig_f64 + gpet_factor_assemble_f64 (low rank) |
                                                      gpet_block_jacobi_* (full rank) | host SVD (parity mode)
    N_samples posterior curves                        gpet_sample_f64          (one DMMA contraction, shared Z)
    cost of every curve, top N_keep                   gpet_score_f64, gpet_topk_f64
    density of the kept curves                        gpet_density_f64
    per-bin best pixel                                gpet_select_f64
    threshold decay + new observation set             gpet_update_obs_f64, gpet_training_sets_f64

torch is used for device memory, streams and host<->device copies only.
"""
import contextlib
import threading

import numpy as np
import torch

import os
import time

from . import _cabi, _gp_host, _lbfgs_worker
from ._cabi import GpetError, call, ptr, query

MAX_TRAIN = 224     # GPET_MAX_TRAIN: largest training set of the shared-memory kernels (beyond: gpet_dense.cu)
MAX_RANK = 160      # GPET_MAX_RANK


def _stream():
    return torch.cuda.current_stream().cuda_stream


_pools = {}
_boost_streams = {}


def _boost_stream(dev, stage):
    """Side stream for the stages named in GPET_BOOST (default "score"; "" = none), created with a priority above the
    final-fit stream's (GPET_BOOST_PRIORITY, default -2; the fit stream has -1, the loop 0).  Measured on B200 (cfg 5
    shard, loop and fit streams overlapped): the scoring kernel shares the SMs with the objective kernel of the final fit
    otherwise and runs at 0.39 of the HBM peak; scheduled ahead of it, it streams its 5 GB of curves at 0.55 and the
    step gets 3 % shorter (298 against 307 ms).  Boosting the WHOLE loop starves the fit instead (329 ms)."""
    names = os.environ.get("GPET_BOOST", "score")
    if stage not in [t.strip() for t in names.split(",") if t.strip()]:
        return None
    key = str(dev)
    if key not in _boost_streams:
        _boost_streams[key] = torch.cuda.Stream(device=dev, priority=int(os.environ.get("GPET_BOOST_PRIORITY", "-2")))
    return _boost_streams[key]


def _host_cores():
    try:
        return max(2, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 2


def fit_pool(n_instances):
    """Process pool that drives the L-BFGS-B instances of the final fit when they run on the host (GPET_FIT_DRIVER=host
    only; the default driver is on the device and needs no pool). Shared by
    every TraceBatch of this process. Small problems run in-process. GPET_FIT_WORKERS overrides the worker count."""
    if n_instances <= 512:
        key = 0
    else:
        env = os.environ.get("GPET_FIT_WORKERS")
        if env is not None:
            key = max(0, int(env))
        else:
            world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1"))))
            key = max(1, min(32, _host_cores() // world - 1))
    if key not in _pools:
        _pools[key] = _lbfgs_worker.LbfgsbPool(key)
    return _pools[key]


_KIND = {("RBF", None): 0, ("Matern", 0.5): 1, ("Matern", 1.5): 2, ("Matern", 2.5): 3}


def device_standard_normal(seed, S, n, kcols, s0, S_loc, d_Zt, d_ok, d_fix, d_work, exact=True):
    """RandomState(seed).standard_normal((S, n)) on the device (gpet_standard_normal_t_f64): columns < kcols of the samples
    s0 .. s0 + S_loc - 1, transposed, into d_Zt[rows >= kcols][S_loc].  exact=True (needs d_fix): the attempts whose
    logarithm is too close to a rounding boundary for the device to know what glibc's log returns are recomputed on the
    host with libm's log and patched in - the result is numpy's array bit for bit (one host synchronisation)."""
    st = _stream()
    call("gpet_standard_normal_t_f64", seed & 0xffffffff, S, n, kcols, s0, S_loc, ptr(d_Zt), ptr(d_ok),
         ptr(d_fix) if exact else None, ptr(d_work), st)
    if not exact:
        return 0
    cnt = int(d_fix[:8].view(torch.int64).cpu()[0])
    cap = (d_fix.numel() - 16) // 40
    if cnt > cap:
        raise GpetError(f"device normal generator: {cnt} flagged attempts, room for {cap}")
    if cnt == 0:
        return 0
    key = str(d_Zt.device)
    pin = _rng_pinned.get(key)
    if pin is None or pin.numel() < 2 * cnt:
        pin = _rng_pinned[key] = torch.empty(2 * max(cnt, 1 << 16), dtype=torch.float64).pin_memory()
    f64view = d_fix[16:].view(torch.float64)                           # r2[cap] | lg[cap] | ...
    pin[:cnt].copy_(f64view[:cnt], non_blocking=True)
    torch.cuda.current_stream().synchronize()
    host = pin.numpy()
    call("gpet_host_log_f64", host[:cnt].ctypes.data, host[cnt:2 * cnt].ctypes.data, cnt)     # libm's log, as legacy_gauss
    f64view[cap:cap + cnt].copy_(pin[cnt:2 * cnt], non_blocking=True)
    call("gpet_standard_normal_fixup_apply_f64", S, n, kcols, s0, S_loc, ptr(d_Zt), ptr(d_fix), cnt, st)
    return cnt


_rng_pinned = {}


class StageTimers:
    """CUDA-event timers around the C-ABI stages (on torch's current stream, where the kernels are launched).
    `collect()` synchronises and returns {stage: (milliseconds, launches)} accumulated since the last reset."""

    def __init__(self):
        self.pending = []
        self.acc = {}

    def start(self):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def stop(self, name, e0):
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        self.pending.append((name, e0, e1))

    def collect(self):
        torch.cuda.synchronize()
        for name, e0, e1 in self.pending:
            ms, k = self.acc.get(name, (0.0, 0))
            self.acc[name] = (ms + e0.elapsed_time(e1), k + 1)
        self.pending = []
        return dict(self.acc)

    def reset(self):
        self.collect()
        self.acc = {}


class NormalDraws:
    """Z = RandomState(seed).standard_normal((S, n)) (numpy legacy polar method; sequential, so it stays on the
    host). Draws for consecutive seeds are produced ahead of use by a worker thread and only the
    first `rp` columns (those a rank-rp factor multiplies) are kept, transposed to [rp, S]. One generator is shared
    by every TraceBatch with the same (S, n, rp, seed) - sub-batches of a pipelined run ask for the same draws."""

    _shared = {}
    _shared_lock = threading.Lock()

    @classmethod
    def shared(cls, S, n, rp, base_seed):
        key = (S, n, rp, base_seed)
        with cls._shared_lock:
            if key not in cls._shared:
                if len(cls._shared) > 8:
                    cls._shared.clear()
                cls._shared[key] = cls(S, n, rp, base_seed)
            return cls._shared[key]

    def __init__(self, S, n, rp, base_seed, lookahead=4, keep=32):
        self.S, self.n, self.rp, self.base = S, n, rp, base_seed
        self.cache = {}
        self.lock = threading.Lock()
        self.lookahead = lookahead
        self.keep = keep
        self.thread = None
        self._want = 0
        self.dcache, self.dbytes, self.dcache_limit = {}, 0, 2 << 30

    def _make(self, it):
        z = np.random.RandomState(self.base + it + 1).standard_normal((self.S, self.n))   # gpet.py:839
        return np.ascontiguousarray(z[:, : self.rp].T)

    def _work(self):
        while True:
            with self.lock:
                todo = [k for k in range(self._want, self._want + self.lookahead) if k not in self.cache]
            if not todo:
                return
            zt = self._make(todo[0])
            with self.lock:
                self.cache[todo[0]] = zt

    def get(self, it):
        with self.lock:
            zt = self.cache.get(it)
            self._want = max(self._want, it + 1) if zt is not None or it >= self._want else self._want
            for k in [k for k in self.cache if k < it - self.keep]:
                del self.cache[k]
        if zt is None:
            zt = self._make(it)
            with self.lock:
                self.cache[it] = zt
                self._want = max(self._want, it + 1)
        with self.lock:
            if self.thread is None or not self.thread.is_alive():
                self.thread = threading.Thread(target=self._work, daemon=True)
                self.thread.start()
        return zt

    def device(self, it, dev, s0, S_loc, rows):
        """The draws of iteration `it` on the device: float64[rows, S_loc] = columns s0 .. s0+S_loc-1 of get(it), zero
        padded to `rows`. Uploaded once per (iteration, device, sample block) and shared by every TraceBatch - Z depends on
        the seed only, so the sub-batches and steps of a workload traced with one seed share it like the traces of a
        batch do. Returns (tensor, event recorded after the upload on the stream that issued it)."""
        key = (it, str(dev), s0, S_loc, rows)
        with self.lock:
            hit = self.dcache.get(key)
        if hit is not None:
            return hit
        zt = self.get(it)
        h = np.zeros((rows, S_loc))
        h[: zt.shape[0]] = zt[:, s0:s0 + S_loc]
        t = torch.from_numpy(h).to(dev)
        ev = torch.cuda.Event()
        ev.record()
        with self.lock:
            self.dbytes += h.nbytes
            while self.dbytes > self.dcache_limit and self.dcache:
                k0 = next(iter(self.dcache))
                self.dbytes -= self.dcache.pop(k0)[0].numel() * 8
            self.dcache[key] = (t, ev)
        return t, ev


class TraceBatch:
    """B traces traced concurrently. Arguments mirror GP_Edge_Tracing (gpet.py:22-35) with a leading batch
    dimension on `init`, `grad_img` and `obs`.

    factor: 'device'   - device factor provider (low-rank eigensolver for kernels whose grid matrix has numerical
                         rank <= 160, e.g. RBF; otherwise full covariance + block Jacobi eigensolver in HBM);
            'host_svd' - parity mode: full covariance on device, numpy.linalg.svd per trace on the host
                         (exactly the factor numpy's multivariate_normal uses), canonical signs.
    """

    def __init__(self, init, grad_img, kernel_options=(1, 3, 3), noise_y=1, obs=None, N_samples=500, score_thresh=1,
                 delta_x=20, keep_ratio=0.1, pixel_thresh=5, seed=42, fix_endpoints=True, factor="device",
                 device=None, record=False, y_budget_bytes=6 << 30, timers=None, final_fit="device",
                 sample_group=None, device_rng="auto", fused=None):
        if not torch.cuda.is_available():
            raise GpetError("TraceBatch needs a CUDA device (there is no CPU fallback)")
        _cabi.load()
        self.dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        init = np.asarray(init)
        if init.ndim == 2:
            init = init[None]
        self.B = B = init.shape[0]
        # ---- scalar options, clamped exactly like gpet.py:95-119 ------------------------------------
        x_st = np.array([int(i[0, 0]) for i in init])
        x_en = np.array([int(i[-1, 0]) for i in init])
        if not (np.all(x_st == x_st[0]) and np.all(x_en == x_en[0])):
            raise GpetError("all traces of a batch must share the edge span [x_st, x_en]")
        self.x_st, self.x_en = int(x_st[0]), int(x_en[0])
        self.init = np.stack([i[np.argsort(i[:, 0])].astype(int) for i in init])       # gpet.py:95
        self.noise_y = noise_y
        self.N_samples = int(N_samples) if N_samples > 100 else 1000
        self.seed = seed
        self.keep_ratio = float(keep_ratio) if 0 < keep_ratio <= 1 else 0.1
        self.pixel_thresh = int(pixel_thresh) if pixel_thresh >= 2 else 2
        st = float(score_thresh) if 0 < score_thresh <= 1 else 1
        self._score_thresh = np.full(B, st, dtype=np.float64)
        self.delta_x = int(delta_x) if delta_x > 3 else 2
        self.fix_endpoints = bool(fix_endpoints)
        self.N_inits = self.init.shape[1]
        if torch.is_tensor(grad_img):
            g = grad_img.to(self.dev)
            if g.ndim == 2:
                g = g[None]
        else:
            g = np.asarray(grad_img)
            if g.ndim == 2:
                g = g[None]
            g = torch.from_numpy(np.ascontiguousarray(g.astype(np.float32))).to(self.dev)
        if g.shape[0] != B:
            raise GpetError(f"{B} init sets but {g.shape[0]} gradient images")
        self.M, self.N = int(g.shape[1]), int(g.shape[2])
        self.x_grid = self.x_st + np.arange(self.x_en - self.x_st + 1).astype(int)
        self.n = n = self.x_grid.shape[0]
        if self.x_st < 0 or self.x_en >= self.N:
            raise GpetError("edge endpoints outside the image")
        self.N_subints = int(n // self.delta_x)
        self.N_keep = int(keep_ratio * N_samples)                                          # gpet.py:118 (raw args)
        self.algo_thresh = self.N_subints - (self.pixel_thresh - 1)
        if not 0 < self.N_keep <= self.N_samples:
            raise GpetError(f"N_keep={self.N_keep} must be in 1..N_samples")
        self.ktype, self.nu, self.sigma_f, self.sigma_l = _gp_host.parse_kernel_options(kernel_options, self.M, n)
        self.alpha_init = np.array(self.N_inits * [[0.5, 1e-7][int(self.fix_endpoints)]])
        if obs is None:
            obs = [np.zeros((0, 2), dtype=np.int64)] * B
        elif isinstance(obs, np.ndarray) and obs.dtype != object and obs.ndim <= 2:
            obs = [obs] * B          # one observation set shared by (or for) every trace
        if len(obs) != B:
            raise GpetError(f"{B} traces but {len(obs)} observation sets")
        obs = [np.asarray(o).reshape(-1, 2).astype(np.int64) for o in obs]
        self.record = [] if record else None
        self.timers = timers
        # sample sharding (second row): every rank of `sample_group` (a torch.distributed process group,
        # or True for the default group) draws / scores its own block of the N_samples curves of EVERY trace; the
        # costs are all-gathered, the fixed-point density grids all-reduced, everything else is replicated
        self.sgroup, self.srank, self.sworld = None, 0, 1
        if sample_group is not None and sample_group is not False:
            import torch.distributed as tdist
            self.sgroup = None if sample_group is True else sample_group
            self.srank, self.sworld = tdist.get_rank(self.sgroup), tdist.get_world_size(self.sgroup)
        self.final_fit_mode = final_fit
        self.factor = factor
        if factor not in ("device", "host_svd"):
            raise GpetError(f"unknown factor provider {factor!r}")

        # ---- one-time device state: normalised gradient (gpet.py:97), its KDE (:127), transposed copy -----
        S = self.N_samples
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.grad = g.to(torch.float32).contiguous().clone()
        self._mm = torch.empty((B, 2), dtype=torch.int32, device=self.dev)
        call("gpet_normalise_f32", ptr(self.grad), B, self.M, self.N, ptr(self._mm), _stream())
        self.grad_kde = torch.empty((B, self.M, self.N), **f32)
        work = torch.empty(query("gpet_grad_kde_workspace_bytes", B, self.M, self.N), dtype=torch.uint8, device=self.dev)
        call("gpet_grad_kde_f32", ptr(self.grad), B, self.M, self.N, ptr(self.grad_kde), ptr(work), _stream())
        self.gradT = torch.empty((B, self.N, self.M + 2), **f32)     # guarded columns, see gpet_transpose_f32
        call("gpet_transpose_f32", ptr(self.grad), B, self.M, self.N, ptr(self.gradT), _stream())
        del work

        # ---- kernel tables ------------------------------------------------------------------------------------
        kd, Ur, lam, r = _gp_host.grid_eigenbasis(self.ktype, self.nu, self.sigma_l, self.x_grid, MAX_RANK)
        self.rank = r
        f64 = dict(dtype=torch.float64, device=self.dev)
        self.kd = torch.from_numpy(kd).to(self.dev)
        self.lowrank = (Ur is not None) and factor == "device"
        if self.lowrank:
            self.rp = Ur.shape[1]
            self.Ur = torch.from_numpy(Ur).to(self.dev)
            self.lam = torch.from_numpy(lam).to(self.dev)
            self.uw = torch.from_numpy(Ur.T @ _gp_host.sign_weights(n)).to(self.dev)
        else:
            self.rp = ((n + 3) // 4) * 4
        self.draws = NormalDraws.shared(S, n, min(self.rp, n), seed)

        # ---- bins / groups for the selection kernel ---------------------------------------------------------
        # band-limited density + selection (gpet_density_bands_f64: one column group of one trace per CTA, all in shared
        # memory; single-rank runs): the groups are narrower so that (M + 8) x (width + 8) 64-bit cells fit one SM
        self.bands_width = 0
        if (sample_group is None or sample_group is False) and os.environ.get("GPET_DENSITY_BANDS", "1") != "0":
            fit_cols = (220 * 1024 - 8) // (8 * (self.M + 8))          # histogram columns that fit beside M + 8 rows
            # 16-column groups (two CTAs of the band kernel per SM, 80 KB each at M = 500): alone on the GPU they are as fast
            # as 32-column groups (2.27 against 2.32 ms per 1250-trace launch), next to the final-fit stream they are faster
            # (density stage 48 against 62 ms per step, step 281 against 287 ms) - smaller CTAs find room beside the
            # objective kernel's
            width = min(int(os.environ.get("GPET_BANDS_WIDTH", "16")), (fit_cols if fit_cols % 2 else fit_cols - 1) - 8)
            try:
                if width >= 1:
                    col_bin, group_cols, self.nb, self.bin_lo = _gp_host.column_bins(self.N, self.x_st, self.x_en, self.delta_x,
                                                                                    self.fix_endpoints, max_group=width)
                    w_max = int(np.diff(group_cols).max())
                    if query("gpet_density_bands_supported", self.M, self.N, w_max):
                        self.bands_width = w_max
            except ValueError:
                pass
        if not self.bands_width:
            col_bin, group_cols, self.nb, self.bin_lo = _gp_host.column_bins(self.N, self.x_st, self.x_en, self.delta_x,
                                                                            self.fix_endpoints)
        self.col_bin = torch.from_numpy(col_bin).to(self.dev)
        self.group_cols = torch.from_numpy(group_cols).to(self.dev)
        self.n_groups = len(group_cols) - 1
        self.max_old = max([self.nb] + [o.shape[0] for o in obs])
        self.mmax = self.N_inits + self.max_old
        # loop-carried state (gpet.py:829-870): lives on the device (gpet_control.cu); these are host mirrors, pulled
        # on demand (properties obs / n_obs / score_thresh / n_iter) and pushed when a caller changed them (set_obs)
        self._obs = np.zeros((B, self.max_old, 2), dtype=np.int64)     # accepted observations (x, y), padded
        self._n_obs = np.zeros(B, dtype=np.int64)
        self._n_iter = np.zeros(B, dtype=np.int64)
        self._host_dirty, self._dev_newer = True, False
        for b, o in enumerate(obs):
            self.set_obs(b, o)
        # more training points than the shared-memory kernels hold: the C-ABI entry points switch to the HBM-resident
        # blocked path by themselves (gpet_dense.cu); the flag only picks the `_big` final-fit entry points
        self.large_m = self.mmax > MAX_TRAIN

        # ---- per-iteration buffers -------------------------------------------------------------------------------
        i32 = dict(dtype=torch.int32, device=self.dev)
        mm = self.mmax
        self.d_init = torch.from_numpy(np.ascontiguousarray(self.init.astype(np.int32))).to(self.dev)
        self.d_alpha_init = torch.from_numpy(np.ascontiguousarray(self.alpha_init, dtype=np.float64)).to(self.dev)
        self.d_obs = torch.zeros((B, self.max_old, 2), **i32)
        self.d_nobs = torch.zeros((B,), **i32)
        self.d_thr = torch.zeros((B,), **f64)
        self.d_niter = torch.zeros((B,), **i32)
        self.d_ctrl = torch.zeros((4,), **i32)
        self.h_ctrl = torch.zeros((4,), dtype=torch.int32).pin_memory()
        self.d_xi = torch.empty((B, mm), **i32)
        self.d_y = torch.empty((B, mm), **f64)
        self.d_w = torch.empty((B, mm), **f64)
        self.d_m = torch.empty((B,), **i32)
        self.d_old = torch.empty((B, self.max_old, 2), **i32)
        self.d_nold = torch.empty((B,), **i32)
        self.d_sigma_f = torch.full((B,), float(self.sigma_f), **f64)
        self.d_mean = torch.empty((B, n), **f64)
        self.d_ys = torch.empty((B,), **f64)
        self.d_status = torch.empty((B,), **i32)
        self.d_A = torch.empty((B, self.rp, n), **f64) if self.lowrank else None
        if self.lowrank:
            self.d_Mr = torch.empty((B, self.rp, self.rp), **f64)
            self.d_d = torch.empty((B, self.rp), **f64)
            self.d_Q = torch.empty((B, self.rp, self.rp), **f64)
            self.d_sweeps = torch.empty((B,), **i32)
            self.d_eig_work = torch.empty(query("gpet_sym_eig_workspace_bytes", B, self.rp), dtype=torch.uint8,
                                          device=self.dev)
            nbytes = query("gpet_posterior_lowrank_workspace_bytes", B, self.mmax, self.rp)
            self.d_post_work = torch.empty(nbytes, dtype=torch.uint8, device=self.dev) if nbytes else None
        if S % self.sworld:
            raise GpetError(f"N_samples={S} must be divisible by the {self.sworld} ranks of the sample group")
        self.S_loc = Sl = S // self.sworld                                  # curves drawn and scored by this rank
        self.s0 = self.srank * Sl
        self.d_Zt = None      # host generator: NormalDraws.device() tensors, shared; device generator: own buffer
        # standard normals: host (numpy itself, bit exact, produced ahead of use by a thread and shared by all traces)
        # or device (gpet_rng.cu: the same array bit for bit, see device_standard_normal); "auto" switches to the device when one draw is large
        # enough for the host generator to become the bottleneck (BASELINE config 2: 50 M normals = 1.8 s per iteration)
        self.device_rng = (S * n >= 4_000_000) if device_rng == "auto" else bool(device_rng)
        if self.device_rng:
            self.d_Zt = torch.zeros((self.rp, Sl), **f64)
            self.d_rng_work = torch.empty(query("gpet_standard_normal_workspace_bytes", S, n), dtype=torch.uint8,
                                          device=self.dev)
            self.d_rng_fix = torch.empty(query("gpet_standard_normal_fixup_bytes", S, n), dtype=torch.uint8, device=self.dev)
            self.d_rng_ok = torch.ones(1, dtype=torch.int32, device=self.dev)
            self.h_rng_ok = torch.ones(1, dtype=torch.int32).pin_memory()
        # fused sampling + scoring (gpet_sample_score_f64; GPET_FUSED=1 or fused=True): the curves never reach HBM, the
        # N_keep kept ones are recomputed for the density.  Needs the low-rank factor with rp <= 80 and an even edge
        # length; a recording batch additionally materialises the curves with the unfused sampler.  OFF by default:
        # measured on the cfg 5 shard it is slower than the unfused pair (sample_score 126 + keep 17 ms per step against
        # sample 69 + score 28) - neither kernel is HBM bound (DMMA pipe / FP64 issue), and the scoring half needs the
        # latency hiding of ~16 warps per SM, which the register-resident GEMM fragments leave no room for.  It saves
        # the Y buffer (4 MB per trace at cfg 1 sizes), which is what matters when N_samples is large.
        want = os.environ.get("GPET_FUSED", "0") == "1" if fused is None else bool(fused)
        self.fused = bool(want and self.lowrank and self.sworld == 1 and
                          query("gpet_sample_score_supported", self.rp, n, Sl))
        need_Y = (not self.fused) or record
        per_trace = n * Sl * 8 if need_Y else self.M * self.N * 12 + n * self.N_keep * 8
        self.Bc = int(max(1, min(B, y_budget_bytes // per_trace)))            # traces per chunk of the sample..select stages
        self.d_Y = torch.empty((self.Bc, n, Sl), **f64) if need_Y else None
        if self.fused:
            self.d_Yk = torch.empty((self.Bc, n, self.N_keep), **f64)
            self.d_idx_id = torch.arange(self.N_keep, dtype=torch.int32, device=self.dev).repeat(self.Bc, 1).contiguous()
        self.d_cost = torch.empty((B, S), **f64)
        if self.sworld > 1:
            self.d_cost_loc = torch.empty((self.Bc, Sl), **f64)
            self.d_idx_loc = torch.empty((self.Bc, self.N_keep), **i32)
        self.d_idx = torch.empty((B, self.N_keep), **i32)
        self.d_best = torch.empty((B, self.N_keep), **f64)
        self.d_wts = torch.empty((B, self.N_keep), **f64)
        self.d_dens = torch.empty((self.Bc, self.M, self.N), **f32)
        self.d_dmm = torch.empty((self.Bc, 2), **i32)
        if self.bands_width:
            self.d_bands = torch.empty((self.Bc, self.n_groups, 2), **i32)
            self.d_dwork = torch.empty(query("gpet_density_bands_workspace_bytes", self.Bc, n, max(self.N_keep, 1)),
                                       dtype=torch.uint8, device=self.dev)
        else:
            self.d_dwork = torch.empty(query("gpet_density_workspace_bytes", self.Bc, self.M, self.N, max(self.N_keep, 1)),
                                       dtype=torch.uint8, device=self.dev)
        self.d_bscore = torch.empty((B, self.nb), **f64)
        self.d_bpos = torch.empty((B, self.nb), **i32)
        self.d_rows = torch.empty((B,), **i32)
        self.d_rows_prev = torch.empty((B,), **i32)
        self._done_ev = torch.cuda.Event()
        self._pending = None
        self.stream = None
        self._n_active = None          # active traces (known to the host after the control block was read)
        self._m_cap = self.mmax        # bound on the training-set sizes of the next posterior launch (control block)
        self._it = 0                   # iterations done by the traces that are still active
        self._released = False
        self.host_ms = {}
        self.curves_scored = 0
        self.kernel_launches = 0

    def _density_bands(self, Y, idx, wts, nbk, S, st):
        """kernel_density_estimate of the kept curves (gpet.py:455-529), band-limited: gpet_density_bands_f64."""
        self._stage("density", "gpet_density_bands_f64", ptr(Y), ptr(idx), ptr(wts), nbk, self.n, S, self.N_keep, self.M, self.N,
                    self.x_st, ptr(self.group_cols), self.n_groups, self.bands_width, ptr(self.d_dens), ptr(self.d_dmm),
                    ptr(self.d_bands), ptr(self.d_dwork), st)

    @contextlib.contextmanager
    def _boosted(self, stage):
        """Yields the CUDA stream handle the stage should launch on: the current stream, or - for the stages of GPET_BOOST
        - the high-priority side stream, ordered after everything queued so far and joined back afterwards."""
        side = _boost_stream(self.dev, stage)
        if side is None:
            yield _stream()
            return
        cur = torch.cuda.current_stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            yield side.cuda_stream
        cur.wait_stream(side)

    # ----------------------------------------------------------------------------------------------------
    def _stage(self, stage, name, *args):
        """One C-ABI call, optionally bracketed by CUDA events."""
        if self.timers is None:
            return call(name, *args)
        e0 = self.timers.start()
        call(name, *args)
        self.timers.stop(stage, e0)

    # ---- loop-carried state: device <-> host mirrors -----------------------------------------------------------------
    def _pull_state(self):
        """Host mirrors of the device-resident loop state (one synchronous copy; only when the device is ahead)."""
        if self._dev_newer:
            self._obs[:] = self.d_obs.cpu().numpy()
            self._n_obs[:] = self.d_nobs.cpu().numpy()
            self._score_thresh[:] = self.d_thr.cpu().numpy()
            self._n_iter[:] = self.d_niter.cpu().numpy()
            self._dev_newer = False

    @property
    def obs(self):
        self._pull_state()
        return self._obs

    @property
    def n_obs(self):
        self._pull_state()
        return self._n_obs

    @property
    def n_iter(self):
        self._pull_state()
        return self._n_iter

    @property
    def score_thresh(self):
        """self.score_thresh of every trace (gpet.py:595: decays in place across iterations). Read-only view; use
        set_score_thresh to change it."""
        self._pull_state()
        return self._score_thresh

    def set_score_thresh(self, b, v):
        self._pull_state()
        self._score_thresh[b] = float(v)
        self._host_dirty = True

    @property
    def fobs(self):
        """Per-trace accepted observations, int64[k, 2] in xy order (the reference's `pre_fobs`)."""
        obs, n_obs = self.obs, self.n_obs
        return [obs[b, : n_obs[b]].copy() for b in range(self.B)]

    def set_obs(self, b, arr):
        self._pull_state()
        arr = np.asarray(arr).reshape(-1, 2)
        if arr.shape[0] > self.max_old:
            raise GpetError(f"{arr.shape[0]} observations, but this batch was built for at most {self.max_old}")
        self._obs[b, : arr.shape[0]] = arr
        self._n_obs[b] = arr.shape[0]
        self._host_dirty = True

    def active(self):
        return self.n_obs < self.algo_thresh

    def curve_buffer(self):
        """Y[Bc][n][S_loc] for callers that want the curves themselves (the stage seams of GP_Edge_Tracing); the fused
        loop does not materialise them, so the buffer is created on first use."""
        if self.d_Y is None:
            self.d_Y = torch.empty((self.Bc, self.n, self.S_loc), dtype=torch.float64, device=self.dev)
        return self.d_Y

    def _ensure_device_state(self, all_traces=False):
        """Pushes host-side changes of the loop state (constructor, set_obs, set_score_thresh) and builds the training
        sets of the active traces (gpet_training_sets_f64). all_traces: every trace gets a slot, converged or not (the
        fit_predict_GP seam of GP_Edge_Tracing)."""
        if not (self._host_dirty or all_traces or self._n_active is None):
            return
        if self._released:
            raise GpetError("this TraceBatch has released its loop buffers (release_loop_buffers)")
        self._pull_state()
        if self._host_dirty or self._n_active is None:
            self.d_obs.copy_(torch.from_numpy(self._obs.astype(np.int32)))
            self.d_nobs.copy_(torch.from_numpy(self._n_obs.astype(np.int32)))
            self.d_thr.copy_(torch.from_numpy(self._score_thresh))
            self.d_niter.copy_(torch.from_numpy(self._n_iter.astype(np.int32)))
            self.d_ctrl.zero_()
        thresh = (1 << 30) if all_traces else self.algo_thresh
        call("gpet_training_sets_f64", ptr(self.d_init), ptr(self.d_alpha_init), self.N_inits, ptr(self.d_obs),
             ptr(self.d_nobs), self.B, self.B, self.max_old, thresh, self.x_st, self.mmax, 0, ptr(self.d_rows),
             ptr(self.d_ctrl), ptr(self.d_xi), ptr(self.d_y), ptr(self.d_w), ptr(self.d_m), ptr(self.d_old),
             ptr(self.d_nold), ptr(self.h_ctrl), _stream())
        torch.cuda.current_stream().synchronize()
        self._n_active = int(self.h_ctrl[0])
        self._m_cap = self.mmax if all_traces else min(self.mmax, self.N_inits + int(self.h_ctrl[3]))
        act = self._n_iter[self._n_obs < self.algo_thresh]
        if act.size and not np.all(act == act[0]):
            raise GpetError("lock-step violated: active traces are at different iterations")
        self._it = int(act[0]) if act.size else 0
        self._host_dirty = bool(all_traces)      # the slots of an all_traces build are not those of the loop
        self.kernel_launches += 2

    def _host_training_sets(self, rows=None):
        """gpet.py:209-224 on the host mirrors (used by the final-fit preparation): concat(init,
        obs), stable sort by x, noise weights. Returns (x int64[b, mmax], y float64[b, mmax], w float64[b, mmax], m)."""
        K, mo = self.N_inits, self.max_old
        obs_all, n_obs_all = self.obs, self.n_obs
        init = self.init if rows is None else self.init[rows]
        obs = obs_all if rows is None else obs_all[rows]
        n_obs = n_obs_all if rows is None else n_obs_all[rows]
        B = init.shape[0]
        m = (K + n_obs).astype(np.int32)
        valid = np.concatenate([np.ones((B, K), dtype=bool), np.arange(mo)[None, :] < n_obs[:, None]], axis=1)
        x = np.concatenate([init[:, :, 0], obs[:, :, 0]], axis=1)
        y = np.concatenate([init[:, :, 1], obs[:, :, 1]], axis=1).astype(np.float64)
        w = np.concatenate([np.broadcast_to(self.alpha_init, (B, K)), np.ones((B, mo))], axis=1)
        key = np.where(valid, x, np.iinfo(np.int64).max)
        order = np.argsort(key, axis=1, kind="stable")
        x = np.take_along_axis(np.where(valid, x, self.x_st), order, axis=1)
        y = np.take_along_axis(np.where(valid, y, 0.0), order, axis=1)
        w = np.take_along_axis(np.where(valid, w, 0.0), order, axis=1)
        return x, y, w, m

    _training_sets = _host_training_sets

    def _factor_full(self, it, B):
        """Full-covariance providers for the B compacted active traces: returns A[B, rp, n] (rp = n padded to 4) on
        the device."""
        n = self.n
        cov = torch.empty((B, n, n), dtype=torch.float64, device=self.dev)
        work = torch.empty(query("gpet_posterior_full_workspace_bytes", B, self.mmax, n), dtype=torch.uint8,
                           device=self.dev)
        call("gpet_posterior_full_f64", ptr(self.d_xi), ptr(self.d_y), ptr(self.d_w), ptr(self.d_m), self.mmax, B, n,
             ptr(self.d_sigma_f), float(self.noise_y), _gp_host.GP_ALPHA, ptr(self.kd), ptr(self.d_mean),
             ptr(self.d_ys), ptr(cov), ptr(self.d_status), ptr(work), _stream())
        self.kernel_launches += 2
        if self.factor == "host_svd":
            A = torch.zeros((B, self.rp, n), dtype=torch.float64, device=self.dev)
            cov_h = cov.cpu().numpy()
            for b in range(B):
                A[b, :n] = torch.from_numpy(_gp_host.canonical_factor_host(cov_h[b])).to(self.dev)
        else:
            # full-rank kernels (Matern): block Jacobi eigensolver in HBM (gpet_jacobi.cu)
            A = self._factor_block_jacobi(cov, B)
        self._last_cov = cov if self.record is not None else None
        return A

    def _factor_block_jacobi(self, cov, B, tol=3e-13, max_sweeps=30):
        """numpy's svd factor of the full covariances cov[B, n, n] (sklearn_gpr.py:460-464) from their eigen-decomposition
        by two-sided block Jacobi on the device: sweeps until the off-diagonal Frobenius norm is below tol * ||A||_F or has
        reached its rounding floor (below 1e-9 and no longer shrinking by a factor of 4). From the second iteration on the
        sweeps start from the eigenvectors of the previous iteration's covariance (same active traces): the rotated matrix
        is nearly diagonal and 2-4 sweeps suffice instead of 8 (GPET_JACOBI_WARM=0: always from the identity).
        Returns F[B, rp, n]."""
        n, rp = self.n, self.rp
        np_ = ((n + 127) // 128) * 128
        f64 = dict(dtype=torch.float64, device=self.dev)
        st = _stream()
        jac = getattr(self, "_jac", None)
        if jac is None:
            jac = self._jac = dict(A=torch.empty((B, np_, np_), **f64), V=torch.empty((B, np_, np_), **f64),
                                   off=torch.empty((B, 2), **f64), tmp=None, B=None, rows=None, sweeps=[],
                                   work=torch.empty(query("gpet_block_jacobi_workspace_bytes", B, np_), dtype=torch.uint8,
                                                    device=self.dev),
                                   w=torch.from_numpy(_gp_host.sign_weights(n)).to(self.dev))
            self.jacobi_sweeps = jac["sweeps"]          # sweeps per iteration (survives release_loop_buffers)
        Aj, Vj, off, work = jac["A"], jac["V"], jac["off"], jac["work"]
        warm = jac["B"] is not None and os.environ.get("GPET_JACOBI_WARM", "1") != "0"
        rows = self.d_rows[:B]
        if warm and jac["B"] != B:
            # converged traces left and the rest was re-packed to the front: move the eigenvectors along (device side)
            pos = (rows[:, None] == jac["rows"][None, :]).to(torch.int32).argmax(dim=1)
            Vj[:B] = Vj.index_select(0, pos)
        jac["rows"] = rows.clone()
        if warm:
            if jac["tmp"] is None:
                jac["tmp"] = torch.empty((2, Aj.shape[0], np_, np_), **f64)
            call("gpet_block_jacobi_warm_f64", ptr(cov), B, n, np_, ptr(Aj), ptr(Vj), ptr(jac["tmp"]), st)
        else:
            call("gpet_block_jacobi_init_f64", ptr(cov), B, n, np_, ptr(Aj), ptr(Vj), st)
        jac["B"] = B
        prev, sweeps = None, 0
        for _ in range(max_sweeps):
            call("gpet_block_jacobi_sweep_f64", ptr(Aj), ptr(Vj), B, np_, ptr(off), ptr(work), st)
            sweeps += 1
            o = off[:B].cpu().numpy()
            rel = np.sqrt(o[:, 0] / np.maximum(o[:, 1], 1e-300))
            if np.all(rel <= tol) or (prev is not None and np.all((rel <= tol) | ((rel <= 1e-9) & (rel > 0.25 * prev)))):
                break
            prev = rel
        jac["sweeps"].append(sweeps)
        F = torch.empty((B, rp, n), **f64)
        call("gpet_block_jacobi_factor_f64", ptr(Aj), ptr(Vj), B, n, np_, rp, ptr(jac["w"]), ptr(F), ptr(work), st)
        self.kernel_launches += (6 if warm else 1) + sweeps * ((np_ // 32 - 1) * 7 + 1) + 2
        return F

    def step(self):
        """One pass of the while-loop body (gpet.py:839-861) for every unfinished trace."""
        if not self.step_launch():
            return False
        self.step_finish()
        return True

    def use_own_stream(self, priority=0):
        """Gives this batch a CUDA stream of its own (ordered after everything queued so far on the current stream).
        In a pipelined run the sub-batches inside the window then overlap ON the GPU as well: the latency-bound
        kernels of one (the serial QL recurrences, the per-trace Cholesky, the top-k sort) run next to the
        bandwidth / FP64-bound kernels of the other instead of each leaving most SMs idle in turn."""
        if self.stream is None:
            self.stream = torch.cuda.Stream(priority=priority)
            self.stream.wait_stream(torch.cuda.current_stream())

    def step_launch(self):
        """Device half of one iteration: enqueues every kernel of the iteration - including the update of the
        observation sets / thresholds and the training sets of the next iteration (gpet_control.cu) - and the copy of
        the 16-byte control block on this batch's stream (the current one unless use_own_stream() was called), without
        waiting. Returns False when every trace is done."""
        if self.stream is not None and torch.cuda.current_stream() != self.stream:
            with torch.cuda.stream(self.stream):
                return self.step_launch()
        self._ensure_device_state()
        B = self._n_active
        if B == 0:
            return False
        n, S, Kp, M, N = self.n, self.N_samples, self.N_keep, self.M, self.N
        it = self._it
        st = _stream()
        if self.device_rng:
            device_standard_normal(self.seed + it + 1, S, n, min(self.rp, n), self.s0, self.S_loc, self.d_Zt, self.d_rng_ok,
                                   self.d_rng_fix, self.d_rng_work)                                # gpet.py:839
            self.h_rng_ok.copy_(self.d_rng_ok, non_blocking=True)
            self.kernel_launches += 5
        else:
            self.d_Zt, z_ev = self.draws.device(it, self.dev, self.s0, self.S_loc, self.rp)
            torch.cuda.current_stream().wait_event(z_ev)
        if self.lowrank:
            self._stage("posterior", "gpet_posterior_lowrank_f64", ptr(self.d_xi), ptr(self.d_y), ptr(self.d_w), ptr(self.d_m), self.mmax,
                 self._m_cap, B, n, ptr(self.d_sigma_f), float(self.noise_y), _gp_host.GP_ALPHA, ptr(self.kd), ptr(self.Ur),
                 ptr(self.lam), self.rp, ptr(self.d_mean), ptr(self.d_ys), ptr(self.d_Mr), ptr(self.d_status),
                 ptr(self.d_post_work), st)
            self._stage("eig", "gpet_sym_eig_f64", ptr(self.d_Mr), B, self.rp, ptr(self.d_d), ptr(self.d_Q), ptr(self.d_sweeps), ptr(self.d_eig_work), st)
            self._stage("assemble", "gpet_factor_assemble_f64", ptr(self.d_d), ptr(self.d_Q), ptr(self.Ur), ptr(self.uw), B, self.rp, n,
                 ptr(self.d_A), st)
            A = self.d_A
            self.kernel_launches += 3
        else:
            A = self._factor_full(it, B)
        rec = None
        if self.record is not None:
            rows = self.d_rows[:B].cpu().numpy()
            act = np.zeros(self.B, dtype=bool)
            act[rows] = True
            rec = dict(it=it, active=act, rows=rows.copy(), A=self._expand(A[:B].cpu().numpy(), rows),
                       mean=self._expand(self.d_mean[:B].cpu().numpy(), rows),
                       ys=self._expand(self.d_ys[:B].cpu().numpy(), rows), obs_in=[f.copy() for f in self.fobs],
                       samples=[], kde=[], thr_in=self.score_thresh.copy())
            if self.lowrank:
                rec["sweeps"] = self._expand(self.d_sweeps[:B].cpu().numpy(), rows)
                rec["d"] = self._expand(self.d_d[:B].cpu().numpy(), rows)
        Sl = self.S_loc
        for b0 in range(0, B, self.Bc):
            b1 = min(B, b0 + self.Bc)
            nbk = b1 - b0
            if self.fused:
                self._stage("sample_score", "gpet_sample_score_f64", ptr(self.d_Zt), ptr(A[b0:b1]), ptr(self.d_mean[b0:b1]),
                            ptr(self.d_ys[b0:b1]), ptr(self.gradT), ptr(self.d_rows[b0:b1]), nbk, self.rp, n, S, M, N, self.x_st,
                            ptr(self.d_cost[b0:b1]), st)
                if rec is not None:       # the record wants every curve: the unfused sampler writes them out as well
                    call("gpet_sample_f64", ptr(self.d_Zt), ptr(A[b0:b1]), ptr(self.d_mean[b0:b1]), ptr(self.d_ys[b0:b1]), nbk,
                         self.rp, n, Sl, ptr(self.d_Y), st)
            else:
                with self._boosted("sample") as sst:
                    self._stage("sample", "gpet_sample_f64", ptr(self.d_Zt), ptr(A[b0:b1]), ptr(self.d_mean[b0:b1]),
                                ptr(self.d_ys[b0:b1]), nbk, self.rp, n, Sl, ptr(self.d_Y), sst)
            if self.fused:
                pass
            elif self.sworld == 1:
                with self._boosted("score") as sst:
                    self._stage("score", "gpet_score_f64", ptr(self.d_Y), ptr(self.gradT), ptr(self.d_rows[b0:b1]), nbk, n, S, M, N,
                                self.x_st, ptr(self.d_cost[b0:b1]), sst)
            else:
                from . import dist as gdist
                self._stage("score", "gpet_score_f64", ptr(self.d_Y), ptr(self.gradT), ptr(self.d_rows[b0:b1]), nbk, n, Sl, M, N,
                     self.x_st, ptr(self.d_cost_loc), st)
                gdist.gather_costs(self.d_cost_loc[:nbk], self.d_cost[b0:b1], self.sgroup)   # every rank needs every cost
            self._stage("topk", "gpet_topk_f64", ptr(self.d_cost[b0:b1]), nbk, S, Kp, ptr(self.d_idx[b0:b1]), ptr(self.d_best[b0:b1]),
                 ptr(self.d_wts[b0:b1]), st)
            if self.fused:
                self._stage("keep", "gpet_sample_keep_f64", ptr(self.d_Zt), ptr(A[b0:b1]), ptr(self.d_mean[b0:b1]), ptr(self.d_ys[b0:b1]),
                            ptr(self.d_idx[b0:b1]), nbk, self.rp, n, S, Kp, ptr(self.d_Yk), st)
                if self.bands_width:
                    self._density_bands(self.d_Yk, self.d_idx_id, self.d_wts[b0:b1], nbk, Kp, st)
                else:
                    self._stage("density", "gpet_density_f64", ptr(self.d_Yk), ptr(self.d_idx_id), ptr(self.d_wts[b0:b1]), nbk, n, Kp, Kp, M, N,
                         self.x_st, ptr(self.d_dens), ptr(self.d_dmm), ptr(self.d_dwork), st)
            elif self.bands_width:
                with self._boosted("density") as sst:
                    self._density_bands(self.d_Y, self.d_idx[b0:b1], self.d_wts[b0:b1], nbk, S, sst)
            elif self.sworld == 1:
                self._stage("density", "gpet_density_f64", ptr(self.d_Y), ptr(self.d_idx[b0:b1]), ptr(self.d_wts[b0:b1]), nbk, n, S, Kp, M, N,
                     self.x_st, ptr(self.d_dens), ptr(self.d_dmm), ptr(self.d_dwork), st)
            else:
                # kept curves owned by this rank, as local sample indices (others -> -1: skipped by the splat)
                self.d_idx_loc[:nbk].copy_(gdist.local_keep_index(self.d_idx[b0:b1], self.s0, Sl))
                self._stage("density", "gpet_density_splat_f64", ptr(self.d_Y), ptr(self.d_idx_loc), ptr(self.d_wts[b0:b1]), nbk, n, Sl,
                     Kp, M, N, self.x_st, ptr(self.d_dwork), st)
                # exact, order-independent integer sums: fixed-point density grid and dropped-point counts
                gdist.reduce_density(self.d_dwork, nbk, M, N, Kp, self.sgroup)
                self._stage("density", "gpet_density_finish_f64", ptr(self.d_wts[b0:b1]), nbk, n, Kp, M, N, ptr(self.d_dens),
                     ptr(self.d_dmm), ptr(self.d_dwork), st)
            if self.bands_width:
                self._stage("select", "gpet_select_bands_f64", ptr(self.d_dens), ptr(self.d_dmm), ptr(self.grad_kde),
                            ptr(self.d_rows[b0:b1]), ptr(self.d_bands), nbk, M, N, ptr(self.col_bin), ptr(self.group_cols),
                            self.n_groups, ptr(self.d_old[b0:b1]), ptr(self.d_nold[b0:b1]), self.max_old, self.nb,
                            ptr(self.d_bscore[b0:b1]), ptr(self.d_bpos[b0:b1]), st)
            else:
                self._stage("select", "gpet_select_f64", ptr(self.d_dens), ptr(self.d_dmm), ptr(self.grad_kde), ptr(self.d_rows[b0:b1]), nbk, M, N,
                     ptr(self.col_bin), ptr(self.group_cols), self.n_groups, ptr(self.d_old[b0:b1]), ptr(self.d_nold[b0:b1]),
                     self.max_old, self.nb, ptr(self.d_bscore[b0:b1]), ptr(self.d_bpos[b0:b1]), st)
            self.kernel_launches += 8
            if rec is not None:
                rec["samples"].append(self.d_Y[:nbk].cpu().numpy())
                kde = torch.empty((nbk, M, N), dtype=torch.float32, device=self.dev)
                if self.bands_width:
                    call("gpet_kde_bands_f32", ptr(self.d_dens), ptr(self.d_dmm), ptr(self.d_bands), ptr(self.group_cols),
                         self.n_groups, nbk, M, N, ptr(kde), st)
                else:
                    call("gpet_kde_normalised_f32", ptr(self.d_dens), ptr(self.d_dmm), nbk, M, N, ptr(kde), st)
                rec["kde"].append(kde.cpu().numpy())
        if rec is not None:
            rec.update(bin_score=self._expand(self.d_bscore[:B].cpu().numpy(), rows),
                       bin_pos=self._expand(self.d_bpos[:B].cpu().numpy(), rows))
        # compute_new_obs (gpet.py:589-616) + the training sets of the next iteration, all on the device
        self.d_rows_prev[:B].copy_(self.d_rows[:B])       # slots of THIS iteration (error reporting)
        self._stage("control", "gpet_update_obs_f64", ptr(self.d_bscore), ptr(self.d_bpos), ptr(self.d_rows), ptr(self.d_status), B,
                    self.nb, N, self.max_old, self.pixel_thresh, self.algo_thresh, ptr(self.d_obs), ptr(self.d_nobs),
                    ptr(self.d_thr), ptr(self.d_niter), ptr(self.d_ctrl), st)
        self._stage("control", "gpet_training_sets_f64", ptr(self.d_init), ptr(self.d_alpha_init), self.N_inits, ptr(self.d_obs),
                    ptr(self.d_nobs), self.B, B, self.max_old, self.algo_thresh, self.x_st, self.mmax, 1, ptr(self.d_rows),
                    ptr(self.d_ctrl), ptr(self.d_xi), ptr(self.d_y), ptr(self.d_w), ptr(self.d_m), ptr(self.d_old),
                    ptr(self.d_nold), ptr(self.h_ctrl), st)
        self.kernel_launches += 3
        self._done_ev.record()
        self._dev_newer = True
        self._pending = (B, rec)
        return True

    def _expand(self, a, rows):
        """Scatters a compacted per-active-trace array back to the full batch (zeros elsewhere)."""
        out = np.zeros((self.B,) + a.shape[1:], dtype=a.dtype)
        out[rows] = a
        return out

    def step_finish(self):
        """Host half of one iteration (after step_launch): waits for the control block - how many traces are still
        inside the while-loop, whether a trace failed - and nothing else."""
        B, rec = self._pending
        self._pending = None
        S = self.N_samples
        self._done_ev.synchronize()
        if self.device_rng and int(self.h_rng_ok[0]) != 1:
            raise GpetError("device normal generator: attempt budget exhausted (probability ~1e-15); rerun")
        n_active, err, bad = (int(v) for v in self.h_ctrl[:3])
        if err == 1:
            raise np.linalg.LinAlgError(f"Cholesky of the training kernel matrix failed (trace {bad}) "
                                        "(sklearn_gpr.py:306-314)")
        if err != 0:
            # cross-check on the host before blaming the data: the same per-bin maxima through the numpy restatement
            rows_prev = self.d_rows_prev[:B].cpu().numpy()
            slot = int(np.flatnonzero(rows_prev == bad)[0])
            best = self.d_bscore[slot:slot + 1].cpu().numpy()
            n_pre = self.d_nobs[bad:bad + 1].cpu().numpy().astype(np.int64)
            thr = self.d_thr[bad:bad + 1].cpu().numpy().copy()
            try:
                _gp_host.threshold_loop_batch(best, n_pre, self.pixel_thresh, self.algo_thresh, thr, np.ones(1, dtype=bool))
            except RuntimeError:
                pass
            else:
                raise GpetError(f"gpet_update_obs_f64 flagged trace {bad} but the host threshold loop ends: kernel defect "
                                f"(non-empty bins {int((best > 0).sum())}, n_pre {int(n_pre[0])}, thr {float(thr[0])})")
            raise RuntimeError(f"compute_new_obs: score threshold decayed to zero without enough new pixels (trace {bad}; "
                               "the reference loops forever here, gpet.py:591-609)")
        self.curves_scored += B * S
        self._n_active = n_active
        self._m_cap = min(self.mmax, self.N_inits + int(self.h_ctrl[3]))   # bound on the training sets of the next iteration
        self._it += 1
        if rec is not None:
            rows = rec["rows"]
            rec.update(costs=self._expand(self.d_cost[:B].cpu().numpy(), rows),
                       keep_idx=self._expand(self.d_idx[:B].cpu().numpy(), rows),
                       best_costs=self._expand(self.d_best[:B].cpu().numpy(), rows),
                       wts=self._expand(self.d_wts[:B].cpu().numpy(), rows), fobs=[f.copy() for f in self.fobs],
                       thr_out=self.score_thresh.copy())
            rec["samples"] = self._expand(np.concatenate(rec["samples"], axis=0), rows)
            rec["kde"] = self._expand(np.concatenate(rec["kde"], axis=0), rows)
            if not self.lowrank and self._last_cov is not None:
                rec["cov"] = self._expand(self._last_cov[:B].cpu().numpy(), rows)
            self.record.append(rec)

    def release_loop_buffers(self):
        """Frees everything only the while-loop needs (posterior curves, density grids, factors, gradient images and
        their transposed / KDE copies: ~7 MB per trace at 500 x 500, S = 1000). The final fit works from the host
        mirrors of the observation sets alone. Called by trace_pipelined when a sub-batch has converged; the batch
        cannot step afterwards."""
        if self._released:
            return
        self._pull_state()
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)
        for name in ("d_Y", "d_Yk", "d_idx_id", "d_dens", "d_dwork", "d_dmm", "d_bands", "d_A", "d_Mr", "d_Q", "d_d", "d_eig_work", "d_post_work", "d_sweeps", "gradT",
                     "grad_kde", "grad", "d_cost", "d_cost_loc", "d_idx_loc", "d_idx", "d_best", "d_wts", "d_bscore",
                     "d_bpos", "d_Zt", "d_rng_work", "d_rng_fix", "d_xi", "d_y", "d_w", "d_old", "d_obs", "_last_cov", "_jac"):
            if hasattr(self, name):
                setattr(self, name, None)
        self._released = True

    def run_loop(self, max_iters=100000):
        k = 0
        while self.step():
            k += 1
            if k > max_iters:
                raise GpetError("trace did not converge")
        return self.fobs

    def final_fit(self, b):
        """Converged branch + outputs for trace b (gpet.py:874-886) with scipy's own L-BFGS-B on the HOST
        (final_fit="host"): a parity instrument for tests - the product path is final_fit_group on the device."""
        X, y, w = _gp_host.assemble_training_set(self.init[b], self.obs[b, : self.n_obs[b]], self.alpha_init)
        y_mean, y_std, theta = _gp_host.final_fit(X.astype(np.float64), y, w, self.x_grid, self.ktype, self.nu,
                                                  self.noise_y, self.seed + int(self.n_iter[b]))
        cred = (y_mean - 1.96 * y_std, y_mean + 1.96 * y_std)
        curve = np.concatenate([self.x_grid[:, np.newaxis], y_mean[:, np.newaxis]], axis=1)
        edge = np.rint(curve[:, [1, 0]]).astype(int)
        return edge, cred, (y_mean, y_std, theta)

    def _fit_inputs(self, seed=None):
        """Host preparation of the final fit (gpet.py:232-248): standardised training sets and the 13 start points of
        every trace. Returns dict(Xs, yt, ws [B, mmax], ms [B], stats [B, 6] = (y_m, y_s, X_m, X_s, tm, ts),
        x0 [B, 13, 3], xc int32 [B, mmax] = the sorted integer pixel columns). seed: random_state of the restarts for
        every trace (the fit_predict_GP seam passes it explicitly); default self.seed + N_iter (gpet.py:874)."""
        B, mm, R = self.B, self.mmax, 13
        t_prep = time.perf_counter()
        Xs = np.zeros((B, mm)); yt = np.zeros((B, mm)); ws = np.zeros((B, mm)); ms = np.zeros(B, dtype=np.int32)
        stats = np.zeros((B, 6))                     # y_m, y_s, X_m, X_s, tm, ts
        x0 = np.zeros((B, R, 3))
        lo, hi = _gp_host.FINAL_BOUNDS[:, 0].copy(), _gp_host.FINAL_BOUNDS[:, 1].copy()
        tx, ty, tw, tm_ = self._training_sets()
        # standardisation (gpet.py:235-238, sklearn_gpr.py:229-234) per group of equal training-set size: numpy reduces
        # every row of an exact-length 2-D block with the same pairwise sum it uses for a 1-D array => same bits
        for k in np.unique(tm_):
            rows = np.flatnonzero(tm_ == k)
            k = int(k)
            X = tx[rows, :k].astype(np.float64)
            y = ty[rows, :k].copy()
            y_m, y_s = np.mean(y, axis=1), np.std(y, axis=1)
            y = (y - y_m[:, None]) / y_s[:, None]
            X_m, X_s = np.mean(X, axis=1), np.std(X, axis=1)
            X = (X - X_m[:, None]) / X_s[:, None]
            tm, ts = np.mean(y, axis=1), np.std(y, axis=1)
            ts = np.where(ts < 10 * np.finfo(np.float64).eps, 1.0, ts)
            Xs[rows, :k], yt[rows, :k], ws[rows, :k] = X, (y - tm[:, None]) / ts[:, None], tw[rows, :k]
            ms[rows] = k
            stats[rows] = np.stack([y_m, y_s, X_m, X_s, tm, ts], axis=1)
        starts = {}
        for sd in np.unique(self.n_iter):
            rng = np.random.RandomState(self.seed + int(sd) if seed is None else seed)   # sklearn_gpr.py:205, gpet.py:874
            t0 = np.empty((R, 3))
            t0[0] = np.log(np.array([5.0, 5.0, float(self.noise_y)]))            # gpet.py:244-245
            for r in range(1, R):
                t0[r] = rng.uniform(lo, hi)                                        # sklearn_gpr.py:285
            starts[int(sd)] = t0
        for b in range(B):
            x0[b] = starts[int(self.n_iter[b])]
        self.host_ms["fit_prep"] = self.host_ms.get("fit_prep", 0.0) + 1e3 * (time.perf_counter() - t_prep)
        return dict(Xs=Xs, yt=yt, ws=ws, ms=ms, stats=stats, x0=x0, xc=tx.astype(np.int32))

    def final_fit_all(self):
        """Converged branch for every trace of this batch at once; see final_fit_group."""
        return final_fit_group([self])[0]

    def trace(self):
        """Runs every trace to convergence. Returns (edge_traces int[B, n, 2] (y, x), list of (lo, hi))."""
        self.run_loop()
        if self.final_fit_mode == "device":
            edges, creds, self.final_info = self.final_fit_all()
            return edges, creds
        edges, creds = [], []
        for b in range(self.B):
            e, c, _ = self.final_fit(b)
            edges.append(e)
            creds.append(c)
        return np.stack(edges), creds


def fit_driver():
    """Where the L-BFGS-B state machines of the final fit run: "device" (default: gpet_lbfgsb_* kernels in lock step
    with the objective kernel, no host arithmetic) or "host" (scipy's setulb in worker processes; GPET_FIT_DRIVER=host)."""
    v = os.environ.get("GPET_FIT_DRIVER", "device").lower()
    if v not in ("device", "host"):
        raise GpetError(f"GPET_FIT_DRIVER={v!r}: expected 'device' or 'host'")
    return v


def _lbfgsb_device(x0, lo, hi, trace_of, dev, stage, n_eval, lml_args):
    """E L-BFGS-B runs advanced on the device: one round = gpet_lbfgsb_advance_f64 (every run that got its objective
    value moves to its next evaluation point or ends) + gpet_lml_f64 over the E slots (ended runs are skipped);
    gpet_fit_rounds_f64 queues several rounds per call. The host only reads three counters, one call late.
    Returns (x [E, 3], f [E], nfev [E], rounds) like LbfgsbPool.minimize_many."""
    dX, dy, dw, dxc, dm, mm, kind = lml_args
    E = x0.shape[0]
    big = mm > MAX_TRAIN      # kernel matrices in HBM (gpet_dense.cu): gpet_fit_rounds_big_f64 with a caller-owned workspace
    if big:
        free_b, _ = torch.cuda.mem_get_info(dev)
        one = int(query("gpet_lml_big_workspace_bytes", 1, mm))
        # all evaluation slots at once when they fit in 45 % of the free memory (a streamed run keeps the loop buffers of the
        # next batch next to this fit), otherwise the objective walks over the slots in chunks
        work_bytes = min(int(query("gpet_lml_big_workspace_bytes", E, mm)), max(one, int(0.45 * free_b)))
        d_lml_work = torch.empty(work_bytes, dtype=torch.uint8, device=dev)
    lib = _cabi.load()
    nd, ni = int(lib.gpet_lbfgsb_state_doubles()), int(lib.gpet_lbfgsb_state_ints())
    f64 = dict(dtype=torch.float64, device=dev)
    i32 = dict(dtype=torch.int32, device=dev)
    d_state = torch.empty((nd, E), **f64)
    i_state = torch.empty((ni, E), **i32)
    d_x0 = torch.from_numpy(np.ascontiguousarray(x0)).to(dev)
    d_lo = torch.from_numpy(np.ascontiguousarray(lo)).to(dev)
    d_hi = torch.from_numpy(np.ascontiguousarray(hi)).to(dev)
    d_tr = torch.from_numpy(np.ascontiguousarray(trace_of)).to(dev)
    d_theta = torch.zeros((E, 3), **f64)
    d_f = torch.zeros((E,), **f64)
    d_g = torch.zeros((E, 3), **f64)
    d_ev = torch.full((E,), -1, **i32)
    d_cnt = torch.zeros((3,), **i32)        # waiting runs after the last round / evaluations / rounds with work
    R = int(os.environ.get("GPET_FIT_ROUNDS_PER_CALL", "8"))
    h_cnt = [torch.zeros((3,), dtype=torch.int32).pin_memory() for _ in range(2)]
    events = [None, None]
    stage("lbfgsb", "gpet_lbfgsb_init_f64", ptr(d_state), ptr(i_state), E, ptr(d_x0), ptr(d_lo), ptr(d_hi), _stream())
    k = 0
    while True:
        slot = k & 1
        # R rounds [advance -> objective] per call; the counters of call k are read after call k + 1 has been queued,
        # so the device never waits for the host (the rounds queued after the last run ended are empty)
        if big:
            stage("lml", "gpet_fit_rounds_big_f64", ptr(dX), ptr(dy), ptr(dw), ptr(dm), mm, kind, _gp_host.GP_ALPHA,
                  ptr(d_state), ptr(i_state), E, 1 if k == 0 else 0, R, ptr(d_tr), ptr(d_f), ptr(d_g), ptr(d_theta),
                  ptr(d_ev), ptr(d_cnt), ptr(h_cnt[slot]), ptr(d_lml_work), work_bytes, _stream())
        else:
            stage("lml", "gpet_fit_rounds_f64", ptr(dX), ptr(dy), ptr(dw), ptr(dxc), ptr(dm), mm, kind, _gp_host.GP_ALPHA,
                  ptr(d_state), ptr(i_state), E, 1 if k == 0 else 0, R, ptr(d_tr), ptr(d_f), ptr(d_g), ptr(d_theta),
                  ptr(d_ev), ptr(d_cnt), ptr(h_cnt[slot]), _stream())
        ev = torch.cuda.Event()
        ev.record()
        events[slot] = ev
        k += 1
        if k >= 2:
            events[slot ^ 1].synchronize()
            if int(h_cnt[slot ^ 1][0].item()) == 0:
                break
        if k * R > 4 * _lbfgs_worker.MAXFUN:
            raise GpetError("device L-BFGS-B did not terminate")
    events[(k - 1) & 1].synchronize()
    last = h_cnt[(k - 1) & 1]
    n_eval[0] += int(last[1].item())
    rounds = int(last[2].item())
    launches = 1 + 2 * k * R
    d_xs = torch.empty((E, 3), **f64)
    d_fs = torch.empty((E,), **f64)
    d_nf = torch.empty((E,), **i32)
    d_task = torch.empty((E,), **i32)
    stage("lbfgsb", "gpet_lbfgsb_result_f64", ptr(d_state), ptr(i_state), E, ptr(d_xs), ptr(d_fs), ptr(d_nf), ptr(d_task),
          _stream())
    n_eval[1] += launches + 1
    return d_xs.cpu().numpy(), d_fs.cpu().numpy(), d_nf.cpu().numpy().astype(np.int64), rounds


def _fit_core(arr, kind, dev, stage):
    """Device + worker-pool part of the final fit on plain arrays.
    arr: dict(Xs, yt, ws [B, mm], ms [B], stats [B, 6], x0 [B, 13, 3], xq [B, n]).  stage(name, cabi_name, *args)
    launches one C-ABI call.  Returns dict(theta [B, 3], nfev [B, 13], rounds, lml_evals, launches, mean, sd, status)."""
    Xs, yt, ws, ms, stats, x0, xq = (arr[k] for k in ("Xs", "yt", "ws", "ms", "stats", "x0", "xq"))
    B, mm = Xs.shape
    n = xq.shape[1]
    R = x0.shape[1]
    lo, hi = _gp_host.FINAL_BOUNDS[:, 0].copy(), _gp_host.FINAL_BOUNDS[:, 1].copy()
    f64 = dict(dtype=torch.float64, device=dev)
    dX, dy, dw = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (Xs, yt, ws))
    dm = torch.from_numpy(np.ascontiguousarray(ms)).to(dev)
    dxc = torch.from_numpy(np.ascontiguousarray(arr["xc"])).to(dev) if arr.get("xc") is not None else None
    E = B * R
    trace_of = np.repeat(np.arange(B, dtype=np.int32), R)
    G = int(os.environ.get("GPET_FIT_GROUPS", "2"))      # worker groups evaluated alternately
    d_theta = [torch.empty((E, 3), **f64) for _ in range(G)]
    d_tr = [torch.empty((E,), dtype=torch.int32, device=dev) for _ in range(G)]
    d_fg = [torch.empty((E, 4), **f64) for _ in range(G)]
    h_theta = [torch.empty((E, 3), dtype=torch.float64).pin_memory() for _ in range(G)]
    h_tr = [torch.empty((E,), dtype=torch.int32).pin_memory() for _ in range(G)]
    h_fg = [torch.empty((E, 4), dtype=torch.float64).pin_memory() for _ in range(G)]
    n_eval = [0, 0]

    big = mm > MAX_TRAIN        # training sets beyond the shared-memory kernels: the `_big` entry points (gpet_dense.cu)
    if big:
        free_b, _ = torch.cuda.mem_get_info(dev)
        one = int(query("gpet_lml_big_workspace_bytes", 1, mm))
        lml_bytes = min(int(query("gpet_lml_big_workspace_bytes", E, mm)), max(one, int(0.4 * free_b)))
        lml_work = [None]       # allocated by the host-driven loop only (the device driver owns its own)

    def submit(gi, ids, thetas):
        k = ids.shape[0]
        h_theta[gi][:k].copy_(torch.from_numpy(np.ascontiguousarray(thetas)))
        h_tr[gi][:k].copy_(torch.from_numpy(trace_of[ids]))
        d_theta[gi][:k].copy_(h_theta[gi][:k], non_blocking=True)
        d_tr[gi][:k].copy_(h_tr[gi][:k], non_blocking=True)
        # f -> column 0, g -> columns 1..3 of one buffer (a single device->host copy per evaluation batch)
        if big:
            if lml_work[0] is None:
                lml_work[0] = torch.empty(lml_bytes, dtype=torch.uint8, device=dev)
            stage("lml", "gpet_lml_big_f64", ptr(dX), ptr(dy), ptr(dw), ptr(dm), mm, ptr(d_tr[gi]), ptr(d_theta[gi]), k, kind,
                  _gp_host.GP_ALPHA, ptr(d_fg[gi]), d_fg[gi].data_ptr() + E * 8, ptr(lml_work[0]), lml_bytes, _stream())
        else:
            stage("lml", "gpet_lml_f64", ptr(dX), ptr(dy), ptr(dw), ptr(dxc), ptr(dm), mm, ptr(d_tr[gi]), ptr(d_theta[gi]),
                  k, kind, _gp_host.GP_ALPHA, ptr(d_fg[gi]), d_fg[gi].data_ptr() + E * 8, _stream())
        hf, df = h_fg[gi].view(-1), d_fg[gi].view(-1)
        hf[:k].copy_(df[:k], non_blocking=True)
        hf[E:E + 3 * k].copy_(df[E:E + 3 * k], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        n_eval[0] += k
        n_eval[1] += 1
        return gi, k, ev

    def wait(handle):
        gi, k, ev = handle
        ev.synchronize()
        flat = h_fg[gi].numpy().reshape(-1)
        return flat[:k].copy(), flat[E:E + 3 * k].reshape(k, 3).copy()

    if fit_driver() == "device":
        xs, fs, nfev, rounds = _lbfgsb_device(x0.reshape(E, 3), lo, hi, trace_of, dev, stage, n_eval,
                                              (dX, dy, dw, dxc, dm, mm, kind))
    else:
        xs, fs, nfev, rounds = fit_pool(E).minimize_many(x0.reshape(E, 3), lo, hi, submit, wait, n_groups=G)
    fs = fs.reshape(B, R)
    best = np.argmin(fs, axis=1)                                             # first minimum, like np.argmin
    theta = xs.reshape(B, R, 3)[np.arange(B), best]
    d_best = torch.from_numpy(np.ascontiguousarray(theta)).to(dev)
    d_xq = torch.from_numpy(np.ascontiguousarray(xq)).to(dev)
    d_tmts = torch.from_numpy(np.ascontiguousarray(stats[:, 4:6])).to(dev)
    d_mean = torch.empty((B, n), **f64)
    d_sd = torch.empty((B, n), **f64)
    d_st = torch.empty((B,), dtype=torch.int32, device=dev)
    if big:
        lml_work[0] = None
        # traces in chunks the workspace holds (L: ld^2, K*^T: ld x n per trace)
        per = int(query("gpet_final_predict_big_workspace_bytes", 1, mm, n))
        free_b, _ = torch.cuda.mem_get_info(dev)
        Tc = max(1, min(B, int(0.5 * free_b) // per))
        work = torch.empty(int(query("gpet_final_predict_big_workspace_bytes", Tc, mm, n)), dtype=torch.uint8, device=dev)
        for a in range(0, B, Tc):
            b = min(B, a + Tc)
            stage("final_predict", "gpet_final_predict_big_f64", ptr(dX[a:b]), ptr(dy[a:b]), ptr(dw[a:b]), ptr(dm[a:b]), mm,
                  b - a, ptr(d_best[a:b]), kind, _gp_host.GP_ALPHA, ptr(d_xq[a:b]), n, ptr(d_tmts[a:b]), ptr(d_mean[a:b]),
                  ptr(d_sd[a:b]), ptr(d_st[a:b]), ptr(work), _stream())
    else:
        stage("final_predict", "gpet_final_predict_f64", ptr(dX), ptr(dy), ptr(dw), ptr(dm), mm, B, ptr(d_best), kind,
              _gp_host.GP_ALPHA, ptr(d_xq), n, ptr(d_tmts), ptr(d_mean), ptr(d_sd), ptr(d_st), _stream())
    return dict(theta=theta, nfev=nfev.reshape(B, R), rounds=rounds, lml_evals=n_eval[0], launches=n_eval[1] + 1,
                mean=d_mean.cpu().numpy(), sd=d_sd.cpu().numpy(), status=d_st.cpu().numpy())


def final_fit_group(tbs, seed=None):
    """Converged branch (gpet.py:232-248, 263-266, 874-886) for every trace of the TraceBatch objects `tbs` (same
    configuration) in ONE lock-step optimisation: the 13 L-BFGS-B runs per trace (sklearn_gpr.py:254-295) advance on
    the device (gpet_lbfgsb_*; GPET_FIT_DRIVER=host: scipy's setulb in worker processes), their objective
    -(log marginal likelihood, gradient) is evaluated in batches by gpet_lml_f64, the final predictive mean/std by
    gpet_final_predict_f64.
    Returns one (edges int[B, n, 2], creds list of (lo, hi), info dict) per TraceBatch."""
    t0 = tbs[0]
    n, mm, dev = t0.n, t0.mmax, t0.dev
    if any((tb.n, tb.mmax, tb.ktype, tb.nu, tb.noise_y) != (n, mm, t0.ktype, t0.nu, t0.noise_y) for tb in tbs):
        raise GpetError("final_fit_group: the batches must share their configuration")
    kind = _KIND.get((t0.ktype, None if t0.ktype == "RBF" else float(t0.nu)))
    if kind is None:
        raise GpetError(f"final fit on the device supports RBF and Matern nu in (0.5, 1.5, 2.5), not nu={t0.nu}")
    parts = [tb._fit_inputs(seed) for tb in tbs]
    arr = {k: np.concatenate([p[k] for p in parts]) for k in ("Xs", "yt", "ws", "ms", "stats", "x0", "xc")}
    x_grid = np.concatenate([np.broadcast_to(tb.x_grid[None, :], (tb.B, n)) for tb in tbs])
    stats = arr["stats"]
    arr["xq"] = (x_grid - stats[:, 2:3]) / stats[:, 3:4]                     # gpet.py:264
    t_fit = time.perf_counter()
    res = _fit_core(arr, kind, dev, t0._stage)
    t0.host_ms["fit_rounds"] = t0.host_ms.get("fit_rounds", 0.0) + 1e3 * (time.perf_counter() - t_fit)
    t0.kernel_launches += res["launches"]
    if np.any(res["status"] != 0):
        raise np.linalg.LinAlgError("Cholesky failed at the optimised hyper-parameters (sklearn_gpr.py:306-314)")
    mean, sd, theta = res["mean"], res["sd"], res["theta"]
    B = mean.shape[0]
    y_mean = stats[:, 1:2] * mean + stats[:, 0:1]                            # gpet.py:266 (std NOT rescaled)
    edges = np.empty((B, n, 2), dtype=int)
    edges[:, :, 0] = np.rint(y_mean).astype(int)                             # gpet.py:885-886
    edges[:, :, 1] = x_grid
    out, o = [], 0
    for i, tb in enumerate(tbs):
        sl = slice(o, o + tb.B)
        creds = [(y_mean[b] - 1.96 * sd[b], y_mean[b] + 1.96 * sd[b]) for b in range(o, o + tb.B)]
        info = dict(theta=theta[sl], nfev=res["nfev"][sl], rounds=res["rounds"] if i == 0 else 0,
                    lml_evals=res["lml_evals"] if i == 0 else 0, y_mean=y_mean[sl], y_std=sd[sl])
        out.append((edges[sl], creds, info))
        o += tb.B
    return out


_fit_executor = None
_fit_stream = None


def _fit_resources():
    """One background thread (fits are serialised: they share the worker processes) and its high-priority stream."""
    global _fit_executor, _fit_stream
    import concurrent.futures
    if _fit_executor is None:
        _fit_executor = concurrent.futures.ThreadPoolExecutor(max_workers=1, thread_name_prefix="gpet-fit")
        # high priority: the small objective kernels of a round must not queue behind the multi-millisecond loop kernels
        _fit_stream = torch.cuda.Stream(priority=int(os.environ.get("GPET_FIT_PRIORITY", "-1")))
    return _fit_executor, _fit_stream


class PipelinedResult:
    """Handle returned by trace_pipelined(..., wait=False): the tracing loops are done, the final fits may still be
    running in the background thread. result() waits for them and returns (edges, creds) in batch order. Only the
    results are kept alive here; the loop buffers of the batches were released when their loops ended."""

    def __init__(self, n_batches, futures, results):
        self.n_batches, self._futures, self._results = n_batches, futures, results

    def done(self):
        return all(f.done() for f in self._futures)

    def result(self):
        for f in self._futures:
            f.result()
        out = [self._results[i] for i in range(self.n_batches)]
        return np.concatenate([e for e, _ in out]), [c for _, cs in out for c in cs]


_outstanding = []        # fit futures of earlier trace_pipelined(wait=False) calls (back-pressure)


def trace_pipelined(batches, window=2, fit_merge=2, wait=True, own_streams=False, max_pending=4, release=True):
    """Runs several TraceBatch objects (sub-batches of one workload) to completion with host and device work
    overlapped; returns (edges int[sum B, n, 2], creds list) in batch order, like TraceBatch.trace().

    * At most `window` sub-batches are inside the while-loop (gpet.py:829-870) at a time. Their iterations alternate:
      while the host waits for the control block of one sub-batch, the kernels of the other one are already queued.
    * A sub-batch that has converged releases its loop buffers (release=True: TraceBatch.release_loop_buffers - device
      memory does not grow with the number of workloads in flight) and hands its final hyper-parameter fit
      (gpet.py:232-248) to a background thread with its own high-priority CUDA stream; the L-BFGS-B rounds (a chain of
      small kernels; with GPET_FIT_DRIVER=host scipy's setulb in worker processes) then overlap with the loop kernels of
      the following sub-batches. `fit_merge` converged sub-batches are fitted together (one larger lock-step
      optimisation instead of several small ones).
      (A helper PROCESS for the fit was tried and measured slower: across processes the GPU is time-sliced and the
      stream priority that lets the small objective kernels overtake the loop kernels does not apply.)
    * own_streams: every sub-batch launches on a CUDA stream of its own (TraceBatch.use_own_stream), so the
      sub-batches of the window also overlap on the device.
    * wait=False returns a PipelinedResult as soon as the loops are done: a caller that streams workloads (bench.py)
      starts the loops of the next workload while the last fits of this one are still running, and collects later.
      Back-pressure: a call first waits until at most `max_pending` fit jobs of earlier calls are unfinished.
    """
    if not batches:
        return np.zeros((0, 0, 2), dtype=int), []
    batches = batches if isinstance(batches, list) else list(batches)   # entries may be TraceBatch factories (replaced in place)
    if any((not callable(tb)) and tb.final_fit_mode != "device" for tb in batches):
        out = [(tb() if callable(tb) else tb).trace() for tb in batches]
        return np.concatenate([e for e, _ in out]), [c for _, cs in out for c in cs]
    pool, fit_stream = _fit_resources()
    loop_prio = os.environ.get("GPET_LOOP_PRIORITY")     # experiment knob: loops on their own stream of this priority
    _outstanding[:] = [f for f in _outstanding if not f.done()]
    while len(_outstanding) > max(0, max_pending):
        _outstanding.pop(0).result()
    results = {}

    def fit(tbs, ids):
        with torch.cuda.stream(fit_stream):
            for i, tb, (edges, creds, info) in zip(ids, tbs, final_fit_group(tbs)):
                tb.final_info = info
                results[i] = (edges, creds)

    futures = []
    todo = list(range(len(batches)))
    inside, finished = [], []

    def flush():
        # sub-batches that converged while the same window was open are fitted together
        if finished:
            ids = list(finished)
            futures.append(pool.submit(fit, [batches[i] for i in ids], ids))
            finished.clear()

    def retire(i):
        if release:
            batches[i].release_loop_buffers()
        finished.append(i)

    def admit():
        while todo and len(inside) < window:
            i = todo.pop(0)
            if callable(batches[i]):       # factory: the sub-batch (its upload, gradient image, device state) is
                batches[i] = batches[i]()  # only created now, so host->device copies overlap earlier sub-batches
            if loop_prio is not None:
                batches[i].use_own_stream(priority=int(loop_prio))
            elif own_streams and window > 1:
                batches[i].use_own_stream()
            if batches[i].step_launch():
                inside.append(i)
            else:
                retire(i)

    admit()
    while inside:
        i = inside.pop(0)
        tb = batches[i]
        tb.step_finish()
        if tb.step_launch():
            inside.append(i)
        else:
            retire(i)
            admit()
            if len(finished) >= fit_merge or not inside:
                flush()
    flush()
    _outstanding.extend(futures)
    handle = PipelinedResult(len(batches), futures, results)
    handle.batches = batches            # host-side statistics only: the device buffers of the loops are gone
    return handle.result() if wait else handle


def trace_stream(factories, prefetch=0, max_pending=1, fit_merge=1):
    """Throughput front end for a STREAM of independent batches (bench.py: the steps of a run; a production job: the
    batches of a long list of images). `factories` yields callables that build one TraceBatch each (their host->device
    copies, gradient stencil and constructor kernels included). Generator: yields (edges int[B, n, 2], creds, batch) per
    batch, in order.

    Three things overlap:
      * the while-loop (gpet.py:829-870) of batch i - on the batch's own CUDA stream, driven by the calling thread;
      * the construction of batch i+1 .. i+prefetch (stencil, normalise, gradient KDE, transposed copy, uploads) - in a
        builder thread on a side stream (a constructor makes small synchronous copies; in the calling thread they would
        stall the launches of the running loop). Measured on B200 (cfg 5 shard): the device is already saturated by the
        loop and the fit streams, a third stream only adds contention (3.0-3.7 k traces/s, erratic, against a steady 3.85 k)
        - so the default is prefetch=0: batches are built in the calling thread between two loops, and only their
        host->device copies are started a batch ahead (a factory may offer .prefetch() for that);
      * the final hyper-parameter fits (gpet.py:232-248) of earlier batches - background thread, high-priority stream;
        at most `max_pending` fit jobs are outstanding before the generator hands out the oldest result, and a
        converged batch releases its loop buffers first, so device memory does not grow with the length of the stream.
    """
    import collections
    import concurrent.futures
    pool, fit_stream = _fit_resources()
    it = iter(factories)
    built = collections.deque()            # futures of batches under construction / constructed
    waiting = collections.deque()          # (future, batches, results) of fits in flight
    builder = concurrent.futures.ThreadPoolExecutor(max_workers=1, thread_name_prefix="gpet-build") if prefetch > 0 else None
    side = torch.cuda.Stream() if prefetch > 0 else None
    dev = torch.cuda.current_device()

    def construct(f):
        torch.cuda.set_device(dev)
        with torch.cuda.stream(side):
            tb = f() if callable(f) else f
            tb.use_own_stream()
        return tb

    ahead = collections.deque()            # factories whose input copies have been started (prefetch == 0)

    def next_factory():
        return ahead.popleft() if ahead else next(it, None)

    def start_copies():
        # prefetch == 0: a factory may offer .prefetch() - a cheap, asynchronous start of its host->device copies - which
        # is called one batch ahead, so that the copies overlap the loop of the current batch
        if not ahead:
            f = next(it, None)
            if f is not None:
                if hasattr(f, "prefetch"):
                    f.prefetch()
                ahead.append(f)

    def build_one():
        f = next_factory()
        if f is None:
            return False
        if builder is None:
            fut = concurrent.futures.Future()
            fut.set_result(f() if callable(f) else f)
        else:
            fut = builder.submit(construct, f)
        built.append(fut)
        return True

    def fit(tbs, out):
        with torch.cuda.stream(fit_stream):
            for tb, (edges, creds, info) in zip(tbs, final_fit_group(tbs)):
                tb.final_info = info
                out.append((edges, creds, tb))

    def drain(limit):
        while len(waiting) > limit:
            fut, tbs, out = waiting.popleft()
            fut.result()
            for r in out:
                yield r

    group = []
    try:
        build_one()
        while built:
            tb = built.popleft().result()
            if tb.final_fit_mode != "device":
                edges, creds = tb.trace()
                yield edges, creds, tb
                continue
            if os.environ.get("GPET_LOOP_PRIORITY") is not None:      # experiment knob: the loop on its own stream of this priority
                tb.use_own_stream(priority=int(os.environ["GPET_LOOP_PRIORITY"]))
            more = tb.step_launch()
            while len(built) < max(prefetch, 0) and build_one():
                pass
            if prefetch <= 0:
                start_copies()
            while more:
                tb.step_finish()
                more = tb.step_launch()
            tb.release_loop_buffers()
            if not built:
                build_one()                    # prefetch == 0: the next batch is built here, between two loops
            group.append(tb)
            if len(group) >= fit_merge or not built:
                out = []
                waiting.append((pool.submit(fit, list(group), out), list(group), out))
                group = []
            yield from drain(max_pending)
        yield from drain(0)
    finally:
        if builder is not None:
            builder.shutdown(wait=True)
