#!/usr/bin/env python
"""bench.py - edges traced / s on BASELINE.json config 5 (a stream of independent 500x500 synthetic test images
traced with the README parameters: RBF sigma_f=75 ls=20, N_samples=1000, delta_x=5, keep_ratio=0.1,
pixel_thresh=5, same seed), sharded over the GPUs of one box (whole traces per rank, no data-path collective).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--traces B] [--impl ours|reference]

A step = one pass of the hot path over one batch of B images per GPU: gradient stencil -> GP_Edge_Tracing set-up
(normalise, gradient KDE) -> recursive-Bayesian loop (posterior, factor, sample, score, top-k, density, select)
-> final hyper-parameter fit -> integer edge_pred + credible interval.
  value : traces / s with the images already resident in HBM when the timed region starts;
  e2e   : traces / s through the public API with HOST buffers (pinned): host->device copy of the images and
          device->host read of edge_pred/credint inside the timed region.
The reference arm (--impl reference) and `cpu_baseline` time the oracle port of the reference's own numpy/scipy
path (per-curve Python loop exactly like gpet.py:437-440) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
os.environ.setdefault("OMP_NUM_THREADS", "1")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

KERNEL_OPTIONS = {"kernel": "RBF", "sigma_f": 75, "length_scale": 20}
TRACE_KW = dict(kernel_options=KERNEL_OPTIONS, noise_y=1, N_samples=1000, score_thresh=1, delta_x=5, keep_ratio=0.1,
                pixel_thresh=5, seed=1, fix_endpoints=True)
IMG = 500
# dram__bytes_read.sum + dram__bytes_write.sum of one full-shard launch (1250 traces x 1000 curves) of the scoring kernel
# from the ncu --set full capture profiles/r01_score_streamN_ncu.csv (6.260 GB read: Y 5.0 GB + gradient columns
# 1.25 GB; 15 MB written)
NCU_TRAFFIC_BYTES_PER_LAUNCH = 6.275e9

WORKLOAD = ("cfg5 shard: B independent 500x500 construct_test_img traces per GPU per step "
            "(RBF sigma_f=75 ls=20, N_samples=1000, delta_x=5, keep_ratio=0.1, pixel_thresh=5, seed=1)")


def image_params(i):
    """SURVEY 8(d) cfg 5: image i uses amplitude in {100..300}, curvature in {2,3,4,5}, noise seed i."""
    return dict(amplitude=100 + (37 * i) % 201, curvature=2 + i % 4, noise_seed=i + 1)


def make_image(i):
    from gaussian_process_edge_trace_b200 import gpet_utils
    p = image_params(i)
    img, edge = gpet_utils.construct_test_img((IMG, IMG), p["amplitude"], p["curvature"], 0.05, "sinusoidal", 0.3,
                                               gaps=True, noise_seed=p["noise_seed"])
    return img, edge[[0, -1], :][:, [1, 0]]


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path on host cores
# ------------------------------------------------------------------------------------------------------------
def ref_kind():
    """'reference' when the unmodified reference package was staged under oracle/_ref (by __graft_entry__.build() in the
    build container; it travels to the GPU box), else 'port' (the oracle restatement). GPET_REF_KIND overrides."""
    env = os.environ.get("GPET_REF_KIND")
    if env in ("reference", "port"):
        return env
    return "reference" if os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "gp_edge_tracing")) else "port"


def _cpu_trace(i):
    """One trace of the workload with the oracle port (per-curve Python loop exactly like gpet.py:437-440)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gpet_oracle as O
    p = image_params(i)
    img, edge = O.construct_test_img((IMG, IMG), p["amplitude"], p["curvature"], 0.05, "sinusoidal", 0.3, gaps=True,
                                     noise_seed=p["noise_seed"])
    t0 = time.time()
    grad = O.comp_grad_img(img, O.kernel_builder((11, 5)))
    tr = O.OracleTracer(edge[[0, -1], :][:, [1, 0]], grad, return_std=True, loop_costs=True, **TRACE_KW)
    tr()
    return time.time() - t0, len(tr.record) * TRACE_KW["N_samples"]


def _cpu_trace_ref(i):
    """One trace of the workload with the UNMODIFIED reference (gp_edge_tracing.gpet_utils.comp_grad_img +
    gp_edge_tracing.gpet.GP_Edge_Tracing.__call__) imported through oracle/ref_harness.py (shims for the packages this
    image lacks: matplotlib stub, KDEpy / skimage stand-ins, scipy.simps and sklearn._validate_data adapters)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gpet_oracle as O
    import ref_harness
    gpet, gu, _ = ref_harness.load_reference()
    p = image_params(i)
    img, edge = O.construct_test_img((IMG, IMG), p["amplitude"], p["curvature"], 0.05, "sinusoidal", 0.3, gaps=True,
                                     noise_seed=p["noise_seed"])
    t0 = time.time()
    grad = gu.comp_grad_img(img, gu.kernel_builder(size=(11, 5)))
    tr = gpet.GP_Edge_Tracing(edge[[0, -1], :][:, [1, 0]], grad, obs=np.array([]), return_std=True, **TRACE_KW)
    calls = [0]
    best = tr.get_best_curves

    def counted(y_samples):
        calls[0] += 1
        return best(y_samples)

    tr.get_best_curves = counted
    tr()
    return time.time() - t0, calls[0] * TRACE_KW["N_samples"]


def cpu_arm(n_workers, rounds, first_image=0, kind="port", budget_s=None):
    """Each round traces one image per worker process. Returns (traces/s over all rounds, per-round seconds, curves).
    budget_s: stop starting new rounds once the elapsed time exceeds it (at least one round runs)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    fn = _cpu_trace_ref if kind == "reference" else _cpu_trace
    per_round, curves = [], 0
    t_all = time.time()
    with ctx.Pool(n_workers) as pool:
        for r in range(rounds):
            if budget_s is not None and r > 0 and time.time() - t_all > budget_s:
                break
            t0 = time.time()
            out = pool.map(fn, range(first_image + r * n_workers, first_image + (r + 1) * n_workers))
            per_round.append(time.time() - t0)
            curves += sum(o[1] for o in out)
    return n_workers * len(per_round) / sum(per_round), per_round, curves


def host_cores():
    """Cores this process may run on (affinity mask when the platform has one, else the machine's count)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


CPU_SAMPLE = {"reference": "the UNMODIFIED reference (gp_edge_tracing comp_grad_img + GP_Edge_Tracing.__call__, staged under "
                           "oracle/_ref, imported through oracle/ref_harness.py: stand-ins for the absent KDEpy / "
                           "scikit-image, matplotlib stub, canonical SVD signs)",
              "port": "oracle port of the reference numpy/scipy path incl. its per-curve Python loop"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    kind = ref_kind()
    w, k = max(args.warmup, 0), max(args.steps, 1)
    # warm-up steps of a CPU arm only page the interpreter and the libraries in; they stop early when they alone would
    # take more than GPET_REF_WARMUP_BUDGET_S (default 40 s) so that the whole run ends within a few minutes
    w_run = 0
    if w:
        _, pr, _ = cpu_arm(cores, w, first_image=10_000, kind=kind,
                           budget_s=float(os.environ.get("GPET_REF_WARMUP_BUDGET_S", "40")))
        w_run = len(pr)
    tps, per_round, curves = cpu_arm(cores, k, kind=kind)
    ms = 1e3 * sum(per_round) / k
    line = {
        "impl": "reference", "metric": "edges_traced_per_sec", "value": tps, "unit": "traces/s", "n_gpus": args.gpus,
        "steps": k, "warmup": w, "warmup_steps_run": w_run, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": {"workload": WORKLOAD, "traces_per_step": cores},
        "curves_scored_per_sec": curves / sum(per_round),
        "cpu_baseline": {"value": tps, "unit": "traces/s", "cores": cores, "kind": kind,
                         "sample": f"{cores} traces per step (one per host core): " + CPU_SAMPLE[kind]},
        "e2e": {"value": tps, "unit": "traces/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
            "clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "reasons": reasons,
                "power_w_max": max(float(r[2]) for r in rows), "samples": len(rows)}


def _oracle_check(job):
    """Worker of the in-bench parity check (spawned process, no CUDA): the oracle traces one image of the benched shard
    with the factor the GPU produced injected per iteration; returns what it got."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gpet_oracle as O
    grad, init, factors = job
    orc = O.OracleTracer(init, grad, factor_fn=lambda cov, it: factors[it], return_std=True, **TRACE_KW)
    edge, cred = orc()
    return edge, np.stack(cred), [r["fobs"] for r in orc.record]


def parity_check(sel, inits, d_imgs, kern, benched_edges, benched_creds, TraceBatch, gpet_utils):
    """The benched path against the oracle on the images `sel` of the shard: (1) the edges / credible intervals the
    timed (pipelined, compacting, merged-fit) path returned for them equal a plain recorded TraceBatch re-trace;
    (2) the oracle, fed the factor of every GPU iteration, selects the same observation sets in every iteration and
    returns the same integer edge_pred; credible interval within 1e-6 relative."""
    import multiprocessing as mp
    import torch
    grad = gpet_utils.comp_grad_img(d_imgs[torch.as_tensor(sel, device=d_imgs.device)], kern, return_tensor=True)
    tb = TraceBatch(inits[sel], grad, record=True, **TRACE_KW)
    edges, creds = tb.trace()
    rec = tb.record
    grad_h = grad.cpu().numpy()
    jobs = []
    for j in range(len(sel)):
        factors = [r["A"][j] for r in rec if r["active"][j]]
        jobs.append((grad_h[j], inits[sel[j]], factors))
    with mp.get_context("spawn").Pool(min(len(sel), host_cores())) as pool:
        out = pool.map(_oracle_check, jobs)
    res = {"traces": len(sel), "images": [int(i) for i in sel], "edge_pred_equal": True, "obs_sets_equal": True,
           "benched_equals_retrace": True, "credint_max_rel": 0.0}
    for j, (e_o, c_o, fobs_o) in enumerate(out):
        its = [r for r in rec if r["active"][j]]
        same_obs = len(its) == len(fobs_o) and all(np.array_equal(r["fobs"][j], f) for r, f in zip(its, fobs_o))
        res["obs_sets_equal"] &= bool(same_obs)
        res["edge_pred_equal"] &= bool(np.array_equal(edges[j], e_o))
        c_g = np.stack(creds[j])
        res["credint_max_rel"] = max(res["credint_max_rel"], float(np.max(np.abs(c_g - c_o) / np.maximum(1.0, np.abs(c_o)))))
        b_e, b_c = benched_edges[sel[j]], np.stack(benched_creds[sel[j]])
        res["benched_equals_retrace"] &= bool(np.array_equal(b_e, edges[j]) and np.allclose(b_c, c_g, rtol=1e-9, atol=1e-9))
    res["ok"] = bool(res["edge_pred_equal"] and res["obs_sets_equal"] and res["benched_equals_retrace"]
                     and res["credint_max_rel"] <= 1e-6)
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 hot path has no CPU fallback")
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # before any CUDA context exists in this process (the worker pool is forked); N = 1 only
        cores = host_cores()
        kind = ref_kind()
        tps, per_round, curves = cpu_arm(cores, 1, kind=kind)
        cpu = {"value": tps, "unit": "traces/s", "cores": cores, "kind": kind,
               "curves_scored_per_sec": curves / sum(per_round),
               "sample": f"{cores} traces of the same workload (one per host core, {per_round[0]:.1f} s): " + CPU_SAMPLE[kind]}
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    __graft_entry__.build()
    from gaussian_process_edge_trace_b200 import TraceBatch, gpet_utils
    from gaussian_process_edge_trace_b200.engine import StageTimers, trace_stream
    dev = torch.device(f"cuda:{local}")
    B = args.traces
    kern = gpet_utils.kernel_builder((11, 5))

    # synthetic inputs of this rank's shard (distinct images across ranks), pinned host + resident device copies
    t_gen = time.time()
    cache = os.environ.get("GPET_BENCH_IMG_CACHE")          # tuning runs only: reuse the generated shard across invocations
    cache = f"{cache}.{rank}.{B}.npz" if cache else None
    if cache and os.path.exists(cache):
        z = np.load(cache)
        imgs, inits = z["imgs"], z["inits"]
    else:
        imgs = np.empty((B, IMG, IMG), dtype=np.float64)
        inits = np.empty((B, 2, 2), dtype=np.int64)
        for i in range(B):
            imgs[i], inits[i] = make_image(rank * B + i)
        if cache:
            np.savez(cache, imgs=imgs, inits=inits)
    h_imgs = torch.from_numpy(imgs).pin_memory()
    d_imgs = h_imgs.to(dev)
    t_gen = time.time() - t_gen
    timers = StageTimers()
    stats = {}

    copy_stream = torch.cuda.Stream()

    cuts = np.linspace(0, B, max(1, min(args.sub_batches, B)) + 1).astype(int)
    spans = list(zip(cuts[:-1], cuts[1:]))

    class BatchFactory:
        """Builds one TraceBatch over images a..b of the shard. e2e: prefetch() starts the host -> device copy of its
        images from pinned memory on a copy stream (non-blocking: it overlaps the loop of the batch before)."""

        def __init__(self, resident, a, b):
            self.resident, self.a, self.b, self.d, self.ev = resident, a, b, None, None

        def prefetch(self):
            if not self.resident and self.d is None:
                with torch.cuda.stream(copy_stream):
                    self.d = h_imgs[self.a:self.b].to(dev, non_blocking=True)
                    self.ev = torch.cuda.Event()
                    self.ev.record(copy_stream)

        def __call__(self):
            cur = torch.cuda.current_stream()
            if self.resident:
                d = d_imgs[self.a:self.b]
            else:
                self.prefetch()
                cur.wait_event(self.ev)
                d = self.d
                d.record_stream(cur)
                self.d = None
            grad = gpet_utils.comp_grad_img(d, kern, return_tensor=True)
            return TraceBatch(inits[self.a:self.b], grad, timers=timers, **TRACE_KW)

    def make_batch(resident, a, b):
        return BatchFactory(resident, a, b)

    def run_steps(resident, k):
        # the k steps are a stream of batches (engine.trace_stream): the loop of one batch runs while the next batch is
        # being built (e2e: while its images are being copied in) and while the final fits of the previous one finish in
        # the background; at most one fitted batch is uncollected and converged batches release their loop buffers, so
        # device memory does not grow with k.  Every result is read back before the function returns.
        facs = [make_batch(resident, a, b) for _ in range(k) for a, b in spans]
        edges, creds, tbs = [], [], []
        n_sub = len(spans)
        for e, c, tb in trace_stream(facs, prefetch=args.prefetch, max_pending=args.max_pending, fit_merge=args.fit_merge):
            edges.append(e)
            creds.extend(c)
            tbs.append(tb)
            if len(tbs) == n_sub:                 # one step complete
                stats["curves"] = sum(t.curves_scored for t in tbs)
                # stencil(3), normalise(3), grad KDE(5), transpose(1) per batch
                stats["launches"] = sum(t.kernel_launches + 3 + 3 + 5 + 1 for t in tbs)
                stats["iters"] = int(max(t.n_iter.max() for t in tbs))
                hm = {}
                for t in tbs:
                    for kk, v in t.host_ms.items():
                        hm[kk] = hm.get(kk, 0.0) + v
                stats["host_ms"] = {kk: round(v, 1) for kk, v in hm.items()}
                stats["fit"] = {kk: int(sum(t.final_info[kk] for t in tbs)) for kk in ("rounds", "lml_evals")}
                stats["edges"], stats["creds"] = np.concatenate(edges), creds
                edges, creds, tbs = [], [], []

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(resident, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_steps(resident, k)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item())

    # warm-up in the same streamed pattern as the timed region (the caching allocator and the fit stream are in their
    # steady state when the clock starts)
    if args.warmup > 0:
        run_steps(True, args.warmup)
        run_steps(False, 1)
    torch.cuda.synchronize()
    mem_warm = torch.cuda.max_memory_allocated()
    timers.reset()
    sampler = ClockSampler(local)
    sampler.start()
    ms_total = timed(True, args.steps)
    stage = timers.collect()
    launches_per_step = stats["launches"]
    curves_per_step = stats["curves"]
    ms_e2e = timed(False, args.steps)
    clocks = sampler.stop()
    mem_end = torch.cuda.max_memory_allocated()

    value = world * B * args.steps / (ms_total / 1e3)
    e2e = world * B * args.steps / (ms_e2e / 1e3)
    h2d = B * IMG * IMG * 8
    d2h = B * IMG * (2 * 8 + 2 * 8)        # edge_pred int64[n,2] + credint 2 x float64[n]

    # parity of the benched path (rank 0): the results of the last timed e2e step against the oracle
    parity = None
    if rank == 0 and args.parity > 0:
        sel = np.unique(np.linspace(0, B - 1, min(args.parity, B)).astype(int))
        parity = parity_check(sel, inits, d_imgs, kern, stats["edges"], stats["creds"], TraceBatch, gpet_utils)

    # rooflines, from CUDA events around every C-ABI call of the TIMED region (on the launching stream).
    # Headline: the scoring kernel, HBM bound, algorithmic bytes = 8 n + 8 per curve.
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    n, S = IMG, TRACE_KW["N_samples"]
    curves_total = curves_per_step * args.steps
    sc_ms, sc_n = stage.get("score", (0.0, 0))
    achieved = curves_total * (8 * n + 8) / (sc_ms * 1e-3) / 1e9 if sc_n else None
    roofline = {"kernel": "gpet::score_stream*_kernel (gpet_score_f64), every launch of the timed region",
                "bound": "hbm", "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None,
                "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH * (curves_total / sc_n) / (1250 * S) if sc_n else None,
                "traffic_note": "ncu --set full dram bytes of a 1250-trace launch (profiles/), scaled to the mean launch",
                "ms_per_launch": sc_ms / sc_n if sc_n else None, "launches": sc_n,
                "bytes_per_launch": curves_total * (8 * n + 8) / sc_n if sc_n else None}
    # secondary figure: the same kernel timed ALONE on one full-shard launch (B traces x S curves at iteration 8 of the
    # benched shard: no other stream on the GPU, no shrinking launches) - what the kernel itself reaches; the headline
    # `frac` above is what it gets inside the timed region, next to the final-fit stream and on partly converged batches
    if rank == 0 and args.dedicated:
        from gaussian_process_edge_trace_b200._cabi import call, ptr
        grad = gpet_utils.comp_grad_img(d_imgs, kern, return_tensor=True)
        tbd = TraceBatch(inits, grad, **TRACE_KW)
        for _ in range(8):
            tbd.step()
        nbd = min(B, tbd.Bc)
        stc = torch.cuda.current_stream().cuda_stream
        run = lambda: call("gpet_score_f64", ptr(tbd.curve_buffer()), ptr(tbd.gradT), None, nbd, n, S, IMG, IMG, tbd.x_st,
                           ptr(tbd.d_cost), stc)
        run()
        torch.cuda.synchronize()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record()
        for _ in range(5):
            run()
        d1.record()
        torch.cuda.synchronize()
        ms_d = d0.elapsed_time(d1) / 5
        gb_d = nbd * S * (8 * n + 8) / (ms_d * 1e-3) / 1e9
        roofline["dedicated_launch"] = {"traces": nbd, "ms": ms_d, "achieved": gb_d, "frac": gb_d / peak,
                                        "note": "one launch over the whole shard, nothing else running; Y (5 GB) larger than L2"}
        del tbd, grad
    rp = 76
    dmma_peak = 37.2        # TF/s, tools/ubench/dmma_peak.cu on B200 (MEASURED_PEAKS.json holds no fp64 figure)
    others = {}
    if "sample" in stage:
        ms_, k_ = stage["sample"]
        tf = 2.0 * curves_total * n * rp / (ms_ * 1e-3) / 1e12
        others["sample"] = {"kernel": "gpet::sample_rows_kernel", "bound": "tensor (fp64 DMMA)", "achieved": tf,
                            "peak": dmma_peak, "peak_source": "tools/ubench/dmma_peak.cu (builder measured)",
                            "unit": "TFLOP/s", "frac": tf / dmma_peak, "ms": ms_ / args.steps, "launches": k_}
    if "select" in stage:
        ms_, k_ = stage["select"]
        gb = curves_total / S * 8.0 * IMG * IMG / (ms_ * 1e-3) / 1e9
        others["select"] = {"kernel": "gpet::select_kernel", "bound": "hbm", "achieved": gb, "peak": peak, "unit": "GB/s",
                            "frac": gb / peak, "ms": ms_ / args.steps, "launches": k_}
    if "lml" in stage and stats.get("fit"):
        # objective kernel + L-BFGS-B advance kernel of the final fit (timed together: gpet_fit_rounds_f64 queues both);
        # ~m^3 flops per evaluation (Cholesky, inverse, K^-1 contraction) at m ~ 100 training points
        ms_, k_ = stage["lml"]
        ev = stats["fit"]["lml_evals"] * args.steps
        others["final_fit"] = {"kernel": "gpet::lml_blocked_kernel + gpet::lbfgsb_advance_kernel", "bound": "shared memory / latency",
                               "evaluations_per_s": ev / (ms_ * 1e-3), "achieved": ev * 1.0e6 / (ms_ * 1e-3) / 1e12,
                               "unit": "TFLOP/s (m^3 = 1e6 flops per evaluation)", "peak": dmma_peak,
                               "frac": ev * 1.0e6 / (ms_ * 1e-3) / 1e12 / dmma_peak, "ms": ms_ / args.steps, "launches": k_}
    for nm, kern_name in (("eig", "gpet::tridiag_reduce/ql/apply kernels"), ("posterior", "gpet::posterior_lowrank_kernel"),
                          ("density", "gpet::keep_gather_kernel + gpet::density_band_kernel"),
                          ("assemble", "gpet::factor_assemble_kernel")):
        if nm in stage:
            others[nm] = {"kernel": kern_name, "bound": "latency / shared memory", "ms": stage[nm][0] / args.steps,
                          "launches": stage[nm][1]}
    stage_ms = {k: round(v[0] / args.steps, 3) for k, v in stage.items()}

    if rank == 0:
        line = {
            "metric": "edges_traced_per_sec", "value": value, "unit": "traces/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "traces_per_gpu_per_step": B, "iterations_per_trace": stats["iters"],
                       "l2": "inputs larger than L2 (B x 2 MB images, B x 4 MB curve sets per iteration)",
                       "factor": "device low-rank Householder+QL eigensolver (rank 73 of 500)",
                       "normal_draws": "numpy RandomState on the host, one draw per (seed, iteration) shared by all traces",
                       "final_fit": "L-BFGS-B state machines on the " + ("device (gpet_lbfgsb_*)" if os.environ.get(
                           "GPET_FIT_DRIVER", "device").lower() == "device" else "host (scipy setulb workers)"),
                       "sub_batches": args.sub_batches, "prefetch": args.prefetch, "max_pending_fits": args.max_pending,
                       "fit_merge": args.fit_merge},
            "curves_scored_per_sec": world * curves_per_step * args.steps / (ms_total / 1e3),
            "e2e": {"value": e2e, "unit": "traces/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": roofline, "roofline_other_kernels": others, "stage_ms_per_step": stage_ms,
            "host_ms_last_step": stats.get("host_ms"),
            "final_fit": stats.get("fit"), "host_cores": host_cores(), "cpu_baseline": cpu, "clocks": clocks,
            "parity_checked": parity,
            "device_memory": {"max_allocated_after_warmup_gb": round(mem_warm / 1e9, 2),
                              "max_allocated_at_end_gb": round(mem_end / 1e9, 2)},
            "input_generation_s": round(t_gen, 2),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        raise SystemExit("bench.py: the benched path disagrees with the oracle: " + json.dumps(parity))


def launch_ranks(args):
    """`python bench.py --gpus N` outside torchrun: starts the N ranks itself (one process per GPU, NCCL rendezvous on
    127.0.0.1) and lets rank 0 print the line."""
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
    raise SystemExit(subprocess.call(cmd))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--traces", type=int, default=int(os.environ.get("GPET_BENCH_TRACES", "1250")),
                    help="traces per GPU per step")
    ap.add_argument("--sub-batches", type=int, default=1,
                    help="TraceBatch objects per step; 1 since the loop state lives on the device: larger launches win once "
                         "no host work has to be hidden (3914 vs 3597 traces/s at 1 and 2)")
    ap.add_argument("--prefetch", type=int, default=0,
                    help="batches BUILT ahead of the one inside the loop by a builder thread (0: only their host->device "
                         "copies are started ahead; building ahead measured slower - the device is already saturated)")
    ap.add_argument("--max-pending", type=int, default=1, help="fit jobs in flight before the oldest result is collected")
    ap.add_argument("--fit-merge", type=int, default=1, help="converged batches fitted together")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dedicated", dest="dedicated", action="store_false",
                    help="skip the dedicated full-shard launch of the scoring kernel (secondary roofline figure)")
    ap.add_argument("--parity", type=int, default=8,
                    help="images of the shard re-traced by the oracle after the timed region (0: skip)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args)
    elif args.gpus > 1 and "RANK" not in os.environ:
        launch_ranks(args)
    elif args.gpus != world:
        raise SystemExit(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun, or run "
                         f"`python bench.py --gpus {args.gpus}` outside torchrun and it starts the ranks itself)")
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
