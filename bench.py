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
def _cpu_trace(i):
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


def cpu_arm(n_workers, rounds, first_image=0):
    """Each round traces one image per worker process. Returns (traces/s over all rounds, per-round seconds, curves)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    per_round, curves = [], 0
    with ctx.Pool(n_workers) as pool:
        for r in range(rounds):
            t0 = time.time()
            out = pool.map(_cpu_trace, range(first_image + r * n_workers, first_image + (r + 1) * n_workers))
            per_round.append(time.time() - t0)
            curves += sum(o[1] for o in out)
    return n_workers * rounds / sum(per_round), per_round, curves


def host_cores():
    """Cores this process may run on (affinity mask when the platform has one, else the machine's count)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    w, k = max(args.warmup, 0), max(args.steps, 1)
    if w:
        cpu_arm(cores, w, first_image=10_000)
    tps, per_round, curves = cpu_arm(cores, k)
    ms = 1e3 * sum(per_round) / k
    line = {
        "impl": "reference", "metric": "edges_traced_per_sec", "value": tps, "unit": "traces/s", "n_gpus": args.gpus,
        "steps": k, "warmup": w, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": {"workload": WORKLOAD, "traces_per_step": cores},
        "curves_scored_per_sec": curves / sum(per_round),
        "cpu_baseline": {"value": tps, "unit": "traces/s", "cores": cores, "kind": "port",
                         "sample": f"{cores} traces per step (one per host core), oracle port of the reference "
                                   "numpy/scipy path incl. its per-curve Python loop"},
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
        tps, per_round, curves = cpu_arm(cores, 1)
        cpu = {"value": tps, "unit": "traces/s", "cores": cores, "kind": "port",
               "curves_scored_per_sec": curves / sum(per_round),
               "sample": f"{cores} traces of the same workload (one per host core, {per_round[0]:.1f} s), oracle "
                         "port of the reference numpy/scipy path incl. its per-curve Python loop"}
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    __graft_entry__.build()
    from gaussian_process_edge_trace_b200 import TraceBatch, gpet_utils
    from gaussian_process_edge_trace_b200.engine import StageTimers, trace_pipelined
    dev = torch.device(f"cuda:{local}")
    B = args.traces
    kern = gpet_utils.kernel_builder((11, 5))

    # synthetic inputs of this rank's shard (distinct images across ranks), pinned host + resident device copies
    t_gen = time.time()
    imgs = np.empty((B, IMG, IMG), dtype=np.float64)
    inits = np.empty((B, 2, 2), dtype=np.int64)
    for i in range(B):
        imgs[i], inits[i] = make_image(rank * B + i)
    h_imgs = torch.from_numpy(imgs).pin_memory()
    d_imgs = h_imgs.to(dev)
    t_gen = time.time() - t_gen
    timers = StageTimers()
    stats = {}

    copy_stream = torch.cuda.Stream()

    cuts = np.linspace(0, B, max(1, min(args.sub_batches, B)) + 1).astype(int)
    spans = list(zip(cuts[:-1], cuts[1:]))

    def upload(resident):
        """Inputs of one step, per sub-batch: resident views, or (e2e) host -> device copies issued on a copy stream -
        non-blocking, so that the copies overlap whatever the GPU is tracing at that moment."""
        if resident:
            return [(d_imgs[a:b], None) for a, b in spans]
        parts = []
        with torch.cuda.stream(copy_stream):
            for a, b in spans:
                d = h_imgs[a:b].to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                parts.append((d, ev))
        return parts

    def step(resident, parts=None):
        # the shard is traced as `--sub-batches` TraceBatch objects whose host and device phases overlap
        main = torch.cuda.current_stream()
        if parts is None:
            parts = upload(resident)

        def factory(k):
            def make():
                d, ev = parts[k]
                if ev is not None:
                    main.wait_event(ev)
                    d.record_stream(main)
                grad = gpet_utils.comp_grad_img(d, kern, return_tensor=True)
                a, b = spans[k]
                return TraceBatch(inits[a:b], grad, timers=timers, **TRACE_KW)
            return make

        tbs = [factory(k) for k in range(len(spans))]
        handle = trace_pipelined(tbs, window=args.window, fit_merge=args.fit_merge, wait=False,
                                 own_streams=args.own_streams)

        def collect():
            edges, creds = handle.result()
            stats["curves"] = sum(tb.curves_scored for tb in tbs)
            # stencil(3), normalise(3), grad KDE(5), transpose(1) per sub-batch
            stats["launches"] = sum(tb.kernel_launches + 3 + 3 + 5 + 1 for tb in tbs)
            stats["iters"] = int(max(tb.n_iter.max() for tb in tbs))
            hm = {}
            for tb in tbs:
                for k, v in tb.host_ms.items():
                    hm[k] = hm.get(k, 0.0) + v
            stats["host_ms"] = {k: round(v, 1) for k, v in hm.items()}
            stats["fit"] = {k: int(sum(tb.final_info[k] for tb in tbs)) for k in ("rounds", "lml_evals")}
            stats["edges"] = edges
            return edges, creds

        return collect

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(resident, k):
        # steps are streamed: the tracing loops of step i+1 start while the last final fits of step i are still running
        # in the background (host bound); every result is collected before returning (--no-stream: one by one)
        pending = []
        nxt = upload(resident)                       # e2e: the copies of step i+1 are issued before step i is traced
        for i in range(k):
            parts, nxt = nxt, (upload(resident) if (i + 1 < k and not args.no_stream) else None)
            c = step(resident, parts)
            if args.no_stream:
                c()
                nxt = upload(resident) if i + 1 < k else None
            else:
                pending.append(c)
        for c in pending:
            c()

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

    # warm-up in the same streamed pattern as the timed region (same number of live sub-batches => the caching
    # allocator, the worker pool and the fit stream are in their steady state when the clock starts)
    if args.warmup > 0:
        run_steps(True, args.warmup)
        run_steps(False, 1)
    timers.reset()
    sampler = ClockSampler(local)
    sampler.start()
    ms_total = timed(True, args.steps)
    stage = timers.collect()
    launches_per_step = stats["launches"]
    curves_per_step = stats["curves"]
    ms_e2e = timed(False, args.steps)
    clocks = sampler.stop()

    value = world * B * args.steps / (ms_total / 1e3)
    e2e = world * B * args.steps / (ms_e2e / 1e3)
    h2d = B * IMG * IMG * 8
    d2h = B * IMG * (2 * 8 + 2 * 8)        # edge_pred int64[n,2] + credint 2 x float64[n]

    # roofline of the scoring kernel: algorithmic bytes = 8 n + 8 per curve, CUDA-event time on the
    # launching stream.  Two live measurements: (1) `in_region`: all launches of the timed region (sub-batch sized,
    # shrinking as traces converge, and sharing the GPU with the final-fit kernels of the high-priority stream);
    # (2) headline: a dedicated pass right after the timed region - the first iterations of the whole shard as ONE
    # TraceBatch, nothing else in flight - i.e. the kernel at the workload's full launch size.
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, peak_src = 6650.0, "fallback"
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured"
    sc_ms, sc_n = stage.get("score", (0.0, 0))
    n, S = IMG, TRACE_KW["N_samples"]
    in_region = curves_per_step * args.steps * (8 * n + 8) / (sc_ms * 1e-3) / 1e9 if sc_n else None
    rt = StageTimers()
    grad_all = gpet_utils.comp_grad_img(d_imgs, kern, return_tensor=True)
    tb_r = TraceBatch(inits, grad_all, timers=rt, **TRACE_KW)
    for _ in range(2):
        tb_r.step()
    rt.reset()
    n_it = 4
    for _ in range(n_it):
        tb_r.step()
    r_ms, r_n = rt.collect().get("score", (0.0, 0))
    achieved = (n_it * B * S * (8 * n + 8)) / (r_ms * 1e-3) / 1e9 if r_n else None
    del tb_r, grad_all
    roofline = {"kernel": "score_streamN_kernel<SCAN=1,STAGES=4,MINB=4,CPT=2> (in_region: mostly score_stream_kernel, the one-curve-per-thread form used for sub-batch sized launches)", "bound": "hbm", "achieved": achieved, "peak": peak, "peak_source": peak_src,
                "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH,
                "ms_per_launch": r_ms / r_n if r_n else None, "launches": r_n,
                "bytes_per_launch": B * S * (8 * n + 8),
                "measured": "dedicated pass inside bench.py after the timed region: full-shard launches, no other stream",
                "in_region": {"achieved": in_region, "frac": (in_region / peak) if in_region else None,
                              "ms_per_launch": sc_ms / sc_n if sc_n else None, "launches": sc_n}}
    stage_ms = {k: round(v[0] / args.steps, 3) for k, v in stage.items()}

    if rank == 0:
        line = {
            "metric": "edges_traced_per_sec", "value": value, "unit": "traces/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "traces_per_gpu_per_step": B, "iterations_per_trace": stats["iters"],
                       "l2": "inputs larger than L2 (B x 2 MB images, B x 4 MB curve sets per iteration)",
                       "factor": "device low-rank Householder+QL eigensolver (rank 73 of 500)",
                       "final_fit": "L-BFGS-B state machines on the " + ("device (gpet_lbfgsb_*)" if os.environ.get(
                           "GPET_FIT_DRIVER", "device").lower() == "device" else "host (scipy setulb workers)"),
                       "sub_batches": args.sub_batches, "window": args.window, "own_streams": args.own_streams,
                       "steps_streamed": not args.no_stream},
            "curves_scored_per_sec": world * curves_per_step * args.steps / (ms_total / 1e3),
            "e2e": {"value": e2e, "unit": "traces/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": roofline, "stage_ms_per_step": stage_ms, "host_ms_last_step": stats.get("host_ms"),
            "final_fit": stats.get("fit"), "host_cores": host_cores(), "cpu_baseline": cpu, "clocks": clocks,
            "input_generation_s": round(t_gen, 2),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--traces", type=int, default=int(os.environ.get("GPET_BENCH_TRACES", "1250")),
                    help="traces per GPU per step")
    ap.add_argument("--sub-batches", type=int, default=2, help="TraceBatch objects per step (pipelined)")
    ap.add_argument("--window", type=int, default=2, help="sub-batches inside the tracing loop at a time")
    ap.add_argument("--fit-merge", type=int, default=2, help="converged sub-batches fitted together")
    ap.add_argument("--own-streams", dest="own_streams", action="store_true", default=False,
                    help="every sub-batch launches on a CUDA stream of its own (+7 %% resident, but an erratic e2e figure)")
    ap.add_argument("--no-own-streams", dest="own_streams", action="store_false")
    ap.add_argument("--no-stream", action="store_true", help="finish every step (incl. its last final fit) before the next")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
