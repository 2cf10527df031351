"""Developer tool (GPU box): one cfg 5 shard traced with the loop and the final fit NOT overlapped - wall/CUDA-event
time of the constructor, the while-loop (per stage) and the fit, each alone on the GPU."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
import numpy as np, torch
import bench
import __graft_entry__
__graft_entry__.build()
from gaussian_process_edge_trace_b200 import TraceBatch, gpet_utils
from gaussian_process_edge_trace_b200.engine import StageTimers

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1250
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
imgs = np.empty((B, 500, 500)); inits = np.empty((B, 2, 2), dtype=np.int64)
for i in range(B):
    imgs[i], inits[i] = bench.make_image(i)
d_imgs = torch.from_numpy(imgs).cuda()
kern = gpet_utils.kernel_builder((11, 5))
out = {}
for rep in range(reps):
    timers = StageTimers()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    grad = gpet_utils.comp_grad_img(d_imgs, kern, return_tensor=True)
    tb = TraceBatch(inits, grad, timers=timers, **bench.TRACE_KW)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    act = []
    while True:
        more = tb.step_launch()
        if not more:
            break
        act.append(tb._n_active)
        tb.step_finish()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    loop_stage = {k: (round(v[0], 2), v[1]) for k, v in timers.collect().items()}
    timers.reset()
    _, _, finfo = tb.final_fit_all()
    torch.cuda.synchronize(); t3 = time.perf_counter()
    fit_stage = {k: (round(v[0], 2), v[1]) for k, v in timers.collect().items()}
    out = dict(B=B, construct_ms=1e3 * (t1 - t0), loop_ms=1e3 * (t2 - t1), fit_ms=1e3 * (t3 - t2), active_per_iteration=act,
               loop_stage_ms=loop_stage, fit_stage_ms=fit_stage, fit_info={k: int(finfo[k]) for k in ("rounds", "lml_evals")},
               host_ms=tb.host_ms)
    print(json.dumps(out))
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "split.json"), "w"), indent=1)
