"""Developer tool (GPU box): times the L-BFGS-B advance kernel on E runs with a synthetic objective evaluated on the
device by torch ops between the rounds (so the timing of the advance launches is not mixed with the objective kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch as T
import __graft_entry__
__graft_entry__.build()
from gaussian_process_edge_trace_b200 import _gp_host as H
from gaussian_process_edge_trace_b200._cabi import call, ptr, load

lib = load()
E = int(sys.argv[1]) if len(sys.argv) > 1 else 16250
lo, hi = H.FINAL_BOUNDS[:, 0].copy(), H.FINAL_BOUNDS[:, 1].copy()
rng = np.random.RandomState(5)
dev = T.device("cuda")
x0 = T.from_numpy(rng.uniform(lo, hi, size=(E, 3))).to(dev)
cen = T.from_numpy(rng.uniform(lo - 3, hi + 3, size=(E, 3))).to(dev)
sc = T.from_numpy(np.exp(rng.uniform(-2, 1, size=(E, 3)))).to(dev)
d_lo, d_hi = T.from_numpy(lo).to(dev), T.from_numpy(hi).to(dev)
nd, ni = lib.gpet_lbfgsb_state_doubles(), lib.gpet_lbfgsb_state_ints()
st = T.cuda.current_stream().cuda_stream
for nt in (64, 0):
    lib.gpet_set_tuning(8, nt)
    d_state = T.empty((nd, E), dtype=T.float64, device=dev)
    i_state = T.empty((ni, E), dtype=T.int32, device=dev)
    d_tr = T.arange(E, dtype=T.int32, device=dev)
    d_theta = T.zeros((E, 3), dtype=T.float64, device=dev)
    d_f = T.zeros(E, dtype=T.float64, device=dev)
    d_g = T.zeros((E, 3), dtype=T.float64, device=dev)
    d_ev = T.full((E,), -1, dtype=T.int32, device=dev)
    d_n = T.zeros(3, dtype=T.int32, device=dev)
    call("gpet_lbfgsb_init_f64", ptr(d_state), ptr(i_state), E, ptr(x0), ptr(d_lo), ptr(d_hi), st)
    first, total, per = 1, 0.0, []
    for r in range(400):
        e0, e1 = T.cuda.Event(enable_timing=True), T.cuda.Event(enable_timing=True)
        e0.record()
        call("gpet_lbfgsb_advance_f64", ptr(d_state), ptr(i_state), E, first, ptr(d_tr), ptr(d_f), ptr(d_g), ptr(d_theta),
             ptr(d_ev), ptr(d_n), st)
        e1.record()
        n = int(d_n[0].item())
        ms = e0.elapsed_time(e1)
        per.append((n, ms))
        total += ms
        if n == 0:
            break
        d = (d_theta - cen) * sc
        d_f.copy_(0.5 * (d * d).sum(dim=1) + T.cos(1.3 * d_theta).sum(dim=1))
        d_g.copy_(sc * d - 1.3 * T.sin(1.3 * d_theta))
        first = 0
    print(f"threads={nt}: {len(per)} rounds, advance total {total:.2f} ms; first rounds (active, ms):",
          [(n, round(ms, 3)) for n, ms in per[:14]], "... evals", int(d_n[1].item()))
