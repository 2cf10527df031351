"""Developer tool (GPU box): block Jacobi eigensolver on covariance-like matrices - sweeps to converge, time per sweep,
time of the batched pivot eigensolver alone, and torch.linalg.eigh (cuSOLVER) on the same input for scale.
usage: python tools/bench_jacobi.py [n] [B]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__
__graft_entry__.build()
from gaussian_process_edge_trace_b200._cabi import call, ptr, query

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = "cuda"
from gaussian_process_edge_trace_b200._cabi import load
if os.environ.get("JAC_BLOCK"):
    load().gpet_set_tuning(10, int(os.environ["JAC_BLOCK"]))    # 32: 64 x 64 pivots, 64: 128 x 128 pivots
if os.environ.get("JAC_INNER"):
    load().gpet_set_tuning(12, int(os.environ["JAC_INNER"]))    # cap on the inner sweeps of a pivot solve (0: to convergence)
if os.environ.get("JAC_SYM"):
    load().gpet_set_tuning(13, int(os.environ["JAC_SYM"]))      # 1: fused two-sided update of the lower half, 0: column + row pass
if os.environ.get("JAC_EIG"):
    load().gpet_set_tuning(2, int(os.environ["JAC_EIG"]))       # parallel cyclic Jacobi for the pivots instead of Householder + QL
st = torch.cuda.current_stream().cuda_stream
x = torch.arange(n, device=dev, dtype=torch.float64)
d = (x[:, None] - x[None, :]).abs() / 20.0
Kss = 5625.0 * (1 + 5 ** 0.5 * d + 5 * d * d / 3) * torch.exp(-(5 ** 0.5) * d)
covs = []
g = torch.Generator(device="cpu").manual_seed(0)
for b in range(B):
    m = 50 + 40 * b
    idx = torch.sort(torch.randperm(n, generator=g)[:m]).values.to(dev)
    Kmm = Kss[idx][:, idx] + torch.eye(m, device=dev, dtype=torch.float64)
    covs.append(Kss - Kss[:, idx] @ torch.linalg.solve(Kmm, Kss[idx, :]))
cov = torch.stack(covs)
cov = 0.5 * (cov + cov.transpose(1, 2))
np_ = (n + 127) // 128 * 128
A = torch.empty((B, np_, np_), dtype=torch.float64, device=dev)
V = torch.empty_like(A)
off = torch.empty((B, 2), dtype=torch.float64, device=dev)
work = torch.empty(int(query("gpet_block_jacobi_workspace_bytes", B, np_)), dtype=torch.uint8, device=dev)


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


for rep in range(2):
    call("gpet_block_jacobi_init_f64", ptr(cov), B, n, np_, ptr(A), ptr(V), st)
    rels, ms = [], []
    for sweep in range(30):
        e0 = ev()
        call("gpet_block_jacobi_sweep_f64", ptr(A), ptr(V), B, np_, ptr(off), ptr(work), st)
        e1 = ev()
        o = off.cpu().numpy()
        ms.append(e0.elapsed_time(e1))
        rels.append(float(np.sqrt(o[:, 0] / o[:, 1]).max()))
        if rels[-1] <= 3e-13:
            break
print("n", n, "B", B, "sweeps", len(rels), "rel", ["%.1e" % r for r in rels])
print("ms per sweep", ["%.1f" % m for m in ms], "total", round(sum(ms), 1))
# the pivot eigensolver alone on the pivots of one step
JP = 2 * int(os.environ.get("JAC_BLOCK", "32"))
nm = B * (np_ // JP)
P = torch.randn((nm, JP, JP), dtype=torch.float64, device=dev)
P = P + P.transpose(1, 2)
dd = torch.empty((nm, JP), dtype=torch.float64, device=dev)
Q = torch.empty_like(P)
sw = torch.empty((nm,), dtype=torch.int32, device=dev)
ework = torch.empty(int(query("gpet_sym_eig_workspace_bytes", nm, JP)), dtype=torch.uint8, device=dev)
for rep in range(2):
    P2 = P.clone()
    e0 = ev()
    call("gpet_sym_eig_f64", ptr(P2), nm, JP, ptr(dd), ptr(Q), ptr(sw), ptr(ework), st)
    e1 = ev()
    torch.cuda.synchronize()
print("pivot eig: %d random matrices of %d x %d: %.2f ms" % (nm, JP, JP, e0.elapsed_time(e1)))
for rep in range(2):
    e0 = ev()
    w_, v_ = torch.linalg.eigh(cov)
    e1 = ev()
    torch.cuda.synchronize()
print("torch.linalg.eigh (cuSOLVER, for scale): %.1f ms" % e0.elapsed_time(e1))
