"""Developer tool: a few iterations of the batched loop (+ optionally the final fit) for ncu captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
import numpy as np, torch
import bench
import __graft_entry__
__graft_entry__.build()
from gaussian_process_edge_trace_b200 import TraceBatch, gpet_utils

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
fit = len(sys.argv) > 3 and sys.argv[3] == "fit"
imgs = np.empty((B, 500, 500)); inits = np.empty((B, 2, 2), dtype=np.int64)
for i in range(B):
    imgs[i], inits[i] = bench.make_image(i)
grad = gpet_utils.comp_grad_img(torch.from_numpy(imgs).cuda(), gpet_utils.kernel_builder((11, 5)), return_tensor=True)
tb = TraceBatch(inits, grad, **bench.TRACE_KW)
if fit:
    tb.run_loop()
    tb.final_fit_all()
else:
    for _ in range(iters):
        tb.step()
torch.cuda.synchronize()
print("ok", tb.n_iter.max())
