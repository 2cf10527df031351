"""GPU check (torchrun, >= 2 ranks): a sample-sharded trace equals the single-GPU trace of the same arguments
bit for bit (edge_pred, observation sets, credible interval).  Run:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/check_sample_sharding.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
import numpy as np, torch, torch.distributed as dist
import gpet_oracle as O
import __graft_entry__
__graft_entry__.build()
from gaussian_process_edge_trace_b200 import TraceBatch, dist as gd

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
kern = O.kernel_builder((11, 5))
imgs, inits = [], []
for s in range(2):
    img, edge = O.construct_test_img((120, 160), 40 + 6 * s, 2, 0.01, "sinusoidal", 0.4, noise_seed=s + 1)
    imgs.append(O.comp_grad_img(img, kern)); inits.append(edge[[0, -1], :][:, [1, 0]])
imgs, inits = np.stack(imgs), np.stack(inits)
for S, keep in ((512, 0.2), (12000, 0.1)):
    kw = dict(kernel_options={"kernel": "RBF", "sigma_f": 25, "length_scale": 15}, noise_y=1, N_samples=S,
              score_thresh=1, delta_x=8, keep_ratio=keep, pixel_thresh=3, seed=9, fix_endpoints=True)
    one = TraceBatch(inits, imgs, **kw)
    e1, c1 = one.trace()
    sh = TraceBatch(inits, imgs, sample_group=True, **kw)
    e2, c2 = sh.trace()
    ok = np.array_equal(e1, e2) and all(np.array_equal(a, b) for a, b in zip(one.fobs, sh.fobs)) and \
        all(np.allclose(a[0], b[0], rtol=1e-9, atol=0) for a, b in zip(c1, c2)) and np.array_equal(one.n_iter, sh.n_iter)
    flag = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"S={S}: sample-sharded over {dist.get_world_size()} ranks == single GPU: {bool(flag.item())} "
              f"(iterations {one.n_iter.tolist()})", flush=True)
    assert ok
dist.destroy_process_group()
