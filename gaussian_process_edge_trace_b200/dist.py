"""Multi-GPU execution: whole-trace sharding.

Traces are independent (the only shared inputs are the read-only standard-normal draws, which every rank
regenerates from the seed), so a batch is block-partitioned over the ranks of a torchrun job and each rank runs
its own `TraceBatch`; there is NO collective on the data path. The only communication is the final gather of the
results (edge_pred int64[n, 2] + credible interval 2 x float64[n] per trace = 16 KB/trace at n = 500) over
torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_items, world, rank):
    """Contiguous block partition: the first n_items % world ranks get one extra item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_results(edges_local, cred_local, n_total, device=None, group=None):
    """All-gathers per-trace results of a block-partitioned batch in global trace order.
    edges_local int64[b, n, 2], cred_local float64[b, 2, n] (b = this rank's shard). Returns the full
    (edges int64[n_total, n, 2], cred float64[n_total, 2, n]) on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return np.asarray(edges_local), np.asarray(cred_local)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    n = edges_local.shape[1]
    bmax = max(shard_bounds(n_total, world, r)[1] - shard_bounds(n_total, world, r)[0] for r in range(world))
    e = torch.zeros((bmax, n, 2), dtype=torch.int64, device=device)
    c = torch.zeros((bmax, 2, n), dtype=torch.float64, device=device)
    b = edges_local.shape[0]
    lo, hi = shard_bounds(n_total, world, rank)
    if b != hi - lo:
        raise ValueError(f"rank {rank} holds {b} traces, expected {hi - lo}")
    e[:b] = torch.as_tensor(np.ascontiguousarray(edges_local), dtype=torch.int64).to(device)
    c[:b] = torch.as_tensor(np.ascontiguousarray(cred_local), dtype=torch.float64).to(device)
    es = [torch.empty_like(e) for _ in range(world)]
    cs = [torch.empty_like(c) for _ in range(world)]
    dist.all_gather(es, e, group=group)
    dist.all_gather(cs, c, group=group)
    out_e = np.empty((n_total, n, 2), dtype=np.int64)
    out_c = np.empty((n_total, 2, n), dtype=np.float64)
    for r in range(world):
        a, z = shard_bounds(n_total, world, r)
        out_e[a:z] = es[r][: z - a].cpu().numpy()
        out_c[a:z] = cs[r][: z - a].cpu().numpy()
    return out_e, out_c


def trace_sharded(init, grad_img, gather=True, **kw):
    """Traces a batch over all ranks of the current process group: rank r traces the r-th block of `init` /
    `grad_img` on its own GPU (one process per GPU). Returns (edges, cred) for the whole batch when `gather`,
    otherwise for the local shard only."""
    from .engine import TraceBatch
    init = np.asarray(init)
    n_total = init.shape[0]
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_bounds(n_total, world, rank)
    obs = kw.pop("obs", None)
    if obs is not None and not isinstance(obs, np.ndarray):
        obs = list(obs)[lo:hi]
    tb = TraceBatch(init[lo:hi], grad_img[lo:hi], obs=obs, **kw)
    edges, creds = tb.trace()
    cred = np.stack([np.stack(c) for c in creds]) if len(creds) else np.zeros((0, 2, edges.shape[1]))
    if not gather:
        return edges, cred
    return gather_results(edges, cred, n_total)
