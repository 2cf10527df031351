"""Multi-GPU execution: whole-trace sharding (default) and sample sharding.

Traces are independent (the only shared inputs are the read-only standard-normal draws, which every rank
regenerates from the seed), so a batch is block-partitioned over the ranks of a torchrun job and each rank runs
its own `TraceBatch`; there is NO collective on the data path. The only communication is the final gather of the
results (edge_pred int64[n, 2] + credible interval 2 x float64[n] per trace = 16 KB/trace at n = 500) over
torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np
import torch
import torch.distributed as dist

# ---- sample sharding: one (or a few) traces, the N_samples posterior curves split over the ranks ----------------------
# Rank g draws and scores the curves s in [g*S/G, (g+1)*S/G) of EVERY trace. Per iteration there are exactly two
# exchange steps, both exact: (1) all-gather of the costs (8*S bytes per trace), after which every rank derives the
# same global top-N_keep and weights; (2) all-reduce (sum) of the fixed-point (2^-60) density grid and of the per-curve
# dropped-point counts - integer sums, so the result does not depend on the rank count or the reduction order.
# Posterior, factor, selection and the final fit are replicated (deterministic). The helpers below run on CUDA/NCCL
# and CPU/gloo tensors alike (the gloo tests exercise them with world_size 2).


def sample_block(S, world, rank):
    """[s0, s1) of the samples owned by `rank`; S must be divisible by `world`."""
    if S % world:
        raise ValueError(f"N_samples={S} must be divisible by the {world} ranks of the sample group")
    return rank * (S // world), (rank + 1) * (S // world)


def gather_costs(cost_loc, cost_out, group=None):
    """cost_loc [b, S/G] of this rank -> cost_out [b, S] on every rank (sample s = rank * S/G + local index)."""
    world = dist.get_world_size(group)
    parts = [torch.empty_like(cost_loc) for _ in range(world)]
    dist.all_gather(parts, cost_loc.contiguous(), group=group)
    b, sl = cost_loc.shape
    cost_out.view(b, world, sl).copy_(torch.stack(parts, dim=1))
    return cost_out


def local_keep_index(idx, s0, s_loc):
    """Global kept-curve indices -> local sample indices of this rank, -1 for curves owned elsewhere."""
    loc = idx - s0
    return torch.where((loc >= 0) & (loc < s_loc), loc, torch.full_like(loc, -1))


def reduce_density(work, b, M, N, Kp, group=None):
    """Sums the splat workspace of gpet_density_splat_f64 (u64 grid[b][M][N] | f64 scale[b] | i32 dropped[b][Kp]) over
    the ranks in place: the grid as int64 (two's complement addition == unsigned addition) and the counts."""
    ngrid = b * M * N
    grid = work[: ngrid * 8].view(torch.int64)
    o = (ngrid + b) * 8
    cnt = work[o: o + b * Kp * 4].view(torch.int32)
    dist.all_reduce(grid, group=group)
    dist.all_reduce(cnt, group=group)


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


def trace_sample_sharded(init, grad_img, group=None, **kw):
    """Traces `init` / `grad_img` (one or a few traces) with the posterior samples split over the ranks of `group`
    (default: all ranks). Every rank returns the full (edges, creds) - identical on all ranks and identical to a
    single-GPU run of the same arguments."""
    from .engine import TraceBatch
    tb = TraceBatch(init, grad_img, sample_group=(True if group is None else group), **kw)
    return tb.trace()

