"""CPU test of the N > 1 host logic: block partition + result gather over torch.distributed (gloo, world 2 and 3,
uneven shards)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, n, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from gaussian_process_edge_trace_b200 import dist as gd
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = gd.shard_bounds(n_total, world, rank)
    idx = np.arange(lo, hi)
    edges = (idx[:, None, None] * 1000 + np.arange(n)[None, :, None] * 2 + np.arange(2)[None, None, :]).astype(np.int64)
    cred = idx[:, None, None] + 0.25 * np.arange(2)[None, :, None] + 1e-3 * np.arange(n)[None, None, :]
    e, c = gd.gather_results(edges, cred, n_total)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), e=e, c=c)
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 7), (3, 8), (2, 4)])
def test_block_partition_and_gather(tmp_path, world, n_total):
    from gaussian_process_edge_trace_b200 import dist as gd
    bounds = [gd.shard_bounds(n_total, world, r) for r in range(world)]
    assert bounds[0][0] == 0 and bounds[-1][1] == n_total
    assert all(bounds[r][1] == bounds[r + 1][0] for r in range(world - 1))
    assert max(b - a for a, b in bounds) - min(b - a for a, b in bounds) <= 1
    n = 6
    mp.spawn(_worker, args=(world, _free_port(), n_total, n, str(tmp_path)), nprocs=world, join=True)
    idx = np.arange(n_total)
    want_e = (idx[:, None, None] * 1000 + np.arange(n)[None, :, None] * 2 + np.arange(2)[None, None, :]).astype(np.int64)
    want_c = idx[:, None, None] + 0.25 * np.arange(2)[None, :, None] + 1e-3 * np.arange(n)[None, None, :]
    for r in range(world):
        got = np.load(tmp_path / f"r{r}.npz")
        assert np.array_equal(got["e"], want_e) and np.array_equal(got["c"], want_c)


# ---- sample sharding: the two exchange steps of an iteration on CPU tensors (gloo) --------------------------------------
def _splat_fixed_point(Y_loc, idx_loc, wts, M, N, x_st):
    """numpy restatement of density_splat_kernel for one trace: u64 fixed-point grid (as int64) and dropped counts."""
    n = Y_loc.shape[0]
    Kp = idx_loc.shape[0]
    grid = np.zeros((M, N), dtype=np.uint64)
    dropped = np.zeros(Kp, dtype=np.int32)
    FX = 2.0 ** 60
    for c in range(Kp):
        s = idx_loc[c]
        if s < 0:
            continue
        for j in range(n):
            y = Y_loc[j, s]
            if y < 0.0 or y > M - 1:
                dropped[c] += 1
                continue
            iy = np.floor(y + 1.0)
            fy = (y + 1.0) - iy
            r0 = int(iy) - 1
            grid[r0, x_st + j] += np.uint64(np.rint(((1.0 - fy) * wts[c]) * FX))
            if fy > 0.0 and r0 + 1 < M:
                grid[r0 + 1, x_st + j] += np.uint64(np.rint((fy * wts[c]) * FX))
    return grid, dropped


def _sample_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from gaussian_process_edge_trace_b200 import dist as gd
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(11)                      # same data on every rank
    b, n, S, Kp, M, N, x_st = 2, 12, 24, 6, 10, 16, 2
    Y = rng.uniform(-1.5, M + 0.5, size=(b, n, S))       # some points outside the image
    cost = rng.uniform(1.0, 2.0, size=(b, S))
    s0, s1 = gd.sample_block(S, world, rank)
    # (1) cost all-gather
    full = torch.empty((b, S), dtype=torch.float64)
    gd.gather_costs(torch.from_numpy(cost[:, s0:s1].copy()), full)
    assert np.array_equal(full.numpy(), cost)
    # global top-Kp (replicated), local share of the kept curves
    idx = np.argsort(cost, axis=1, kind="stable")[:, :Kp].astype(np.int32)
    inv = 1 / np.take_along_axis(cost, idx.astype(np.int64), axis=1)
    wts = inv / inv.sum(axis=1, keepdims=True)
    loc = gd.local_keep_index(torch.from_numpy(idx), s0, s1 - s0).numpy()
    assert np.array_equal(loc >= 0, (idx >= s0) & (idx < s1))
    # (2) local splat into the kernel's workspace layout, exact integer all-reduce
    work = torch.zeros((b * M * N + b) * 8 + b * Kp * 4, dtype=torch.uint8)
    grid = work[: b * M * N * 8].view(torch.int64).view(b, M, N)
    cnt = work[(b * M * N + b) * 8:].view(torch.int32).view(b, Kp)
    for t in range(b):
        g, d = _splat_fixed_point(Y[t][:, s0:s1], loc[t], wts[t], M, N, x_st)
        grid[t] = torch.from_numpy(g.view(np.int64))
        cnt[t] = torch.from_numpy(d)
    gd.reduce_density(work, b, M, N, Kp)
    np.savez(os.path.join(out_dir, f"s{rank}.npz"), grid=grid.numpy().copy(), cnt=cnt.numpy().copy(), idx=idx, wts=wts, Y=Y)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sample_sharding_exchanges_are_exact(tmp_path, world):
    """The sharded density (sum over ranks of the fixed-point splats of the curves each rank owns) equals the
    single-process splat of all kept curves BIT FOR BIT, for any rank count; costs gather into global sample order."""
    mp.spawn(_sample_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = [np.load(tmp_path / f"s{r}.npz") for r in range(world)]
    Y, idx, wts = got[0]["Y"], got[0]["idx"], got[0]["wts"]
    for t in range(Y.shape[0]):
        g, d = _splat_fixed_point(Y[t], idx[t], wts[t], 10, 16, 2)
        for r in range(world):
            assert np.array_equal(got[r]["grid"][t], g.view(np.int64))
            assert np.array_equal(got[r]["cnt"][t], d)
    from gaussian_process_edge_trace_b200 import dist as gd
    with pytest.raises(ValueError):
        gd.sample_block(10, 3, 0)
