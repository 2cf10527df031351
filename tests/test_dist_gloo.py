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
