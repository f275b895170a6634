"""world_size-2 gloo tests (CPU) of the host-side sharding logic in rlvi_b200.dist: shard bounds, group
creation from the torchrun environment, the statistics all-reduce, and the algebra the sharded path
relies on -- per-rank partial sums combined in rank order reproduce the oracle's fixed point and
statistics."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

from oracle import rlvi_np
from rlvi_b200 import dist as rdist


def test_shard_bounds_partition():
    for n in (0, 1, 7, 64, 1 << 20, (1 << 20) + 5):
        for world in (1, 2, 3, 8):
            b = [rdist.shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        rdist.shard_bounds(10, 2, 2)
    lo, hi = rdist.shard_bounds(1 << 26, 3, 8)
    assert lo % 2 == 0 and hi - lo == 1 << 23


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    g = rdist.ShardGroup.create("cpu")
    assert (g.rank, g.world) == (rank, world) and td.get_backend() == "gloo"
    rng = np.random.default_rng(0)
    n, d = 4001, 8
    X = rng.normal(size=(n, d))
    y = rng.normal(size=n)
    losses = 0.5 * rng.chisquare(1, size=n)
    lo, hi = rdist.shard_bounds(n, rank, world)
    e = np.exp(-losses[lo:hi])

    # the sharded fixed point: per-rank partial sums, summed in rank order on every rank
    pi_prev = np.full(hi - lo, 0.95)
    mean_prev = 0.95
    k = 0
    for k in range(1, 101):
        eps = 1 - mean_prev
        rho = eps / (1 - eps)
        pi_new = e / (rho + e)
        part = torch.tensor([pi_new.sum(), ((pi_new - pi_prev) ** 2).sum()], dtype=torch.float64)
        parts = [torch.empty_like(part) for _ in range(world)]
        td.all_gather(parts, part)
        tot = torch.stack(parts).sum(dim=0)           # rank order, identical on every rank
        mean_prev = float(tot[0]) / n
        pi_prev = pi_new
        if float(tot[1]) ** 0.5 < 1e-3:
            break
    with pytest.raises(RuntimeError):
        g.fp_dist(n)                                   # peer windows exist on CUDA only

    # statistics: local moments + all_reduce == global moments
    m = rlvi_np.weighted_moments(X[lo:hi], pi_prev, y[lo:hi])
    flat = torch.from_numpy(np.concatenate([[m["S0"], m["Swy"]], m["S1"], m["Sy"], m["G"].ravel()]))
    g.all_reduce(flat)
    q.put((rank, k, pi_prev, flat.numpy()))
    td.barrier()
    td.destroy_process_group()


def test_two_rank_sharding_reproduces_the_oracle():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    n, d = 4001, 8
    X = rng.normal(size=(n, d))
    y = rng.normal(size=n)
    losses = 0.5 * rng.chisquare(1, size=n)
    pi, eps, k, err = rlvi_np.fixed_point_trace(losses)
    assert got[0][1] == got[1][1] == k                 # same stop decision on both ranks
    pi_sharded = np.concatenate([got[0][2], got[1][2]])
    assert np.max(np.abs(pi_sharded - pi)) <= 1e-9 * np.max(pi)
    assert np.array_equal(got[0][3], got[1][3])         # all-reduce: same bits on both ranks
    m = rlvi_np.weighted_moments(X, pi, y)
    ref = np.concatenate([[m["S0"], m["Swy"]], m["S1"], m["Sy"], m["G"].ravel()])
    assert np.max(np.abs(got[0][3] - ref)) <= 1e-9 * np.max(np.abs(ref))
