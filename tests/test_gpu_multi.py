"""Multi-GPU parity (needs >= 2 GPUs: `gpurun --gpus 2`): the sample-sharded E+M step -- fixed point with the
in-kernel NVLink exchange, statistics with the NCCL all-reduce -- against the oracle on the unsharded data
and against the single-GPU kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, d, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    from rlvi_b200 import dist as rdist, ops, synth
    g = rdist.ShardGroup.create(dev)
    X, y, theta = synth.logistic_data(n, d, seed=7)
    params = np.concatenate([[0.1], theta])
    lo, hi = rdist.shard_bounds(n, rank, world)
    Xd = torch.from_numpy(X[lo:hi]).to(dev)
    yd = torch.from_numpy(y[lo:hi]).to(dev)
    pd = torch.from_numpy(params).to(dev)
    out = []
    for rep in range(3):                                   # several calls: exercises call_index / window reuse
        _, e, _ = ops.loss(ops.LOSS_LOGISTIC_CE, Xd, pd, y=yd, intercept=True, want_losses=False, want_e=True)
        pi, res = ops.fixed_point(None, e_work=e, dist=g.fp_dist(n))
        mom = ops.weighted_moments(Xd, pi)
        g.all_reduce(mom)
        r = ops.read_result(res)
        out.append((r, pi.cpu().numpy(), mom.cpu().numpy()))
    # the host-buffer entry point on this rank's shard (bench.py's e2e at N > 1)
    host = ops.em_step_logistic_host(X[lo:hi], y[lo:hi], params, device=rank, group=g, n_global=n)
    # deep variant (FP32) sharded: min / max reductions cross the ranks too
    rng = np.random.default_rng(3)
    resid = rng.exponential(1.0, size=n).astype(np.float32) + 0.25
    w0 = np.ones(n, dtype=np.float32)
    rt = torch.from_numpy(resid[lo:hi]).to(dev)
    wt = torch.from_numpy(w0[lo:hi]).to(dev)
    dres = ops.fixed_point_deep(rt, wt, dist=g.fp_dist(n))
    q.put((rank, out, ops.read_result(dres), rt.cpu().numpy(), wt.cpu().numpy(), host))
    g.close()
    torch.distributed.destroy_process_group()


# the last case shards 2^23 + 4099 samples over TWO GPUs: above 2^22 samples per GPU the fixed point streams its shard through
# the bulk-copy ring (ragged segments, partial last trips) while exchanging its sums over NVLink
@pytest.mark.parametrize("n,max_world", [(4096 + 37, 4), (1 << 20, 4), ((1 << 23) + 4099, 2)])
def test_sharded_em_step_matches_oracle(n, max_world):
    world = min(torch.cuda.device_count(), max_world)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    from oracle import deep_ref, rlvi_np
    from rlvi_b200 import dist as rdist, synth
    d = 64
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, d, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=300) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    X, y, theta = synth.logistic_data(n, d, seed=7)
    ref = rlvi_np.em_step_logistic(X, y, np.concatenate([[0.1], theta]))
    tol = max(1e-9, 4 * 2.0 ** -53 / float(ref["pi"].mean()))
    for rep in range(3):
        results = [g[1][rep][0] for g in got]
        assert all(r["iters"] == ref["iters"] for r in results)
        # every rank computed bitwise-identical scalars (same stop decision everywhere)
        assert all(r == results[0] for r in results)
        pi = np.concatenate([g[1][rep][1] for g in got])
        assert np.max(np.abs(pi - ref["pi"])) <= tol * np.max(ref["pi"])
        moms = [g[1][rep][2] for g in got]
        assert all(np.array_equal(m, moms[0]) for m in moms)
        G = moms[0][2 + 2 * d:].reshape(d, d)
        assert np.max(np.abs(G / moms[0][0] - ref["G"] / ref["S0"])) <= 1e-9 * np.max(np.abs(ref["G"] / ref["S0"]))
    # sharded host-buffer step: global statistics on every rank, this shard's posteriors
    hosts = [g[5] for g in got]
    assert all(h["result"]["iters"] == ref["iters"] for h in hosts)
    assert all(np.array_equal(h["moments"], hosts[0]["moments"]) for h in hosts)
    Gh = hosts[0]["moments"][2 + 2 * d:].reshape(d, d)
    assert np.max(np.abs(Gh / hosts[0]["moments"][0] - ref["G"] / ref["S0"])) <= 1e-9 * np.max(np.abs(ref["G"] / ref["S0"]))
    pih = np.concatenate([h["pi"] for h in hosts])
    assert np.max(np.abs(pih - ref["pi"])) <= tol * np.max(ref["pi"])
    # deep variant against the oracle on the unsharded vectors
    rng = np.random.default_rng(3)
    resid = torch.from_numpy(rng.exponential(1.0, size=n).astype(np.float32) + 0.25)
    w = torch.ones(n)
    k = deep_ref.update_sample_weights(resid, w)
    assert all(g[2]["iters"] == k for g in got)
    assert np.max(np.abs(np.concatenate([g[3] for g in got]) - resid.numpy())) < 1e-5
    assert np.max(np.abs(np.concatenate([g[4] for g in got]) - w.numpy())) < 1e-5
