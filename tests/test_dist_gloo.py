"""CPU suite, part 3: the multi-process host logic with world_size 2 over gloo (no GPU): chain-shard slices,
row-shard ranges, the NCCL-unique-id broadcast, and the arithmetic identity the row-sharded path relies on
(sum of per-shard partial log-likelihood / gradient + prior once == unsharded), checked with the oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import mcmc_jl_b200 as mj
    from mcmc_jl_b200 import api, _capi
    import oracle as O
    from conftest import make_regression
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # 1. chain sharding: slices are disjoint, ordered, and cover [0, C)
    lo, n = api._rank_slice(10001)
    t = torch.tensor([lo, n]); gathered = [torch.zeros(2, dtype=torch.long) for _ in range(world)]
    dist.all_gather(gathered, t)
    spans = [(int(a), int(b)) for a, b in gathered]
    assert spans[0][0] == 0 and all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    assert spans[-1][0] + spans[-1][1] == 10001
    # 2. the NCCL unique id travels over whatever backend torch.distributed uses
    uid = mj.broadcast_unique_id(dist, _capi.Context.comm_unique_id, rank)
    ids = [None] * world
    dist.all_gather_object(ids, uid)
    assert len(uid) == 128 and all(i == ids[0] for i in ids)
    # 3. row-sharded evaluation identity: partial sums all-reduced, prior added once
    N, d = 1001, 6
    X, y, hy, b0 = make_regression("logistic", N, d, 3)
    rlo, rhi = mj.shard_rows(N, rank, world)
    flat = (1e300, hy[1])                          # prior sd -> infinity: the shard model's log-target is its partial loglik (+ const)
    om_full, om_shard = O.Model("logistic", d, X, y, hy), O.Model("logistic", d, X[rlo:rhi], y[rlo:rhi], flat)
    om_none = O.Model("logistic", d, X[:0].reshape(0, d), y[:0], flat)
    beta = b0 + 0.1
    c0, _ = om_none.evalallg(beta)                 # the flat prior's constant
    pl, pg = om_shard.evalallg(beta)
    part = torch.tensor(np.concatenate([[pl - c0], pg]))
    dist.all_reduce(part)                          # what ncclAllReduce does on the device
    prior = sum(-(0.9189385332046727 + 0.5 * (b / hy[0]) ** 2 + np.log(hy[0])) for b in beta)
    lt, g = om_full.evalallg(beta)
    assert abs((part[0].item() + prior) - lt) <= 1e-10 * abs(lt)
    assert np.allclose(part[1:].numpy() - beta / hy[0] ** 2, g, rtol=1e-9, atol=1e-9)
    dist.barrier()
    dist.destroy_process_group()
    q.put(rank)


def test_world_size_two_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    for p in ps:
        p.join(timeout=180)
    assert all(p.exitcode == 0 for p in ps), [p.exitcode for p in ps]
    assert sorted(q.get() for _ in range(2)) == [0, 1]
