"""Run under torchrun with >= 2 GPUs (tests/test_gpu_multi.py launches it): row-sharded logistic regression
(NCCL all-reduce of partial loglik/gradient per evaluation) must reproduce the unsharded single-GPU result:
log-target / gradient within 1e-12 (reduction order differs), identical accept flags, every rank identical."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mcmc_jl_b200 as mj
from mcmc_jl_b200 import _capi as capi
from conftest import make_regression


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = mj.init_row_sharding()
    ctx = mj.default_context()
    N, d, C = 5003, 24, 96
    X, y, hy, b0 = make_regression("logistic", N, d, 5)
    lo, hi = mj.shard_rows(N, rank, world)
    full = capi.DeviceModel(ctx, "logistic", d, X, y, hy)
    shard = capi.DeviceModel(ctx, "logistic", d, X[lo:hi], y[lo:hi], hy, row_sharded=True)
    rng = np.random.default_rng(1)
    B = b0 + 0.2 * rng.standard_normal((C, d))
    lt0, g0 = full.logtarget_grad(B)
    lt1, g1 = shard.logtarget_grad(B)
    assert np.all(np.abs(lt1 - lt0) <= 1e-12 * np.abs(lt0)), np.abs(lt1 - lt0).max()
    assert np.all(np.abs(g1 - g0) <= 1e-12 * np.abs(X).sum(0))
    zn = rng.standard_normal((C, 41, d)); un = rng.random((C, 41))
    outs = []
    for m in (full, shard):
        for kind, kw in (("HMC", dict(scale=0.01, nleaps=5)), ("MALA", dict(scale=2e-4)), ("RWM", dict(scale=0.005))):
            run = capi.DeviceRun(m, capi.sampler_cfg(kind, **kw), (1, 1, 40), C, np.zeros(d), normals=zn, uniforms=un, engine="wave")
            run.execute(); outs.append(run.fetch()); run.close()
    for a, b in zip(outs[:3], outs[3:]):
        assert np.array_equal(a["accept"], b["accept"])
        assert np.allclose(a["samples"], b["samples"], rtol=1e-9, atol=1e-12)
    # all ranks hold identical chains (same Philox keys, same reduced sums)
    t = torch.from_numpy(outs[3]["samples"]).cuda()
    ref = t.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(t, ref)
    full.close(); shard.close()
    dist.barrier()
    if rank == 0:
        print("ROWSHARD_OK world=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
