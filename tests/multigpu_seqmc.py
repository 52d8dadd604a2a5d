"""Run under torchrun with 2 GPUs (tests/test_gpu_multi.py launches it): a SeqMC population sharded over the ranks
(ncclAllGather of particle states and weights per target, every rank resampling its own slots from the global
weights) must reproduce the single-GPU run of the same population bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mcmc_jl_b200 as mj
from mcmc_jl_b200 import _capi as capi


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = mj.init_row_sharding()          # the library's communicator (shared by row sharding and SeqMC)
    ctx = mj.default_context()
    solo = capi.Context(local)                    # no communicator: the whole population on one GPU
    nt, d, npl, steps, burnin = 6, 2, 192, 5, 1
    sig = np.logspace(0.8, -0.5, nt)
    hypers = [(0.0, float(s)) for s in sig]
    smp = [capi.sampler_cfg("RWM", scale=float(s)) if t % 2 == 0 else capi.sampler_cfg("HMC", scale=0.3 * float(s), nleaps=2) for t, s in enumerate(sig)]
    rng = np.random.default_rng(11)
    gp = npl * world
    parts = rng.standard_normal((gp, d))
    zn = rng.standard_normal((steps, nt, gp, d)); un = rng.random((steps, nt, gp)); ru = rng.random((steps, nt, gp))
    for trigger in (1e300, 0.3, 1e-10):
        for inj in (True, False):
            kw = dict(normals=zn, uniforms=un, res_uniforms=ru) if inj else {}
            ref = solo.run_seqmc("normal_dsl", d, hypers, smp, steps, burnin, trigger, parts, seed=9, **kw)
            out = ctx.run_seqmc("normal_dsl", d, hypers, smp, steps, burnin, trigger, parts[rank * npl:(rank + 1) * npl], seed=9, **kw)
            assert out["n_resamples"] == ref["n_resamples"], (trigger, inj, out["n_resamples"], ref["n_resamples"])
            rs = ref["samples"].reshape(steps - burnin, gp, d)[:, rank * npl:(rank + 1) * npl].reshape(-1, d)
            rw = ref["weights"].reshape(steps - burnin, gp)[:, rank * npl:(rank + 1) * npl].reshape(-1)
            assert np.array_equal(out["samples"], rs), (trigger, inj)
            assert np.array_equal(out["weights"], rw), (trigger, inj)
    dist.barrier()
    if rank == 0:
        print("SEQMC_SHARDED_OK world=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
