"""fused per-chain kernel throughput probe (scratch tool)"""
import sys, os, json, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mcmc_jl_b200
from mcmc_jl_b200 import _capi as capi
ctx = capi.Context(0)
d = 3
dm = capi.DeviceModel(ctx, "normal_fn", d)
for C in (65536, 1 << 20):
    for kind, kw in (("RWM", dict(scale=0.5)), ("MALA", dict(scale=0.5)), ("HMC", dict(scale=0.75, nleaps=10))):
        for sg in (False, True):
            steps = 1000 if C == 65536 else 200
            run = capi.DeviceRun(dm, capi.sampler_cfg(kind, **kw), (1, 1, steps), C, np.ones(d), seed=1, engine="fused", store_grad=sg, store_logtarget=sg)
            run.execute(); info = run.execute() if False else run.info
            run.close()
            run = capi.DeviceRun(dm, capi.sampler_cfg(kind, **kw), (1, 1, steps), C, np.ones(d), seed=2, engine="fused", store_grad=sg, store_logtarget=sg)
            info = run.execute(); run.close()
            bytes_per = 8 * d + 1 + (8 * d + 8 if sg else 0)
            print(json.dumps(dict(C=C, kind=kind, store_grad=sg, ms=info["gpu_ms"], chain_steps_per_s=C * steps / info["gpu_ms"] * 1e3,
                                  store_GBs=bytes_per * C * steps / info["gpu_ms"] / 1e6)))
