"""stats kernel (K3) throughput probe (scratch tool)"""
import sys, os, json, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mcmc_jl_b200, torch
from mcmc_jl_b200 import _capi as capi
ctx = capi.Context(0)
d = 3
dm = capi.DeviceModel(ctx, "normal_fn", d)
for C, steps in ((65536, 2000), (65536, 9000), (1 << 18, 2000)):
    run = capi.DeviceRun(dm, capi.sampler_cfg("HMC", scale=0.75, nleaps=10), (1, 1, steps), C, np.ones(d), seed=1, engine="fused", store_grad=False, store_logtarget=False)
    run.execute()
    for vt in ("iid", "bm", "imse", "ipse"):
        torch.cuda.synchronize(); t0 = time.perf_counter(); st = run.stats(vt); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        byt = 8.0 * steps * d * C
        print(json.dumps(dict(C=C, S=steps, vtype=vt, ms=dt * 1e3, GBs_one_pass=byt / dt / 1e9, ess_med=float(np.median(st.get("ess", [0]))))))
    run.close()
