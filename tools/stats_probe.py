"""K3 probe (scratch tool): one stats call on device-resident HMC(0.75) draws; usage: stats_probe.py [vtype] [S] [C]"""
import sys, os, json, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mcmc_jl_b200
from mcmc_jl_b200 import _capi as capi
vt = sys.argv[1] if len(sys.argv) > 1 else "imse"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 9000
C = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
ctx = capi.Context(0)
dm = capi.DeviceModel(ctx, "normal_fn", 3)
run = capi.DeviceRun(dm, capi.sampler_cfg("HMC", scale=0.75, nleaps=10), (1, 1, S), C, np.ones(3), seed=1, engine="fused", store_grad=False, store_logtarget=False)
run.execute()
for rep in range(2):
    t0 = time.perf_counter(); st = run.stats(vt); dt = time.perf_counter() - t0
print(json.dumps(dict(vtype=vt, S=S, C=C, ms=dt * 1e3, GBs_one_pass=8.0 * S * 3 * C / dt / 1e9)))
