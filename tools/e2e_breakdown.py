"""Where the end-to-end time of a short run goes (scratch tool): create / execute / fetch / close, host-timed."""
import sys, os, json, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mcmc_jl_b200, torch
from mcmc_jl_b200 import _capi as capi
fam, N, d, C, kind = (sys.argv[1:] + ["probit", "100000", "20", "16384", "MALA"])[:5]
N, d, C = int(N), int(d), int(C)
ctx = capi.Context(0)
if os.environ.get("TIME_EVAL"): ctx.set_option("time_eval", 1)
r = np.random.default_rng(3)
X = r.standard_normal((d, N)).T; X[:, 0] = 1.0
b0 = r.standard_normal(d) / np.sqrt(d)
y = (r.random(N) < 0.5).astype(float)
hy = {"logistic": (1.0, -1.0), "probit": (10.0,)}[fam]
dm = capi.DeviceModel(ctx, fam, d, X, y, hy)
kw = dict(MALA=dict(scale=1e-6), HMC=dict(scale=1e-3, nleaps=10), HMCDA=dict(len=0.02))[kind]
init = np.tile(b0, (C, 1))
K = 5
pin = lambda shape, dt=torch.float64: torch.empty(shape, dtype=dt).pin_memory().numpy()
bufs = dict(samples=pin((C, K, d)), grads=pin((C, K, d)), accept=pin((C, K), torch.uint8), logtarget=pin((C, K)))
for rep in range(3):
    torch.cuda.synchronize(); t = [time.perf_counter()]
    run = capi.DeviceRun(dm, capi.sampler_cfg(kind, **kw), (1, 1, K), C, init, seed=rep, engine="wave"); t.append(time.perf_counter())
    info = run.execute(); t.append(time.perf_counter())
    run.fetch(out=bufs); torch.cuda.synchronize(); t.append(time.perf_counter())
    run.close(); t.append(time.perf_counter())
    print(json.dumps(dict(rep=rep, create_ms=(t[1] - t[0]) * 1e3, execute_ms=(t[2] - t[1]) * 1e3, gpu_ms=info["gpu_ms"], waves=info["n_waves"],
                          fetch_ms=(t[3] - t[2]) * 1e3, close_ms=(t[4] - t[3]) * 1e3)))
