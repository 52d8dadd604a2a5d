import sys, os, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, mcmc_jl_b200
from mcmc_jl_b200 import _capi as capi
ctx = capi.Context(0)
N = 100000
X, y, b0 = bench.synth_logistic(N, 100, 4)
dm = capi.DeviceModel(ctx, "logistic", 100, X, y, (1.0, -1.0))
eps = bench.CFG4_EPS * (1e6 / N) ** 0.5
C = 256
rng = np.random.default_rng(100)
init = b0[None, :] + 2e-3 * rng.standard_normal((C, 100))
scfg = capi.sampler_cfg("HMCDA", len=10 * eps, max_leaps=64)
rr = capi.DeviceRun(dm, scfg, (1051, 1, 1300), C, init, seed=11, engine="wave", store_grad=False, store_logtarget=False)
rr.set_state(1000, np.full(C, eps), np.full(C, eps), np.zeros(C))
info = rr.execute(); print(info)
st = rr.stats("imse")
out = rr.fetch(grads=False, logtarget=False)
print("acc", out["accept"].mean(), "nan samples", np.isnan(out["samples"]).sum(), "ess nan", np.isnan(st["ess"]).sum(), st["ess"].shape)
print("ess min/median", np.nanmin(st["ess"]), np.nanmedian(st["ess"]), "var nan", np.isnan(st["var"]).sum(), "viid<=0", (st["var_iid"] <= 0).sum())
bad = np.argwhere(np.isnan(st["ess"]))[:5]; print(bad)
for c, j in bad[:2]:
    s = out["samples"][c, :, j]; print(c, j, s[:5], s.std(), st["var"][c, j], st["var_iid"][c, j], st["mean"][c, j])
