"""BASELINE config 2 at the README's full shape: HMC(0.75) on -dot(v,v), 65 536 chains x 10 000 steps, 1000 burn-in
(README.md:104-204 reports one unseeded chain: acceptance 79.76 %, ESS 5333.5/9000, IAT 1.687, var IMSE 9.27e-5,
iid 5.49e-5, BM 9.14e-5).  Draws stay on the device; only the per-chain statistics come back."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mcmc_jl_b200
from mcmc_jl_b200 import _capi as capi
ctx = capi.Context(0)
dm = capi.DeviceModel(ctx, "normal_fn", 3)
C = 65536
run = capi.DeviceRun(dm, capi.sampler_cfg("HMC", scale=0.75, nleaps=10), (1001, 1, 10000), C, np.ones(3), seed=1, engine="fused",
                     store_grad=False, store_logtarget=False)
t0 = time.time(); info = run.execute(); t1 = time.time()
res = {"info": info, "execute_s": t1 - t0, "chain_steps_per_s": C * 10000 / (info["gpu_ms"] / 1e3)}
for vt in ("imse", "ipse", "bm"):
    t0 = time.time(); st = run.stats(vt); dt = time.time() - t0
    res[vt] = dict(seconds=dt, var_median=float(np.median(st["var"])), ess_median=float(np.median(st["ess"])), ess_mean=float(st["ess"].mean()),
                   actime_median=float(np.median(st["actime"])))
res["acceptance_pct_mean"] = float(st["accept_rate"].mean()); res["acceptance_pct_sd_across_chains"] = float(st["accept_rate"].std())
res["var_iid_median"] = float(np.median(st["var_iid"])); res["mean_abs_max"] = float(np.abs(st["mean"]).max())
res["readme"] = dict(acceptance=79.7556, ess=5333.5, actime=1.687, var_imse=9.27e-5, var_iid=5.49e-5, var_bm=9.14e-5)
print(json.dumps(res, indent=1))
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "readme_band_cfg2.json"), "w"), indent=1)
