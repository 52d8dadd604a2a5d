"""Pilot adaptation for the headline workload (BASELINE config 4): HMCDA on the synthetic logistic regression
N=1e6, d=100 from init = 0 (the example's `vars=zeros`), to find the adapted dual-averaged step size that
bench.py restores.  Writes profiles/pilot_cfg4.json."""
import importlib.util, sys, os, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import mcmc_jl_b200  # noqa
from mcmc_jl_b200 import _capi as capi
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
L = float(sys.argv[2]) if len(sys.argv) > 2 else 0.02
C = 512
ctx = capi.Context(0)
X, y, b0 = bench.synth_logistic(N, 100, 4)
dm = capi.DeviceModel(ctx, "logistic", 100, X, y, (1.0, -1.0))
burn, steps = 150, 200
run = capi.DeviceRun(dm, capi.sampler_cfg("HMCDA", len=L, max_leaps=64), (burn + 1, 1, steps), C, np.zeros(100), seed=4,
                     engine="wave", store_grad=False)
t0 = time.time(); info = run.execute(); dt = time.time() - t0
eps, nl = run.fetch_diag(); out = run.fetch(grads=False)
st = run.get_state()
res = dict(N=N, len=L, chains=C, burnin=burn, steps=steps, seconds=dt, info=info,
           eps_quantiles=np.quantile(st["dual_leapstep"], [0.05, 0.25, 0.5, 0.75, 0.95]).tolist(),
           nleaps_quantiles=np.quantile(nl[:, -1], [0.05, 0.5, 0.95]).tolist(),
           accept_rate_post_burnin=float(out["accept"].mean()),
           dist_to_truth=float(np.median(np.linalg.norm(st["pars"] - b0, axis=1))))
print(json.dumps(res))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "pilot_cfg4.json"), "w"), indent=1)
