"""Launch-bound regime probe (scratch tool): the smallN workload of bench.py for a few steps, stream launches (no graph), so that
an ncu launch list shows k1_kernel and transition_kernel side by side.  usage: smalln_probe.py [steps] [chains] [graphs 0|1]"""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import mcmc_jl_b200  # noqa
from mcmc_jl_b200 import _capi as capi
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
wl = dict(bench.WORKLOADS["smallN"])
if len(sys.argv) > 2:
    wl["chains"] = int(sys.argv[2])
graphs = int(sys.argv[3]) if len(sys.argv) > 3 else 0
ctx = capi.Context(0)
ctx.set_option("use_graphs", graphs)
X, y, hy, b0 = bench.make_problem(wl)
dm = capi.DeviceModel(ctx, "logistic", wl["d"], X, y, hy)
C, d = wl["chains"], wl["d"]
init = b0[None, :] + 0.05 * np.random.default_rng(1).standard_normal((C, d))
r = capi.DeviceRun(dm, bench.sampler_for(wl, capi), (1, 1, 3 + steps), C, init, seed=6, engine="wave", store_grad=False, store_logtarget=False)
r.execute_steps(3)
info = r.execute_steps(steps)
print(json.dumps(dict(info, evals_per_s=info["n_grad_evals"] / info["gpu_ms"] * 1e3, ms_per_wave=info["gpu_ms"] / info["n_waves"])))
r.close(); dm.close(); ctx.close()
