"""K1 throughput probe: logistic regression, HMC fixed leapfrogs (scratch tool)."""
import importlib.util, sys, os, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("capi", os.path.join(ROOT, "mcmc.jl_b200", "_capi.py"))
capi = importlib.util.module_from_spec(spec); spec.loader.exec_module(capi)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
d = int(sys.argv[2]) if len(sys.argv) > 2 else 100
C = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
fam = sys.argv[4] if len(sys.argv) > 4 else "logistic"
splits = int(sys.argv[5]) if len(sys.argv) > 5 else 0
ctx = capi.Context(0)
ctx.set_option("time_eval", 1)
if splits: ctx.set_option("force_splits", splits)
if os.environ.get("K1_DEBUG"): ctx.set_option("k1_debug", int(os.environ["K1_DEBUG"]))
r = np.random.default_rng(4)
t0 = time.time()
X = r.standard_normal((d, N)).T  # F-ordered view
X[:, 0] = 1.0
b0 = r.standard_normal(d) / np.sqrt(d)
eta = X @ b0
y = (r.random(N) < 1 / (1 + np.exp(-eta))).astype(float)
print("gen", time.time() - t0)
hy = {"logistic": (1.0, -1.0), "probit": (10.0,), "linear": (1.0, 1.0)}[fam]
t0 = time.time(); dm = capi.DeviceModel(ctx, fam, d, X, y, hy); print("model create", time.time() - t0)
for kind, kw, last in [("HMC", dict(scale=1e-3, nleaps=8), 3), ("RWM", dict(scale=1e-3), 6), ("MALA", dict(scale=1e-6), 6)]:
    init = np.tile(b0, (C, 1)) + 1e-3 * r.standard_normal((C, d))
    run = capi.DeviceRun(dm, capi.sampler_cfg(kind, **kw), (1, 1, last), C, init, seed=1, engine="wave", store_grad=False, store_logtarget=False)
    info = run.execute()
    nw = info["n_waves"]
    per = info["eval_ms"] / nw
    flops = 4.0 * N * d * C if kind != "RWM" else 2.0 * N * d * C
    print(json.dumps(dict(kind=kind, N=N, d=d, C=C, fam=fam, **info, eval_ms_per_wave=per, tflops=flops / per / 1e9,
                          acc=float(run.fetch(samples=False)["accept"].mean()))))
    run.close()
