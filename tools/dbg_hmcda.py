import importlib.util, sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O
spec = importlib.util.spec_from_file_location("capi", os.path.join(ROOT, "mcmc.jl_b200", "_capi.py"))
capi = importlib.util.module_from_spec(spec); spec.loader.exec_module(capi)
ctx = capi.Context(0)
rng = np.random.default_rng(5)
d = 3; last = 30
zn = rng.standard_normal((1, last + 1, d)); un = rng.random((1, last + 1))
om = O.Model("normal_fn", d); dm = capi.DeviceModel(ctx, "normal_fn", d)
for first in (1, 2, 3, 5, 21):
    for eng in ("fused", "wave"):
        run = capi.DeviceRun(dm, capi.sampler_cfg("HMCDA", len=2.0), (first, 1, last), 1, np.ones(d), normals=zn, uniforms=un, engine=eng)
        run.execute(); out = run.fetch(); eps, nl = run.fetch_diag()
        ref = O.run_chain(om, O.sampler("HMCDA", len=2.0), (first, 1, last), np.ones(d), None, zn[0], un[0])
        print(first, eng, "eps gpu", eps[0][:6], "ora", ref["eps"][:6])
        print("   nl gpu", nl[0][:8], "ora", ref["nleaps"][:8], "acc", out["accept"][0][:8], ref["accept"][:8])
        run.close()
