"""First on-GPU check: parity of every path against the oracle on small cases (scratch tool)."""
import importlib.util, sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O
spec = importlib.util.spec_from_file_location("capi", os.path.join(ROOT, "mcmc.jl_b200", "_capi.py"))
capi = importlib.util.module_from_spec(spec); spec.loader.exec_module(capi)

ctx = capi.Context(0)
rng = np.random.default_rng(0)

def cmp(name, a, b, tol=1e-12):
    a = np.asarray(a, float); b = np.asarray(b, float)
    same_nan = np.array_equal(np.isnan(a), np.isnan(b))
    m = ~np.isnan(a)
    err = np.max(np.abs(a[m] - b[m]) / (np.abs(b[m]) + 1e-300)) if m.any() else 0.0
    bit = np.array_equal(a[m], b[m])
    print(f"  {name:28s} maxrel={err:.3e} bitexact={bit} nan_match={same_nan}")
    return err

def data(N, d, fam, seed):
    r = np.random.default_rng(seed)
    X = np.concatenate([np.ones((N, 1)), r.standard_normal((N, d - 1))], axis=1)
    b0 = r.standard_normal(d) / np.sqrt(d)
    eta = X @ b0
    if fam == "linear": y = eta + r.standard_normal(N); hy = (1.0, 1.0)
    elif fam == "logistic": y = (r.random(N) < 1 / (1 + np.exp(-eta))).astype(float); hy = (1.0, -1.0)
    else:
        from scipy.special import ndtr
        y = (r.random(N) < ndtr(eta)).astype(float); hy = (10.0,)
    return X, y, hy, b0

print("== logtarget_grad ==")
for fam in ["linear", "logistic", "probit"]:
    for (N, d, Cn) in [(1000, 10, 70), (333, 3, 5), (4100, 100, 130), (50, 17, 64)]:
        X, y, hy, b0 = data(N, d, fam, 1)
        om = O.Model(fam, d, X, y, hy)
        dm = capi.DeviceModel(ctx, fam, d, X, y, hy)
        B = b0 + 0.3 * rng.standard_normal((Cn, d)) / np.sqrt(d)
        lt, g = dm.logtarget_grad(B)
        olt = np.array([om.evalallg(b)[0] for b in B]); og = np.array([om.evalallg(b)[1] for b in B])
        print(fam, N, d, Cn)
        cmp("lt", lt, olt)
        gs = np.abs(X).T @ np.ones(N)
        print("  grad max abs err / |X|'1:", np.max(np.abs(g - og) / gs))
        dm.close()
for fam, d, hy, ydata in [("normal_fn", 3, (), None), ("normal_dsl", 5, (0.5, 2.0), None), ("ou", 3, (100., 2., 20.), "ou")]:
    if ydata == "ou":
        x = np.empty(1000); x[0] = 1.0
        for i in range(1, 1000): x[i] = x[i-1]*np.exp(-1/20) + 10*(1-np.exp(-1/20)) + 0.1*rng.standard_normal()
        yv = x; B = np.array([[0.05, 1., 1.], [20., 0.1, 10.], [18., 0.11, 9.], [-1, 1, 1], [5, 3, 1]])
    else:
        yv = None; B = rng.standard_normal((9, d))
    om = O.Model(fam, d, None, yv, hy); dm = capi.DeviceModel(ctx, fam, d, None, yv, hy)
    lt, g = dm.logtarget_grad(B)
    olt = np.array([om.evalallg(b)[0] for b in B]); og = np.array([om.evalallg(b)[1] for b in B])
    print(fam); cmp("lt", lt, olt); cmp("grad", g, og)
    dm.close()

print("== runs (injected draws) ==")
def run_case(fam, d, X, y, hy, kind, skw, rngt, Cn, init, engine, scale=None):
    om = O.Model(fam, d, X, y, hy); dm = capi.DeviceModel(ctx, fam, d, X, y, hy)
    first, step, last = rngt
    zn = rng.standard_normal((Cn, last + 1, d)); un = rng.random((Cn, last + 1))
    run = capi.DeviceRun(dm, capi.sampler_cfg(kind, **skw), rngt, Cn, init, scale=scale, normals=zn, uniforms=un, engine=engine)
    info = run.execute(); out = run.fetch()
    print(f"{fam} d={d} {kind} {skw} engine={engine} C={Cn}: {info}")
    osm = O.sampler(kind, **skw)
    worst = 0
    acc_mismatch = 0
    for c in range(Cn):
        ini = init[c] if np.ndim(init) == 2 else init
        ref = O.run_chain(om, osm, rngt, ini, scale, zn[c], un[c])
        acc_mismatch += int(np.sum(ref["accept"] != out["accept"][c]))
        e = np.nanmax(np.abs(ref["samples"] - out["samples"][c]) / (np.abs(ref["samples"]) + 1e-12))
        worst = max(worst, e)
        if c == 0:
            cmp("samples[0]", out["samples"][0], ref["samples"]); cmp("grads[0]", out["grads"][0], ref["grads"]); cmp("lt[0]", out["logtarget"][0], ref["logtarget"])
            if run.has_diag:
                eps, nl = run.fetch_diag(); cmp("eps[0]", eps[0], ref["eps"]); print("   nleaps equal:", np.array_equal(nl[0], ref["nleaps"]))
    print(f"  all chains: accept mismatches={acc_mismatch}, worst sample rel err={worst:.3e}, acc rate={out['accept'].mean():.3f}")
    st = run.stats("imse")
    s0 = out["samples"][0]
    print("  stats ess[0]:", st["ess"][0], "oracle:", [O.ess(s0[:, j]) for j in range(d)][:3], "acc rate", st["accept_rate"][:2])
    run.close(); dm.close()

for engine in ["fused", "wave"]:
    run_case("normal_fn", 3, None, None, (), "RWM", dict(scale=0.1), (101, 1, 1000), 70, np.ones(3), engine)
    run_case("normal_fn", 3, None, None, (), "HMC", dict(scale=0.75, nleaps=10), (101, 2, 1000), 70, np.ones(3), engine)
    run_case("normal_dsl", 4, None, None, (0.0, 1.0), "MALA", dict(scale=0.5), (1, 1, 500), 33, np.ones(4), engine)
    run_case("normal_fn", 3, None, None, (), "HMCDA", dict(len=2.0), (201, 1, 600), 40, np.ones(3), engine)
    run_case("normal_fn", 2, None, None, (), "HMC", dict(scale=0.3, nleaps=5, tuner=dict(target_rate=0.7, adapt_step=50)), (201, 1, 600), 10, np.ones(2), engine)
    run_case("normal_dsl", 2, None, None, (0.0, 1.0), "MALA", dict(scale=0.3, tuner=dict(target_rate=0.5, adapt_step=50)), (201, 1, 600), 10, np.ones(2), engine)
x = np.empty(300); x[0] = 1.0
for i in range(1, 300): x[i] = x[i-1]*np.exp(-1/20) + 10*(1-np.exp(-1/20)) + 0.1*rng.standard_normal()
for engine in ["fused", "wave"]:
    run_case("ou", 3, None, x, (100., 2., 20.), "RWM", dict(scale=0.01), (1, 1, 300), 20, np.array([20., 0.1, 10.]), engine, scale=np.array([1000., 1., 10.]))
    run_case("ou", 3, None, x, (100., 2., 20.), "HMC", dict(scale=0.002, nleaps=5), (1, 1, 200), 20, np.array([20., 0.1, 10.]), engine)
for fam in ["linear", "logistic", "probit"]:
    X, y, hy, b0 = data(1000, 10, fam, 2)
    run_case(fam, 10, X, y, hy, "RWM", dict(scale=0.02), (1, 1, 200), 70, np.zeros(10), "wave")
    run_case(fam, 10, X, y, hy, "MALA", dict(scale=0.001), (1, 1, 200), 70, np.zeros(10), "wave")
    run_case(fam, 10, X, y, hy, "HMC", dict(scale=0.02, nleaps=4), (51, 1, 150), 70, np.zeros(10), "wave")
    run_case(fam, 10, X, y, hy, "HMCDA", dict(len=0.2), (101, 1, 200), 20, np.zeros(10), "wave")
print("== philox ==")
z, u = ctx.philox_draws(12345, 7, 3, 5, 4)
zo = np.array([[O.draw_normals(12345, 7 + c, i, 5) for i in range(5)] for c in range(3)])
uo = np.array([[O.draw_uniform(12345, 7 + c, i) for i in range(5)] for c in range(3)])
print("normals max abs diff", np.abs(z - zo).max(), "uniform bit-exact", np.array_equal(u, uo))
print("DONE")
