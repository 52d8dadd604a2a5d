"""Generates the golden fixtures under tests/golden/ (committed together with this script).

Sources of truth, none of them the product and none of them the C oracle:
  * model log-targets / gradients: the mathematical definitions of README.md:60-72 and examples/*.jl
    evaluated with mpmath at 60 digits (stable closed forms: log-sigmoid, log-Phi via mp.ncdf, Mills
    ratio), on seeded, well-conditioned inputs;
  * probit on the vaso-constriction data the reference ships (examples/vaso.txt, 39 x 3), prepared as
    examples/probit_regression.jl:7-16 does (standardise, add intercept); read from /root/reference
    only here, at generation time;
  * stats: src/stats/var.jl + ess.jl restated independently in mpmath.
Sampler trajectories are pinned in golden_chains.npz from the C oracle itself (a regression pin of the
oracle plus a fixed target for the GPU tests) -- labelled "oracle-generated" in the file.

Run:  python tests/golden/make_golden.py
"""
import os
import sys

import mpmath as mp
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
mp.mp.dps = 60


def f(x):
    return float(x)


def normal_logpdf(x, mu, sd):
    z = (x - mu) / sd
    return -(mp.log(2 * mp.pi) / 2 + z * z / 2 + mp.log(sd))


def fam_truth(fam, X, y, hy, beta):
    b = [mp.mpf(float(v)) for v in beta]
    d = len(b)
    if fam == "normal_fn":
        return f(-sum(v * v for v in b)), [f(-2 * v) for v in b]
    if fam == "normal_dsl":
        mu, sd = mp.mpf(hy[0]), mp.mpf(hy[1])
        return f(sum(normal_logpdf(v, mu, sd) for v in b)), [f((mu - v) / sd ** 2) for v in b]
    if fam == "ou":
        x = [mp.mpf(float(v)) for v in y]
        tau, sig, mu = b
        if not (0 <= tau <= hy[0] and 0 <= sig <= hy[1] and 0 <= mu <= hy[2]):
            return float("-inf"), [0.0, 0.0, 0.0]
        fac = mp.e ** (-1 / tau)
        lt = -mp.log(hy[0]) - mp.log(hy[1]) - mp.log(hy[2])
        dfac = dsig = dmu = mp.mpf(0)
        for t in range(len(x) - 1):
            r = x[t + 1] - x[t] * fac - mu * (1 - fac)
            lt += normal_logpdf(r, 0, sig)
            dres = -r / sig ** 2
            dsig += (r * r / sig ** 2 - 1) / sig
            dfac += dres * (mu - x[t])
            dmu += dres * (-(1 - fac))
        return f(lt), [f(dfac * fac / tau ** 2), f(dsig), f(dmu)]
    N = X.shape[0]
    Xm = [[mp.mpf(float(X[i, j])) for j in range(d)] for i in range(N)]
    ym = [mp.mpf(float(v)) for v in y]
    eta = [sum(Xm[i][j] * b[j] for j in range(d)) for i in range(N)]
    if fam == "linear":
        psd, nsd = mp.mpf(hy[0]), mp.mpf(hy[1])
        lt = sum(normal_logpdf(v, 0, psd) for v in b) + sum(normal_logpdf(ym[i] - eta[i], 0, nsd) for i in range(N))
        r = [(ym[i] - eta[i]) / nsd ** 2 for i in range(N)]
        g = [sum(Xm[i][j] * r[i] for i in range(N)) - b[j] / psd ** 2 for j in range(d)]
        return f(lt), [f(v) for v in g]
    if fam == "logistic":
        psd, sgn = mp.mpf(hy[0]), mp.mpf(hy[1])
        lt = sum(normal_logpdf(v, 0, psd) for v in b)
        r = []
        for i in range(N):
            e = mp.e ** (sgn * eta[i])
            p = 1 / (1 + e)
            if ym[i] != 0:
                lt += mp.log(p); r.append(-sgn * e * p)
            else:
                lt += mp.log(e * p); r.append(sgn * p)
        g = [sum(Xm[i][j] * r[i] for i in range(N)) - b[j] / psd ** 2 for j in range(d)]
        return f(lt), [f(v) for v in g]
    if fam == "probit":
        pvar = mp.mpf(hy[0]) ** 2
        lt = -(d * mp.log(2 * mp.pi) + d * mp.log(pvar)) / 2 - sum(v * v for v in b) / (2 * pvar)
        r = []
        for i in range(N):
            lp, lm = mp.log(mp.ncdf(eta[i])), mp.log(mp.ncdf(-eta[i]))
            lt += lp * ym[i] + lm * (1 - ym[i])
            phi = mp.npdf(eta[i])
            r.append(ym[i] * phi / mp.ncdf(eta[i]) - (1 - ym[i]) * phi / mp.ncdf(-eta[i]))
        g = [sum(Xm[i][j] * r[i] for i in range(N)) - b[j] / pvar for j in range(d)]
        return f(lt), [f(v) for v in g]
    raise ValueError(fam)


def make_models():
    rng = np.random.default_rng(20261018)
    out = {}
    N, d = 40, 4
    X = np.concatenate([np.ones((N, 1)), rng.standard_normal((N, d - 1))], axis=1)
    b0 = rng.standard_normal(d) * 0.7
    eta = X @ b0
    cases = {
        "normal_fn": (None, None, (), 3),
        "normal_dsl": (None, None, (0.25, 1.5), 3),
        "linear": (X, eta + rng.standard_normal(N), (1.0, 1.0), d),
        "logistic": (X, (rng.random(N) < 1 / (1 + np.exp(-eta))).astype(float), (1.0, -1.0), d),
        "logistic_plus": (X, (rng.random(N) < 1 / (1 + np.exp(eta))).astype(float), (1.0, 1.0), d),  # test/test_syntax.jl:18
        "probit": (X, (rng.random(N) < 0.5 * (1 + np.vectorize(lambda t: float(mp.erf(t / mp.sqrt(2))))(eta))).astype(float), (10.0,), d),
    }
    x = np.empty(60); x[0] = 1.0
    for i in range(1, 60):
        x[i] = x[i - 1] * np.exp(-1 / 20) + 10 * (1 - np.exp(-1 / 20)) + 0.1 * rng.standard_normal()
    cases["ou"] = (None, x, (100.0, 2.0, 20.0), 3)
    for name, (Xc, yc, hy, dd) in cases.items():
        fam = "logistic" if name == "logistic_plus" else name
        if fam == "ou":
            B = np.array([[20.0, 0.1, 10.0], [18.0, 0.12, 9.5], [0.05, 1.0, 1.0], [25.0, 0.3, 11.0], [-1.0, 1.0, 1.0], [5.0, 2.5, 1.0]])
        else:
            B = (b0[:dd] if dd == d else np.zeros(dd)) + 0.4 * rng.standard_normal((5, dd))
        lts, gs = [], []
        for bvec in B:
            lt, g = fam_truth(fam, Xc, yc, hy, bvec)
            lts.append(lt); gs.append(g)
        out[name] = dict(family=fam, X=Xc, y=yc, hyper=np.array(hy, dtype=float), B=B, lt=np.array(lts), grad=np.array(gs))
    # vaso probit (examples/probit_regression.jl:7-16)
    vaso_path = "/root/reference/examples/vaso.txt"
    if os.path.exists(vaso_path):
        vaso = np.loadtxt(vaso_path)
        cov, yv = vaso[:, :-1], vaso[:, -1]
        cov = (cov - cov.mean(0)) / cov.std(0, ddof=1)
        Xv = np.concatenate([np.ones((cov.shape[0], 1)), cov], axis=1)
        B = np.array([[0.0, 0.0, 0.0], [-0.5, 1.0, 1.5], [0.3, -0.2, 0.1], [-1.0, 2.0, 2.5]])
        lts, gs = [], []
        for bvec in B:
            lt, g = fam_truth("probit", Xv, yv, (10.0,), bvec)
            lts.append(lt); gs.append(g)
        out["probit_vaso"] = dict(family="probit", X=Xv, y=yv, hyper=np.array([10.0]), B=B, lt=np.array(lts), grad=np.array(gs))
    else:
        old = np.load(os.path.join(HERE, "golden_models.npz"), allow_pickle=True)["cases"].item()
        out["probit_vaso"] = old["probit_vaso"]
    np.savez(os.path.join(HERE, "golden_models.npz"), cases=np.array(out, dtype=object))
    print("models:", list(out))


def stats_truth(x, vtype, maxlag=None, batchlen=100):
    xm = [mp.mpf(float(v)) for v in x]
    n = len(xm)
    mu = sum(xm) / n
    viid = sum((v - mu) ** 2 for v in xm) / (n - 1) / n
    if vtype == "iid":
        return f(viid)
    if vtype == "bm":
        nb = n // batchlen
        bmeans = [sum(xm[j * batchlen:(j + 1) * batchlen]) / batchlen for j in range(nb)]
        mb = sum(bmeans) / nb
        return f(batchlen * (sum((v - mb) ** 2 for v in bmeans) / (nb - 1)) / (nb * batchlen))
    maxlag = n - 1 if maxlag is None else maxlag
    k = (maxlag - 1) // 2
    acv = lambda lag: sum((xm[t] - mu) * (xm[t + lag] - mu) for t in range(n - lag)) / n
    g, m = [], k + 1
    for j in range(k + 1):
        gj = acv(2 * j) + acv(2 * j + 1)
        g.append(gj)
        if gj <= 0:
            m = j
            break
    if vtype == "imse":
        for j in range(1, m):
            if g[j] > g[j - 1]:
                g[j] = g[j - 1]
    return f((-acv(0) + 2 * sum(g[:m])) / n)


def make_stats():
    rng = np.random.default_rng(7)
    series = {}
    n = 400
    for name, rho in [("ar_pos", 0.6), ("ar_neg", -0.5), ("iid", 0.0), ("ar_strong", 0.95)]:
        x = np.empty(n); x[0] = rng.standard_normal()
        for t in range(1, n):
            x[t] = rho * x[t - 1] + rng.standard_normal()
        series[name] = x
    out = {}
    for name, x in series.items():
        out[name] = dict(x=x, iid=stats_truth(x, "iid"), bm=stats_truth(x, "bm", batchlen=20), bm_len=20,
                         imse=stats_truth(x, "imse"), ipse=stats_truth(x, "ipse"),
                         imse_lag21=stats_truth(x, "imse", maxlag=21), mean=f(sum(mp.mpf(float(v)) for v in x) / n))
    np.savez(os.path.join(HERE, "golden_stats.npz"), cases=np.array(out, dtype=object))
    print("stats:", list(out))


def make_chains():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    out = {}
    d = 3
    m = O.Model("normal_fn", d)
    cfgs = {
        "rwm": ("RWM", dict(scale=0.1), (101, 1, 1000)),              # BASELINE config 1
        "mala": ("MALA", dict(scale=0.3), (1, 2, 400)),
        "hmc": ("HMC", dict(scale=0.75, nleaps=10), (101, 1, 500)),    # README.md:104
        "hmcda": ("HMCDA", dict(len=2.0), (1, 1, 25)),
        "hmc_tuned": ("HMC", dict(scale=0.3, nleaps=5, tuner=dict(target_rate=0.7, adapt_step=50)), (201, 1, 400)),
    }
    for name, (kind, kw, rngt) in cfgs.items():
        rng = np.random.default_rng(abs(hash(name)) % 1000 + 11)
        seed = int(rng.integers(1 << 30))
        r2 = np.random.default_rng(seed)
        zn = r2.standard_normal((rngt[2] + 1, d)); un = r2.random(rngt[2] + 1)
        res = O.run_chain(m, O.sampler(kind, **kw), rngt, np.ones(d), None, zn, un)
        out[name] = dict(kind=kind, kw=kw, range=rngt, draw_seed=seed, samples=res["samples"], accept=res["accept"],
                         logtarget=res["logtarget"], eps=res["eps"], nleaps=res["nleaps"], source="oracle-generated")
    np.savez(os.path.join(HERE, "golden_chains.npz"), cases=np.array(out, dtype=object))
    print("chains:", list(out))


if __name__ == "__main__":
    make_models()
    make_stats()
    make_chains()
