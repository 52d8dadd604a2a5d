"""CPU suite, part 1: pins the oracle (oracle/mcmc_oracle.c) against known-answer vectors, mpmath golden
fixtures, finite differences, independent restatements and the README's statistical band.
"parity unpinned" (no numeric golden vectors exist in the reference's own tests): these
are the anchors that stand in."""
import math

import numpy as np
import pytest

from conftest import load_golden, make_regression, ou_series


def test_philox_known_answer_vectors(O):
    # Random123 kat_vectors for philox4x32_10
    assert O.philox((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert O.philox((0xffffffff,) * 4, (0xffffffff,) * 2) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert O.philox((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_philox_draw_moments(O):
    z = np.array([O.draw_normals(11, c, 3, 4) for c in range(4000)]).ravel()
    u = np.array([O.draw_uniform(11, c, 3) for c in range(4000)])
    assert abs(z.mean()) < 4 / math.sqrt(z.size) and abs(z.var() - 1) < 0.05
    assert 0 < u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 0.02


@pytest.mark.parametrize("name", ["normal_fn", "normal_dsl", "linear", "logistic", "logistic_plus", "probit", "ou", "probit_vaso"])
def test_models_against_mpmath_golden(O, name):
    g = load_golden("golden_models.npz")[name]
    d = g["B"].shape[1]
    m = O.Model(g["family"], d, g["X"], g["y"], g["hyper"])
    for b, lt, gr in zip(g["B"], g["lt"], g["grad"]):
        olt, og = m.evalallg(b)
        assert m.eval(b) == olt
        if np.isinf(lt):
            assert olt == lt and np.all(og == 0)       # LLAcc: (-Inf, zeros), modelparser.jl:64-72
            continue
        assert abs(olt - lt) <= 1e-12 * max(1.0, abs(lt))
        scale = 1.0 if g["X"] is None else np.abs(g["X"]).sum(0)
        assert np.all(np.abs(og - gr) <= 1e-11 * np.maximum(scale, np.abs(gr)))


@pytest.mark.parametrize("fam", ["linear", "logistic", "probit"])
def test_gradient_finite_difference(O, fam):
    # the reference's own idea (test/dsl/helper_diff.jl:15-37), with a central difference and a real tolerance
    X, y, hy, b0 = make_regression(fam, 200, 6, 5)
    m = O.Model(fam, 6, X, y, hy)
    lt, g = m.evalallg(b0)
    for j in range(6):
        h = 1e-6
        e = np.zeros(6); e[j] = h
        fd = (m.eval(b0 + e) - m.eval(b0 - e)) / (2 * h)
        assert abs(fd - g[j]) <= 1e-6 * max(1.0, abs(g[j]))


def test_llacc_out_of_support_semantics(O):
    X, y, hy, b0 = make_regression("logistic", 50, 3, 1)
    m = O.Model("logistic", 3, X, y, hy)
    lt, g = m.evalallg(np.array([0.0, 800.0, 0.0]))          # p underflows to 0/1 -> log -> -Inf
    assert lt == -np.inf and np.all(g == 0)                   # AccumulatorDerivRules.jl:10-20
    lt, g = m.evalallg(np.array([np.nan, 0.0, 0.0]))
    assert lt == -np.inf and np.all(g == 0)
    ou = O.Model("ou", 3, None, ou_series(50, 1), (100.0, 2.0, 20.0))
    for bad in ([-1.0, 1.0, 1.0], [5.0, 2.5, 1.0], [5.0, 1.0, 21.0], [101.0, 1.0, 1.0]):
        lt, g = ou.evalallg(np.array(bad))
        assert lt == -np.inf and np.all(g == 0)
    # probit is a plain user function: 0 * -Inf gives NaN, nothing is caught (probit_regression.jl:29)
    Xp, yp, hyp, _ = make_regression("probit", 30, 2, 2)
    pm = O.Model("probit", 2, Xp, yp, hyp)
    assert np.isnan(pm.eval(np.array([1e200, 1e200])))


def test_log_ndtr_against_scipy(O):
    from scipy.special import log_ndtr
    xs = np.linspace(-150, 30, 3601)
    v = np.array([O.lib().orc_log_ndtr(x) for x in xs])
    w = log_ndtr(xs)
    assert np.all(np.abs(v - w) <= 1e-13 * np.abs(w) + 1e-300)


def _py_rwm(eval_fn, init, scale, z, u, first, step, last):
    """independent restatement of RWM.jl:58-71 + SerialMC.jl:47-66"""
    pars, lt, out, acc = init.copy(), eval_fn(init), [], []
    for i in range(1, last + 1):
        prop = pars + z[i] * scale
        plt = eval_fn(prop)
        ratio = plt - lt
        a = ratio > 0 or ratio > math.log(u[i])
        if a:
            pars, lt = prop, plt
        if i >= first and (i - first) % step == 0:
            out.append(pars.copy()); acc.append(a)
    return np.array(out), np.array(acc)


def test_rwm_and_serialmc_range_semantics(O):
    rng = np.random.default_rng(3)
    d, last = 3, 300
    z = rng.standard_normal((last + 1, d)); u = rng.random(last + 1)
    m = O.Model("normal_fn", d)
    for first, step in [(1, 1), (101, 1), (101, 5), (7, 13)]:
        res = O.run_chain(m, O.sampler("RWM", scale=0.1), (first, step, last), np.ones(d), None, z, u)
        s, a = _py_rwm(lambda v: -sum(x * x for x in v), np.ones(d), 0.1, z, u, first, step, last)
        assert res["samples"].shape == s.shape == (len(range(first, last + 1, step)), d)
        assert np.array_equal(res["samples"], s) and np.array_equal(res["accept"].astype(bool), a)
        assert np.all(np.isnan(res["grads"]))                   # SerialMC.jl:42: no gradient from RWM -> NaN
    # model.scale multiplies the sampler scale (RWM.jl:52)
    res = O.run_chain(m, O.sampler("RWM", scale=0.1), (1, 1, 50), np.ones(d), np.array([2.0, 1.0, 0.5]), z, u)
    s, _ = _py_rwm(lambda v: -sum(x * x for x in v), np.ones(d), 0.1 * np.array([2.0, 1.0, 0.5]), z, u, 1, 1, 50)
    assert np.array_equal(res["samples"], s)
    # bad ranges (SerialMC.jl:25-27) and initial values out of support (RWM.jl:55)
    assert O.run_chain(m, O.sampler("RWM", scale=0.1), (0, 1, 10), np.ones(d), None, z, u)["rc"] == -2
    ou = O.Model("ou", 3, None, ou_series(20, 1), (100.0, 2.0, 20.0))
    assert O.run_chain(ou, O.sampler("RWM", scale=0.1), (1, 1, 10), np.array([-1.0, 1, 1]), None, z, u)["rc"] == -1


def test_hmc_leapfrog_restatement(O):
    """HMC.jl:93-102,136-158 restated in numpy for -dot(v,v): separate half steps, rand() < exp(H0 - H)."""
    rng = np.random.default_rng(9)
    d, last, eps, nl = 3, 200, 0.75, 10
    z = rng.standard_normal((last + 1, d)); u = rng.random(last + 1)
    res = O.run_chain(O.Model("normal_fn", d), O.sampler("HMC", scale=eps, nleaps=nl), (1, 1, last), np.ones(d), None, z, u)
    pars = np.ones(d); out = []
    for i in range(1, last + 1):
        m = z[i].copy(); p = pars.copy()
        H0 = (p[0] * p[0] + p[1] * p[1] + p[2] * p[2]) + 0.5 * (m[0] * m[0] + m[1] * m[1] + m[2] * m[2])
        for _ in range(nl):
            m = m + (0.5 * (-2 * p)) * eps
            p = p + eps * m
            m = m + (0.5 * (-2 * p)) * eps
        H = (p[0] * p[0] + p[1] * p[1] + p[2] * p[2]) + 0.5 * (m[0] * m[0] + m[1] * m[1] + m[2] * m[2])
        if u[i] < math.exp(min(H0 - H, 700)):
            pars = p
        out.append(pars.copy())
    assert np.allclose(res["samples"], np.array(out), rtol=1e-13, atol=1e-15)
    assert np.allclose(res["grads"], -2 * res["samples"])


def test_mala_ratio_restatement(O):
    rng = np.random.default_rng(10)
    d, last, h = 4, 150, 0.3
    z = rng.standard_normal((last + 1, d)); u = rng.random(last + 1)
    res = O.run_chain(O.Model("normal_dsl", d, hyper=(0.0, 1.0)), O.sampler("MALA", scale=h), (1, 1, last), np.ones(d), None, z, u)
    lp = lambda v: float(np.sum(-0.5 * v * v - 0.5 * math.log(2 * math.pi)))
    pars = np.ones(d); out = []
    for i in range(1, last + 1):
        mean = pars + (h / 2) * (-pars)
        prop = mean + math.sqrt(h) * z[i]
        q1 = np.sum(-(mean - prop) ** 2 / (2 * h) - math.log(2 * math.pi * h) / 2)
        mean2 = prop + (h / 2) * (-prop)
        q2 = np.sum(-(mean2 - pars) ** 2 / (2 * h) - math.log(2 * math.pi * h) / 2)
        ratio = lp(prop) + q2 - lp(pars) - q1
        if ratio > 0 or ratio > math.log(u[i]):
            pars = prop
        out.append(pars.copy())
    assert np.allclose(res["samples"], np.array(out), rtol=1e-12, atol=1e-14)


def test_hmcda_quirks(O):
    """HMCDA.jl:86-94,104,133-140: start step 1.0 (the search is a no-op), mu = log 10, adaptation for i < burnin,
    nLeaps = max(1, round(len/eps)), frozen dual step afterwards."""
    rng = np.random.default_rng(5)
    d, last = 3, 40
    z = rng.standard_normal((last + 1, d)); u = rng.random(last + 1)
    res = O.run_chain(O.Model("normal_fn", d), O.sampler("HMCDA", len=2.0), (1, 1, last), np.ones(d), None, z, u)
    # burnin = 0: `i < burnin` never holds, so the step stays at the initial dual step 1.0 (:140), nLeaps = round(2/1)
    assert np.all(res["eps"] == 1.0) and np.all(res["nleaps"] == 2)
    # burnin = 2: step 1 adapts once; its acceptance probability p1 restated here
    m = z[1].copy(); p = np.ones(d)
    H0 = (p @ p) + 0.5 * (m @ m)
    for _ in range(2):
        m = m + (0.5 * (-2 * p)) * 1.0; p = p + 1.0 * m; m = m + (0.5 * (-2 * p)) * 1.0
    p1 = min(1.0, math.exp(min(H0 - ((p @ p) + 0.5 * (m @ m)), 700)))
    dualH = (1 / 11) * (0.65 - p1)                                          # eta = 1/(1 + t0), dualH0 = 0
    eps2 = math.exp(math.log(10.0) - math.sqrt(1.0) * dualH / 0.05)         # mu = log(10 * 1.0)
    res2 = O.run_chain(O.Model("normal_fn", d), O.sampler("HMCDA", len=2.0), (3, 1, last), np.ones(d), None, z, u)
    assert abs(res2["eps"][0] - eps2) < 1e-12 * eps2                        # dual step after one update == eps2 (eta = 1)
    assert res2["nleaps"][0] == max(1, math.floor(2.0 / eps2 + 0.5))
    res0 = O.run_chain(O.Model("normal_fn", d), O.sampler("HMCDA", len=2.0), (21, 1, last), np.ones(d), None, z, u)
    assert np.all(res0["eps"] == res0["eps"][0])                            # frozen after burn-in
    # teacher forcing reproduces itself
    fe = np.concatenate([[np.nan], res["eps"]])
    res2 = O.run_chain(O.Model("normal_fn", d), O.sampler("HMCDA", len=2.0, force_eps=fe), (1, 1, last), np.ones(d), None, z, u)
    assert np.array_equal(res["samples"], res2["samples"]) and np.array_equal(res["eps"], res2["eps"])


def test_readme_statistical_band(O):
    """README.md:121-204: HMC(0.75) on -dot(v,v), 10 000 steps / 1000 burn-in: acceptance 79.76 %, ESS/n 0.593,
    IAT 1.687 -- an unseeded run, so a band, not a golden value."""
    rng = np.random.default_rng(2013)
    last = 10000
    z = rng.standard_normal((last + 1, 3)); u = rng.random(last + 1)
    res = O.run_chain(O.Model("normal_fn", 3), O.sampler("HMC", scale=0.75, nleaps=10), (1001, 1, last), np.ones(3), None, z, u)
    assert 78.0 < res["accept"].mean() * 100 < 82.0
    for j in range(3):
        x = res["samples"][:, j]
        assert 0.50 < O.ess(x) / len(x) < 0.68
        assert 1.45 < O.actime(x) < 2.0
        assert abs(x.mean()) < 0.03 and abs(x.var() - 0.5) < 0.03        # target is N(0, I/2)
        assert abs(O.mcvar(x, "imse") - 9.3e-5) < 2e-5 and abs(O.mcvar(x, "iid") - 5.5e-5) < 0.5e-5


@pytest.mark.parametrize("name", ["ar_pos", "ar_neg", "iid", "ar_strong"])
def test_stats_against_mpmath_golden(O, name):
    g = load_golden("golden_stats.npz")[name]
    x = g["x"]
    rel = lambda a, b: abs(a - b) <= 1e-11 * abs(b)
    assert rel(O.mean(x), g["mean"]) or abs(O.mean(x) - g["mean"]) < 1e-15
    assert rel(O.mcvar(x, "iid"), g["iid"])
    assert rel(O.mcvar(x, "bm", batchlen=g["bm_len"]), g["bm"])
    assert rel(O.mcvar(x, "imse"), g["imse"])
    assert rel(O.mcvar(x, "ipse"), g["ipse"])
    assert rel(O.mcvar(x, "imse", maxlag=21), g["imse_lag21"])
    assert rel(O.ess(x), len(x) * g["iid"] / g["imse"])
    assert rel(O.actime(x, "ipse"), g["ipse"] / g["iid"])


def test_golden_chains_regression(O):
    """the committed oracle trajectories still come out of the oracle (guards the fixtures the GPU tests use)"""
    for name, g in load_golden("golden_chains.npz").items():
        r = np.random.default_rng(g["draw_seed"])
        last = g["range"][2]
        z = r.standard_normal((last + 1, 3)); u = r.random(last + 1)
        res = O.run_chain(O.Model("normal_fn", 3), O.sampler(g["kind"], **g["kw"]), g["range"], np.ones(3), None, z, u)
        assert np.array_equal(res["samples"], g["samples"]) and np.array_equal(res["accept"], g["accept"]), name


# ---- second source for the headline path: the logistic AD chain + HMC / HMCDA steps, written independently ------------------
def _logistic_ad_numpy(X, y, b, psd=1.0, sgn=-1.0):
    """examples/logistic_regression.jl:16-20 through the reverse rules of src/dsl/definitions/MCMCDerivRules.jl, restated
    in numpy from the reference text alone (forward sweep, then one adjoint per statement, as ReverseDiffSource emits):
        vars ~ Normal(0, psd)            acc1 = sum logpdf(Normal(0, psd), vars);   dvars += (0 - vars)/(psd*psd)   (:57)
        t1 = X * vars                    dvars += X' * dt1
        t2 = sgn * t1 ; t3 = exp(t2)     dt2 = t3 .* dt3 ; dt1 = sgn * dt2
        t4 = 1. + t3 ; prob = 1 / t4     dt3 = dt4 ; dt4 = -dprob ./ (t4 .* t4)
        Y ~ Bernoulli(prob)              dprob = 1 ./ (prob - 1. + Y)                                             (:111)"""
    t1 = X @ b
    t3 = np.exp(sgn * t1)
    t4 = 1.0 + t3
    prob = 1.0 / t4
    ll = np.where(y == 1.0, np.log(prob), np.log(1.0 - prob))
    prior = -(0.5 * math.log(2 * math.pi) + 0.5 * (b / psd) ** 2 + math.log(psd))
    lt = float(prior.sum() + ll.sum())
    dprob = 1.0 / (prob - 1.0 + y)
    dt4 = -dprob / (t4 * t4)
    dt1 = sgn * (t3 * dt4)
    grad = X.T @ dt1 + (0.0 - b) / (psd * psd)
    return lt, grad


def _logistic_mpmath(X, y, b, psd=1.0, sgn=-1.0):
    """the same target from its mathematical definition at 50 digits (no AD chain): sum log sigma(+-eta) + log prior;
    gradient X'(y - p) - b/psd^2 with p = P(y = 1)"""
    import mpmath as mp
    mp.mp.dps = 50
    N, d = X.shape
    lt = mp.mpf(0)
    g = [mp.mpf(0)] * d
    for j in range(d):
        lt += -(mp.log(2 * mp.pi) / 2 + (mp.mpf(b[j]) / psd) ** 2 / 2 + mp.log(psd))
        g[j] = -mp.mpf(b[j]) / (mp.mpf(psd) ** 2)
    for i in range(N):
        eta = mp.fsum(mp.mpf(X[i, j]) * mp.mpf(b[j]) for j in range(d))
        p1 = 1 / (1 + mp.exp(sgn * eta))                  # prob: P(Y = 1)
        lt += mp.log(p1) if y[i] == 1.0 else mp.log(1 - p1)
        # d/d eta of log p1 = -sgn (1 - p1);  of log(1 - p1) = sgn p1
        r = -sgn * (1 - p1) if y[i] == 1.0 else sgn * p1
        for j in range(d):
            g[j] += r * mp.mpf(X[i, j])
    return float(lt), np.array([float(v) for v in g])


@pytest.mark.parametrize("sgn", [-1.0, 1.0])
def test_logistic_ad_chain_two_sources(O, sgn):
    """oracle (C, written from the same reference lines) == numpy restatement of the AD chain (1e-13) == the mathematical
    definition in 50-digit arithmetic (1e-12): three implementations by construction independent in their arithmetic."""
    N, d = 300, 7
    X, y, hy, b0 = make_regression("logistic", N, d, 21)
    if sgn > 0:
        y = 1.0 - y
    m = O.Model("logistic", d, X, y, (1.0, sgn))
    rng = np.random.default_rng(3)
    gs = np.abs(X).sum(0)
    for k in range(6):
        b = b0 + (0.3 * k) * rng.standard_normal(d)
        olt, og = m.evalallg(b)
        nlt, ng = _logistic_ad_numpy(X, y, b, 1.0, sgn)
        mlt, mg = _logistic_mpmath(X, y, b, 1.0, sgn)
        assert abs(olt - nlt) <= 1e-13 * abs(nlt) and np.all(np.abs(og - ng) <= 1e-13 * gs)
        assert abs(olt - mlt) <= 1e-12 * abs(mlt) and np.all(np.abs(og - mg) <= 1e-12 * gs)


def test_hmc_and_hmcda_steps_on_logistic_two_sources(O):
    """HMC.jl:93-102,136-158 and HMCDA.jl:97-142 (frozen and adapting, continued from a restored dual-averaging state)
    restated in numpy on top of the numpy AD chain, against the oracle's run_chain on a logistic posterior at a
    realistic step size (acceptance strictly inside (0, 1))."""
    N, d, last = 400, 8, 40
    X, y, hy, b0 = make_regression("logistic", N, d, 22)
    m = O.Model("logistic", d, X, y, hy)
    rng = np.random.default_rng(4)
    z = rng.standard_normal((last + 1, d)); u = rng.random(last + 1)
    f = lambda b: _logistic_ad_numpy(X, y, b)

    def leapfrogs(p, mom, g, eps, n):
        lt = None
        for _ in range(n):
            mom = mom + (0.5 * g) * eps          # HMC.jl:95
            p = p + eps * mom                    # :96
            lt, g = f(p)                         # :97
            mom = mom + (0.5 * g) * eps          # :98
        return p, mom, g, lt

    # --- plain HMC ---
    eps, nl = 0.09, 6
    res = O.run_chain(m, O.sampler("HMC", scale=eps, nleaps=nl), (1, 1, last), b0, None, z, u)
    pars = b0.copy(); lt, g = f(pars); out, acc = [], []
    for i in range(1, last + 1):
        mom = z[i].copy()
        H0 = -lt + 0.5 * float(mom @ mom)
        p, mom, gp, ltp = leapfrogs(pars.copy(), mom, g.copy(), eps, nl)
        H = -ltp + 0.5 * float(mom @ mom)
        a = u[i] < math.exp(min(H0 - H, 700.0))  # :154
        if a:
            pars, lt, g = p, ltp, gp
        out.append(pars.copy()); acc.append(a)
    assert 0.3 < np.mean(acc) < 1.0
    assert np.array_equal(res["accept"], np.array(acc, dtype=np.uint8))
    assert np.allclose(res["samples"], np.array(out), rtol=1e-10, atol=1e-13)
    # --- HMCDA from a restored state (step0 = 60): 12 adapting steps (i < burnin = 72), then frozen ---
    s0, B, last2 = 60, 72, 95
    z2 = rng.standard_normal((last2 + 1, d)); u2 = rng.random(last2 + 1)
    L, eps0, rate, shrink, t0, kappa = 0.6, 0.08, 0.65, 0.05, 10.0, 0.75
    dualH0 = (math.log(10.0) - math.log(eps0)) * shrink / math.sqrt(s0)
    res = O.run_chain(m, O.sampler("HMCDA", len=L, start_step=s0, da_state=[eps0, eps0, dualH0]), (B + 1, 1, last2), b0, None, z2, u2)
    pars = b0.copy(); lt, g = f(pars)
    ls, dual, dualH, mu = eps0, eps0, dualH0, math.log(10.0)
    keep_eps, keep_nl, keep_acc, keep_s = [], [], [], []
    for i in range(s0 + 1, last2 + 1):
        mom = z2[i].copy()
        H0 = -lt + 0.5 * float(mom @ mom)
        n = max(1, int(math.floor(L / ls + 0.5)))                       # round(len/leapStep), HMCDA.jl:104
        p, mom, gp, ltp = leapfrogs(pars.copy(), mom, g.copy(), ls, n)
        pacc = min(1.0, math.exp(min(H0 - (-ltp + 0.5 * float(mom @ mom)), 700.0)))   # :120
        a = u2[i] < pacc                                                # :121
        if a:
            pars, lt, g = p, ltp, gp
        if i > B:
            keep_eps.append(ls); keep_nl.append(n); keep_acc.append(a); keep_s.append(pars.copy())
        if i < B:                                                       # :133-138
            eta = 1.0 / (i + t0)
            dualH = (1 - eta) * dualH + eta * (rate - pacc)
            ls = math.exp(mu - math.sqrt(i) * dualH / shrink)
            eta = i ** (-kappa)
            dual = math.exp((1 - eta) * math.log(dual) + eta * math.log(ls))
        else:
            ls = dual                                                   # :140
    assert 0.3 < np.mean(keep_acc) < 1.0 and abs(keep_eps[0] / eps0 - 1) > 1e-3
    assert np.array_equal(res["accept"], np.array(keep_acc, dtype=np.uint8)) and np.array_equal(res["nleaps"], keep_nl)
    assert np.allclose(res["eps"], keep_eps, rtol=1e-9) and np.allclose(res["samples"], np.array(keep_s), rtol=1e-9, atol=1e-12)


def test_fast_baseline_variant_agrees(O):
    """the row-blocked evaluation timed by bench.py's CPU legs is the same function (reassociated sums: 1e-13)"""
    X, y, hy, b0 = make_regression("logistic", 1300, 9, 33)
    m = O.Model("logistic", 9, X, y, hy)
    rng = np.random.default_rng(1)
    gs = np.abs(X).sum(0)
    try:
        for k in range(4):
            b = b0 + 0.2 * k * rng.standard_normal(9)
            O.set_fast_baseline(False); lt0, g0 = m.evalallg(b); e0 = m.eval(b)
            O.set_fast_baseline(True); lt1, g1 = m.evalallg(b); e1 = m.eval(b)
            assert abs(lt1 - lt0) <= 1e-13 * abs(lt0) and abs(e1 - e0) <= 1e-13 * abs(e0) and np.all(np.abs(g1 - g0) <= 1e-13 * gs)
        O.set_fast_baseline(True)
        assert m.evalallg(np.array([0.0, 800.0] + [0.0] * 7))[0] == -np.inf          # LLAcc semantics kept
    finally:
        O.set_fast_baseline(False)
