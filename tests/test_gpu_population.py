"""GPU suite, population runners: SeqMC and SerialTempMC on the device against the CPU oracle
with injected draws (bit-exact: closed-form targets, identical operation order), plus the README's SeqMC example."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ladder(O, capi, fam, d, hypers, kinds):
    models = [O.Model(fam, d, hyper=h) for h in hypers]
    osmp = [O.sampler(k, **kw) for k, kw in kinds]
    gsmp = [capi.sampler_cfg(k, **kw) for k, kw in kinds]
    return models, osmp, gsmp


@pytest.mark.parametrize("fam,d,kinds", [
    ("abs_normal", 1, [("RWM", dict(scale=float(s))) for s in np.logspace(1, -1, 10)]),                 # README.md:248-263
    ("normal_dsl", 3, [("MALA", dict(scale=0.3)), ("HMC", dict(scale=0.4, nleaps=3)), ("RWM", dict(scale=0.5))]),
    ("normal_fn", 2, [("HMC", dict(scale=0.3, nleaps=2)), ("RWM", dict(scale=0.2))]),
])
@pytest.mark.parametrize("trigger", [1e-10, 0.5, 1e300])
def test_seqmc_matches_oracle(O, capi, ctx, fam, d, kinds, trigger):
    nt = len(kinds)
    sig = np.logspace(1, -1, nt) if fam != "normal_fn" else np.ones(nt)
    hypers = [(1.0, float(s)) for s in sig] if fam == "abs_normal" else ([(0.0, float(s)) for s in sig] if fam == "normal_dsl" else [()] * nt)
    models, osmp, gsmp = _ladder(O, capi, fam, d, hypers, kinds)
    rng = np.random.default_rng(nt + d)
    npart, steps, burnin = 257, 6, 2
    parts = rng.standard_normal((npart, d))
    zn = rng.standard_normal((steps, nt, npart, d)); un = rng.random((steps, nt, npart)); ru = rng.random((steps, nt, npart))
    ref = O.run_seqmc(models, osmp, steps, burnin, trigger, parts, zn, un, ru)
    out = ctx.run_seqmc(fam, d, hypers, gsmp, steps, burnin, trigger, parts, normals=zn, uniforms=un, res_uniforms=ru)
    assert ref["rc"] == 0 and out["n_resamples"] == ref["n_resamples"]
    if trigger == 1e300:
        assert out["n_resamples"] == steps * nt           # resampling after every target
    assert np.array_equal(out["samples"], ref["samples"])
    assert np.allclose(out["weights"], ref["weights"], rtol=1e-14, atol=0)      # exp(): CUDA vs glibc, last bit


def test_seqmc_readme_example(O):
    """README.md:245-272 through the host API: 10 tempered targets |x| ~ N(1, s_i), RWM(s_i), 1000 particles, 10 steps"""
    import mcmc_jl_b200 as mj
    nmod = 10
    sts = np.logspace(1, -1, nmod)
    mods = [mj.model(f"y = abs(x)\n y ~ Normal(1, {s})", x=0.0, gradient=False) for s in sts]
    targets = [mods[i] * mj.RWM(float(sts[i])) * mj.SeqMC(steps=10, burnin=0) for i in range(nmod)]
    particles = [[v] for v in np.random.default_rng(1).standard_normal(1000)]
    chain = mj.run(targets, particles=particles, seed=3)
    assert chain.samples.shape == (10000, 1) and list(chain.samples.columns) == ["x"]
    w = chain.diagnostics["weigths"]; x = chain.samples["x"].values
    assert w.shape == (10000,) and np.all(np.isfinite(w)) and chain.diagnostics["particle"][1000] == 1
    last = slice(9000, 10000)
    m_abs = np.sum(np.abs(x[last]) * w[last]) / np.sum(w[last])
    assert abs(m_abs - 1.0) < 0.05                        # the final target is |x| ~ N(1, 0.1)
    assert 0.25 < np.mean(x[last] > 0) < 0.75             # both modes +-1 are populated
    with pytest.raises(AssertionError):
        mj.SeqMC(steps=3, burnin=3)


@pytest.mark.parametrize("fam,d,kinds", [
    ("normal_dsl", 2, [("RWM", dict(scale=0.5)), ("RWM", dict(scale=1.0)), ("RWM", dict(scale=2.0))]),
    ("normal_dsl", 3, [("HMC", dict(scale=0.4, nleaps=3)), ("MALA", dict(scale=0.6)), ("RWM", dict(scale=1.5)), ("HMC", dict(scale=0.9, nleaps=2))]),
    ("abs_normal", 1, [("RWM", dict(scale=2.0)), ("RWM", dict(scale=0.3))]),
])
def test_serialtemp_matches_oracle(O, capi, ctx, fam, d, kinds):
    nt = len(kinds)
    sig = [1.0, 2.0, 4.0, 8.0][:nt]
    hypers = [(1.0 if fam == "abs_normal" else 0.0, s) for s in sig]
    models, osmp, gsmp = _ladder(O, capi, fam, d, hypers, kinds)
    rng = np.random.default_rng(7 + nt)
    nrep, steps, burnin, swap = 70, 300, 50, 5
    inits = rng.standard_normal((nt, d))
    zn = rng.standard_normal((nrep, steps + 2, d)); un = rng.random((nrep, steps + 2))
    pk = rng.random((nrep, steps + 1)); sw = rng.random((nrep, steps + 1))
    out = ctx.run_serialtemp(fam, d, hypers, gsmp, steps, burnin, swap, nrep, inits, normals=zn, uniforms=un, pick=pk, swap=sw)
    visited = set()
    for c in range(nrep):
        ref = O.run_serialtemp(models, osmp, steps, burnin, swap, inits, zn[c], un[c], pk[c], sw[c])
        assert ref["rc"] == 0
        assert np.array_equal(out["at"][c], ref["at"]), c
        assert np.array_equal(out["samples"][c], ref["samples"]), c
        visited |= set(ref["at"].tolist())
    assert visited == set(range(nt))                      # the replicas do move between tasks


def test_serialtemp_host_api():
    import mcmc_jl_b200 as mj
    mods = [mj.model("v ~ Normal(0, %g)" % s, v=np.ones(2), gradient=True) for s in (1.0, 3.0, 9.0)]
    tasks = [mods[i] * mj.HMC(3, 0.3 * s) * mj.SerialTempMC(steps=4000, burnin=500, swapPeriod=5) for i, s in enumerate((1.0, 3.0, 9.0))]
    ch = mj.run(tasks, seed=5)
    assert ch.samples.shape == (3500, 2) and set(np.unique(ch.diagnostics["task"])) <= {1, 2, 3}
    assert np.isfinite(ch.samples.values).all() and len(np.unique(ch.diagnostics["task"])) >= 2   # the chain changes task
    # (no per-task moment check: with logW fixed at 0 (SerialTempMC.jl:48,70) and swaps made from the PRE-step state
    #  (:53,62), the reference's chain is not a per-task sampler of N(0, sigma_t); parity is what is checked above)
    reps = mj.run(tasks, nreplicas=8, seed=6)
    assert len(reps) == 8 and not np.array_equal(reps[0].samples.values, reps[1].samples.values)


def test_seqmc_models_equals_closed_form_runner(O, capi, ctx):
    """the model-array SeqMC (wave engine) on a closed-form ladder reproduces the in-register runner and the oracle"""
    fam, d = "normal_dsl", 3
    kinds = [("MALA", dict(scale=0.3)), ("HMC", dict(scale=0.4, nleaps=3)), ("RWM", dict(scale=0.5))]
    hypers = [(0.0, float(s)) for s in np.logspace(1, -1, 3)]
    models, osmp, gsmp = _ladder(O, capi, fam, d, hypers, kinds)
    rng = np.random.default_rng(5)
    npart, steps, burnin, trigger = 130, 5, 1, 0.5
    parts = rng.standard_normal((npart, d))
    zn = rng.standard_normal((steps, 3, npart, d)); un = rng.random((steps, 3, npart)); ru = rng.random((steps, 3, npart))
    ref = O.run_seqmc(models, osmp, steps, burnin, trigger, parts, zn, un, ru)
    dms = [capi.DeviceModel(ctx, fam, d, hyper=h) for h in hypers]
    out = ctx.run_seqmc_models(dms, gsmp, steps, burnin, trigger, parts, normals=zn, uniforms=un, res_uniforms=ru)
    assert out["n_resamples"] == ref["n_resamples"] and np.array_equal(out["samples"], ref["samples"])
    assert np.allclose(out["weights"], ref["weights"], rtol=1e-14, atol=0)
    # Philox mode: both runners key the draws by (seed, particle, (iter-1) nt + t + 1): identical populations
    a = ctx.run_seqmc(fam, d, hypers, gsmp, steps, burnin, trigger, parts, seed=9)
    b = ctx.run_seqmc_models(dms, gsmp, steps, burnin, trigger, parts, seed=9)
    assert a["n_resamples"] == b["n_resamples"] and np.array_equal(a["samples"], b["samples"])
    for m in dms:
        m.close()


@pytest.mark.parametrize("fam", ["logistic", "probit", "linear"])
def test_seqmc_over_regression_models(O, capi, ctx, fam):
    """SeqMC with the regression families (the population reuses K1): a ladder of priors from wide to the
    model's own over the same data, RWM / MALA / HMC tasks, particles = chains of the likelihood kernel; against the oracle
    with injected draws: same resampling events, samples to 1e-9, weights to 1e-8."""
    from conftest import make_regression
    N, d, npart, steps, burnin = 400, 12, 200, 4, 1
    X, y, hy, b0 = make_regression(fam, N, d, 31)
    sds = [4.0, 2.0, 1.0] if fam != "probit" else [40.0, 20.0, 10.0]
    hys = [(sd,) + tuple(hy[1:]) for sd in sds]
    kinds = [("RWM", dict(scale=0.02)), ("MALA", dict(scale=0.002)), ("HMC", dict(scale=0.03, nleaps=3))]
    oms = [O.Model(fam, d, X, y, h) for h in hys]
    dms = [capi.DeviceModel(ctx, fam, d, X, y, h) for h in hys]
    osmp = [O.sampler(k, **kw) for k, kw in kinds]; gsmp = [capi.sampler_cfg(k, **kw) for k, kw in kinds]
    rng = np.random.default_rng(3)
    parts = b0 + 0.1 * rng.standard_normal((npart, d))
    zn = rng.standard_normal((steps, 3, npart, d)); un = rng.random((steps, 3, npart)); ru = rng.random((steps, 3, npart))
    # a trigger between the weight variances seen, so that some targets resample and some do not
    probe = O.run_seqmc(oms, osmp, steps, burnin, 0.0, parts, zn, un, ru)
    trig = float(np.median(np.var(probe["weights"].reshape(steps - burnin, npart), axis=1, ddof=1)))
    ref = O.run_seqmc(oms, osmp, steps, burnin, trig, parts, zn, un, ru)
    out = ctx.run_seqmc_models(dms, gsmp, steps, burnin, trig, parts, normals=zn, uniforms=un, res_uniforms=ru)
    assert ref["rc"] == 0 and out["n_resamples"] == ref["n_resamples"]
    assert np.allclose(out["samples"], ref["samples"], rtol=1e-9, atol=1e-12)
    assert np.allclose(out["weights"], ref["weights"], rtol=1e-7, atol=0)
    assert out["info"]["n_grad_evals"] == npart * steps * (2 + 2 + 4)
    for m in dms:
        m.close()


def test_serialtemp_models_equals_closed_form_runner(O, capi, ctx):
    """the model-array SerialTempMC (replicas regrouped by task, one-step runs of the wave engine) reproduces the
    one-launch closed-form runner and the oracle, with injected draws and with the engine's own Philox draws"""
    fam, d = "normal_dsl", 3
    kinds = [("HMC", dict(scale=0.4, nleaps=3)), ("MALA", dict(scale=0.6)), ("RWM", dict(scale=1.5))]
    hypers = [(0.0, 1.0), (0.0, 2.0), (0.0, 4.0)]
    models, osmp, gsmp = _ladder(O, capi, fam, d, hypers, kinds)
    rng = np.random.default_rng(11)
    nrep, steps, burnin, swap = 70, 60, 10, 3
    inits = rng.standard_normal((3, d))
    zn = rng.standard_normal((nrep, steps + 2, d)); un = rng.random((nrep, steps + 2))
    pk = rng.random((nrep, steps + 1)); sw = rng.random((nrep, steps + 1))
    dms = [capi.DeviceModel(ctx, fam, d, hyper=h) for h in hypers]
    out = ctx.run_serialtemp_models(dms, gsmp, steps, burnin, swap, nrep, inits, normals=zn, uniforms=un, pick=pk, swap=sw)
    visited = set()
    for c in range(nrep):
        ref = O.run_serialtemp(models, osmp, steps, burnin, swap, inits, zn[c], un[c], pk[c], sw[c])
        assert np.array_equal(out["at"][c], ref["at"]) and np.array_equal(out["samples"][c], ref["samples"]), c
        visited |= set(ref["at"].tolist())
    assert visited == {0, 1, 2}
    a = ctx.run_serialtemp(fam, d, hypers, gsmp, steps, burnin, swap, nrep, inits, seed=4)
    b = ctx.run_serialtemp_models(dms, gsmp, steps, burnin, swap, nrep, inits, seed=4)
    assert np.array_equal(a["at"], b["at"]) and np.array_equal(a["samples"], b["samples"])
    # replicas keyed by global id: two shards reproduce the unsharded run
    lo = ctx.run_serialtemp_models(dms, gsmp, steps, burnin, swap, 30, inits, seed=4)
    hi = ctx.run_serialtemp_models(dms, gsmp, steps, burnin, swap, 40, inits, seed=4, rep_offset=30)
    assert np.array_equal(b["samples"][:30], lo["samples"]) and np.array_equal(b["samples"][30:], hi["samples"])
    for m in dms:
        m.close()


@pytest.mark.parametrize("fam", ["logistic", "probit"])
def test_serialtemp_over_regression_models(O, capi, ctx, fam):
    """SerialTempMC with regression models: a ladder of priors over the same data; every replica's visited tasks identical
    to the oracle's, samples to 1e-9 (the swap test compares log-targets of two models: decided by K1's sums)."""
    from conftest import make_regression
    N, d, nrep, steps, burnin, swap = 300, 8, 40, 40, 5, 4
    X, y, hy, b0 = make_regression(fam, N, d, 41)
    sds = [1.0, 1.5, 2.5] if fam != "probit" else [10.0, 15.0, 25.0]
    hys = [(sd,) + tuple(hy[1:]) for sd in sds]
    kinds = [("HMC", dict(scale=0.05, nleaps=3)), ("MALA", dict(scale=0.004)), ("RWM", dict(scale=0.03))]
    oms = [O.Model(fam, d, X, y, h) for h in hys]
    dms = [capi.DeviceModel(ctx, fam, d, X, y, h) for h in hys]
    osmp = [O.sampler(k, **kw) for k, kw in kinds]; gsmp = [capi.sampler_cfg(k, **kw) for k, kw in kinds]
    rng = np.random.default_rng(2)
    inits = b0 + 0.05 * rng.standard_normal((3, d))
    zn = rng.standard_normal((nrep, steps + 2, d)); un = rng.random((nrep, steps + 2))
    pk = rng.random((nrep, steps + 1)); sw = rng.random((nrep, steps + 1))
    out = ctx.run_serialtemp_models(dms, gsmp, steps, burnin, swap, nrep, inits, normals=zn, uniforms=un, pick=pk, swap=sw)
    visited = set()
    for c in range(nrep):
        ref = O.run_serialtemp(oms, osmp, steps, burnin, swap, inits, zn[c], un[c], pk[c], sw[c])
        assert ref["rc"] == 0 and np.array_equal(out["at"][c], ref["at"]), c
        assert np.allclose(out["samples"][c], ref["samples"], rtol=1e-9, atol=1e-12), c
        visited |= set(ref["at"].tolist())
    assert len(visited) >= 2
    for m in dms:
        m.close()


def test_host_api_population_and_streaming(O):
    """the Python mirror drives the new paths: SeqMC / SerialTempMC over regression models (dispatch on family / size),
    GPUMC(store_draws=False)"""
    import mcmc_jl_b200 as mj
    from conftest import make_regression
    X, y, hy, b0 = make_regression("logistic", 300, 6, 51)
    mods = [mj.model("logistic", X=X, Y=y, vars=b0, prior_sd=sd) for sd in (4.0, 2.0, 1.0)]
    targets = [mods[i] * mj.RWM(0.05) * mj.SeqMC(steps=4, burnin=1) for i in range(3)]
    parts = b0 + 0.1 * np.random.default_rng(0).standard_normal((150, 6))
    chain = mj.run(targets, particles=list(parts), seed=2)
    assert chain.samples.shape == (3 * 150, 6) and np.all(np.isfinite(chain.diagnostics["weigths"])) and chain.info["n_grad_evals"] == 150 * 4 * 3 * 2
    temps = [mods[i] * mj.HMC(3, 0.05) * mj.SerialTempMC(steps=30, burnin=5, swapPeriod=3) for i in range(3)]
    chains = mj.run(temps, nreplicas=8, seed=3)
    assert len(chains) == 8 and chains[0].samples.shape == (25, 6) and set(np.unique(np.concatenate([c.diagnostics["task"] for c in chains]))) <= {1, 2, 3}
    m = mj.model("normal", init=np.ones(3))
    full = mj.run(m * mj.HMC(0.75) * mj.GPUMC(steps=1200, burnin=200, nchains=256, seed=5))
    st = mj.run(m * mj.HMC(0.75) * mj.GPUMC(steps=1200, burnin=200, nchains=256, seed=5, store_draws=False))
    assert np.array_equal(mj.mean(st), mj.mean(full)) and np.array_equal(mj.acceptance(st), mj.acceptance(full))
    assert np.allclose(mj.var(st, vtype="bm"), mj.var(full, vtype="bm"), rtol=1e-10) and np.allclose(mj.ess(st, vtype="bm"), mj.ess(full, vtype="bm"), rtol=1e-10)
    with pytest.raises(mj.MCMCGPUError):
        st.arrays()
    full.close(); st.close()
