"""CPU suite, part 2: the C-ABI library loads and exports every symbol of include/mcmcgpu.h, the host
mirror reproduces the reference's constructor semantics, and nothing computes without a GPU."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def test_library_exports_every_declared_symbol(capi):
    hdr = open(os.path.join(ROOT, "include", "mcmcgpu.h")).read()
    declared = sorted(set(re.findall(r"\b(mcmcgpu_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 20
    lib = capi.lib()
    for name in declared:
        assert hasattr(lib, name), f"libmcmcgpu.so does not export {name}"
    assert sorted(capi.EXPORTS) == declared, "ctypes binding and header disagree"
    assert lib.mcmcgpu_abi_version() == 3


def test_no_cpu_fallback(capi):
    """without a device every entry point that computes fails with MCMCGPU_E_CUDA"""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present; the no-device error path is checked on the CPU box")
    with pytest.raises(capi.MCMCGPUError) as e:
        capi.Context(0)
    assert e.value.code == capi.E_CUDA and "no CPU fallback" in str(e.value)
    import mcmc_jl_b200 as mj
    with pytest.raises(capi.MCMCGPUError):
        mj.model("normal", init=np.ones(3))      # model() checks isfinite(eval(init)) on the device (likmodel.jl:54)


def test_product_does_not_import_oracle():
    """the oracle is test infrastructure: no file of the product package may mention it"""
    pkg = os.path.join(ROOT, "mcmc.jl_b200")
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".jl")):
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                assert "mcmc_oracle" not in txt and "import oracle" not in txt and "orc_" not in txt, fn


def test_sampler_constructors():
    import mcmc_jl_b200 as mj
    assert (mj.HMC().nLeaps, mj.HMC().leapStep) == (10, 0.1)                 # HMC.jl:64
    assert (mj.HMC(0.75).nLeaps, mj.HMC(0.75).leapStep) == (10, 0.75)        # HMC.jl:69
    assert (mj.HMC(2, 0.1).nLeaps, mj.HMC(2, 0.1).leapStep) == (2, 0.1)      # HMC.jl:67
    assert mj.HMC(3).nLeaps == 3 and mj.HMC(3).leapStep == 0.1               # HMC.jl:65
    t = mj.EmpMCTuner(0.8, verbose=True)
    assert (t.adaptStep, t.maxStep, t.targetPath, t.targetRate) == (100, 200, 1.0, 0.8)   # samplers.jl:49-50
    assert mj.HMC(5, 0.2, t).tuner is t and mj.MALA(t).tuner is t
    h = mj.HMCDA()
    assert (h.rate, h.len, h.shrinkage, h.t0, h.step) == (0.65, 2.0, 0.05, 10.0, 0.75)    # HMCDA.jl:42-43
    assert mj.RWM().scale == 1.0 and mj.RWM(0.1).scale == 0.1 and mj.MALA().driftStep == 1.0
    for bad in (lambda: mj.RWM(-1.0), lambda: mj.MALA(0.0), lambda: mj.HMC(0, 0.1), lambda: mj.HMC(2, -0.1),
                lambda: mj.HMCDA(rate=1.5), lambda: mj.HMCDA(len=-1.0), lambda: mj.HMCDA(shrinkage=0.0),
                lambda: mj.EmpMCTuner(1.2)):
        with pytest.raises(AssertionError):
            bad()
    cfg = mj.HMC(7, 0.3, t)._cfg()
    assert (cfg.kind, cfg.nleaps, cfg.scale, cfg.tuner_on, cfg.adapt_step, cfg.target_rate) == (2, 7, 0.3, 1, 100, 0.8)


def test_serialmc_ranges():
    import mcmc_jl_b200 as mj
    r = mj.SerialMC(steps=1000, burnin=100)                                  # README.md:84
    assert (r.burnin, r.thinning, r.len, len(r.r)) == (100, 1, 1000, 900)
    r = mj.SerialMC(steps=1000, burnin=100, thinning=5)                      # README.md:87 "180 post-burnin iterations"
    assert len(r.r) == 180 and list(r.r)[:2] == [101, 106]
    assert list(mj.SerialMC(range(101, 1001, 5)).r) == list(r.r)             # README.md:90  101:5:1000
    r = mj.SerialMC(100, 1000)                                               # SerialMC(100:1000)
    assert (r.burnin, r.thinning, r.len) == (99, 1, 1000)
    r = mj.SerialMC(10000, 10, 100000)                                       # linear_regression.jl:22
    assert (r.burnin, r.thinning, r.len) == (9999, 10, 100000)
    assert (mj.SerialMC().burnin, mj.SerialMC().len) == (0, 100)             # SerialMC.jl:35 defaults
    with pytest.raises(AssertionError):
        mj.SerialMC(steps=10, burnin=10)                                     # SerialMC.jl:26
    with pytest.raises(AssertionError):
        mj.SerialMC(0, 10)                                                   # SerialMC.jl:25
    g = mj.GPUMC(steps=400, burnin=200, nchains=10000, seed=4)
    assert (g.burnin, g.len, g.nchains, g.seed, g.shard) == (200, 400, 10000, 4, "chains")


def test_task_operator_broadcasting():
    import mcmc_jl_b200 as mj
    from mcmc_jl_b200 import api
    m = object.__new__(mj.MCMCLikelihoodModel)                               # no device needed for the operator
    m.size, m.has_gradient = 3, True
    t = m * mj.RWM(0.1) * mj.SerialMC(steps=10)
    assert isinstance(t, mj.MCMCTask) and t.model is m and isinstance(t.sampler, mj.RWM)
    ts = m * [mj.RWM(0.1), mj.MALA(0.1), mj.HMC(3, 0.1)] * mj.SerialMC(steps=1000)     # test/test_syntax.jl:79
    assert [type(x.sampler).__name__ for x in ts] == ["RWM", "MALA", "HMC"] and all(x.model is m for x in ts)
    ts = [m, m] * mj.HMC(0.75) * mj.SerialMC(steps=5) if False else api._combine(api._combine([m, m], mj.HMC(0.75)), mj.SerialMC(steps=5))
    assert len(ts) == 2
    with pytest.raises(ValueError):
        api._combine(api._combine([m, m], [mj.RWM(), mj.RWM(), mj.RWM()]), mj.SerialMC(steps=5))
    # column names from pmap (SerialMC.jl:70-79)
    m.pmap = {"pars": (1, (3,))}
    assert api._colnames(m) == ["pars.1", "pars.2", "pars.3"]
    m.pmap = {"tau": (1, ()), "sigma": (2, ()), "mu": (3, ())}
    assert api._colnames(m) == ["tau", "sigma", "mu"]
    assert api._ispartition({"a": (1, (2,)), "b": (3, ())}, 3) and not api._ispartition({"a": (1, (2,)), "b": (2, ())}, 3)


def test_rank_slices_cover_all_chains(monkeypatch):
    from mcmc_jl_b200 import api
    for world in (1, 2, 4, 8):
        seen = []
        for rank in range(world):
            monkeypatch.setenv("WORLD_SIZE", str(world)); monkeypatch.setenv("RANK", str(rank))
            lo, n = api._rank_slice(10000)
            seen += list(range(lo, lo + n))
        assert seen == list(range(10000))


def test_dsl_recogniser():
    """the example models of the reference, written in its DSL, map onto the GPU families"""
    from mcmc_jl_b200.dsl import recognise
    X, Y = np.ones((5, 3)), np.arange(5.0)
    r = recognise("v ~ Normal(0, 1)", dict(v=np.ones(3)))                          # README.md:67-72
    assert (r["family"], r["hyper"], r["pmap"]) == ("normal_dsl", (0.0, 1.0), {"v": (1, (3,))})
    lin = """
        vars ~ Normal(0, 1.0)  # Normal prior, std 1.0 for predictors
        resid = Y - X * vars
        resid ~ Normal(0, 1.0)
    """                                                                            # examples/linear_regression.jl:14-18
    r = recognise(lin, dict(vars=np.zeros(3), X=X, Y=Y))
    assert r["family"] == "linear" and r["hyper"] == (1.0, 1.0) and r["X"] is not None and list(r["pmap"]) == ["vars"]
    log_ = "vars ~ Normal(0, 1.0)\n prob = 1 / (1. + exp(- X * vars))\n Y ~ Bernoulli(prob)"   # examples/logistic_regression.jl:16-20
    assert recognise(log_, dict(vars=np.zeros(3), X=X, Y=Y))["hyper"] == (1.0, -1.0)
    log2 = "vars ~ Normal(0, 1.0); prob = 1 / (1. + exp(X * vars)); Y ~ Bernoulli(prob)"         # test/test_syntax.jl:16-20
    assert recognise(log2, dict(vars=np.zeros(3), X=X, Y=Y))["hyper"] == (1.0, 1.0)
    ou = """
        tau ~ Uniform(0, 100)
        sigma ~ Uniform(0, 2)
        mu ~ Uniform(0, 20)
        fac = exp(- 1. / tau)
        resid = x[2:end] - x[1:end-1] * fac - mu * (1. - fac)
        resid ~ Normal(0, sigma)
    """                                                                            # examples/ornstein.jl:19-27
    r = recognise(ou, dict(tau=0.05, sigma=1.0, mu=1.0, x=np.arange(10.0)))
    assert r["family"] == "ou" and r["hyper"] == (100.0, 2.0, 20.0) and list(r["init"]) == [0.05, 1.0, 1.0]
    assert list(r["pmap"]) == ["tau", "sigma", "mu"]
    with pytest.raises(NotImplementedError):
        recognise("y = abs(x)\n y ~ Gamma(2, 3)", dict(x=0.0))
    with pytest.raises(ValueError):
        recognise(lin, dict(vars=np.zeros(3), X=X))                                # Y missing


def _build_c_demo(tmp_path):
    import subprocess
    exe = os.path.join(str(tmp_path), "c_abi_demo")
    pkg = os.path.join(ROOT, "mcmc.jl_b200")
    subprocess.check_call(["gcc", "-O2", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_abi_demo.c"),
                           "-L" + pkg, "-lmcmcgpu", "-Wl,-rpath," + pkg, "-lm", "-o", exe])
    return exe


def test_struct_layouts_match_the_header(capi, tmp_path):
    """the ctypes Structures of the binding have exactly the C layout of include/mcmcgpu.h (sizes and field offsets)"""
    import subprocess
    out = subprocess.check_output([_build_c_demo(tmp_path), "layout"], text=True).strip().splitlines()
    for line, st in zip(out, (capi.SamplerCfg, capi.RunnerCfg, capi.RunInfo)):
        toks = line.split()
        import ctypes
        assert int(toks[1]) == ctypes.sizeof(st), line
        fields = dict(zip(toks[2::2], map(int, toks[3::2])))
        assert set(fields) == {n for n, _ in st._fields_}, line
        for name, off in fields.items():
            assert getattr(st, name).offset == off, (st.__name__, name)


def test_generated_link_tables_are_reproducible():
    """csrc/exp_table.h and csrc/log_table.h hold exactly what tools/gen_exp_table.py / gen_log_table.py compute (mpmath),
    and the replayed device arithmetic built on them is accurate to a few 1e-16 (the same replay the generators print)."""
    import importlib.util
    import re
    import mpmath as mp
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    def load(name):
        spec = importlib.util.spec_from_file_location(name, os.path.join(root, "tools", name + ".py"))
        mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
        return mod

    def header_values(fn):
        txt = open(os.path.join(root, "mcmc.jl_b200", "csrc", fn)).read()
        return [float.fromhex(v) for v in re.findall(r"-?0x[0-9a-f.]+p[+-]\d+", txt)]

    ge, gl = load("gen_exp_table"), load("gen_log_table")
    T = ge.table()
    assert header_values("exp_table.h") == T
    R, L = gl.table()
    assert header_values("log_table.h") == [v for pair in zip(R, L) for v in pair]
    rng = np.random.default_rng(7)
    for x in np.concatenate([rng.uniform(-40, 40, 300), rng.uniform(-700, 700, 100)]):
        ref = mp.exp(mp.mpf(float(x)))
        assert abs(mp.mpf(ge.exp_dev(float(x), T)) - ref) <= mp.mpf(3e-16) * ref
    for x in np.concatenate([rng.uniform(1e-12, 1, 300), 1 - 10 ** rng.uniform(-15, -2, 100)]):
        ref = mp.log(mp.mpf(float(x)))
        assert abs(mp.mpf(gl.log_dev(float(x), R, L)) - ref) <= mp.mpf(3e-16) * abs(ref) + mp.mpf(3e-18)


def test_probit_table_is_reproducible():
    """csrc/probit_table.h holds what tools/gen_probit_table.py computes (log Phi and phi/Phi at the grid points in double,
    the float pair of the cubic / quartic tail), and the generator's operation-by-operation replay of K1's probit link
    (probit_eval, k1_regress.cu: grid index in float, W', W'' from W' = -W (z + W) in double, tail in float) is accurate to
    the figures its header states: relative for z < 0, absolute for z >= 0."""
    import importlib.util
    import re
    import struct
    import mpmath as mp
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_probit_table", os.path.join(root, "tools", "gen_probit_table.py"))
    g = importlib.util.module_from_spec(spec); spec.loader.exec_module(g)
    txt = open(os.path.join(root, "mcmc.jl_b200", "csrc", "probit_table.h")).read()
    assert f"#define PROBIT_INV_W {g.INV_W}" in txt and f"#define PROBIT_ZMAX {g.ZMAX}" in txt
    words = [int(w, 16) for w in re.findall(r"0x([0-9a-f]{16})ull", txt)]
    nint = 2 * g.ZMAX * g.INV_W + 1
    assert len(words) == 3 * nint
    as_double = lambda w: struct.unpack("<d", struct.pack("<Q", w))[0]
    as_floats = lambda w: struct.unpack("<ff", struct.pack("<Q", w))
    rng = np.random.default_rng(3)
    ks = np.concatenate([[0, 1, nint // 2, nint - 1], rng.integers(0, nint, 40)])
    tab = {}
    for k in ks:
        k = int(k)
        e = g.entry(k)
        assert (as_double(words[k]), as_double(words[nint + k])) == e[:2], k
        assert as_floats(words[2 * nint + k]) == (np.float32(e[2]), np.float32(e[3])), k
        tab[k] = e
    # the replay on the header's own numbers, at points of the checked intervals
    for k, e in tab.items():
        c = -g.ZMAX + k / g.INV_W
        for t in (0.0, 0.2 / g.INV_W, -0.49 / g.INV_W, 0.4999 / g.INV_W):
            z = c + t
            if not (-36.9 < z < 36.9) or int(np.rint(np.float32(z) * np.float32(g.INV_W))) + g.ZMAX * g.INV_W != k:
                continue
            f, w = g.device_eval(tab, z)
            rf, rw = g.F(mp.mpf(z)), g.W(mp.mpf(z))
            if z < 0:
                assert abs((mp.mpf(f) - rf) / rf) < 4e-16 and abs((mp.mpf(w) - rw) / rw) < 2e-15, (z, f, w)
            else:
                assert abs(mp.mpf(f) - rf) < 3e-16 and abs(mp.mpf(w) - rw) < 3e-15, (z, f, w)


def test_handle_lifetimes_children_first(capi, monkeypatch):
    """runs, models and contexts are released children-first -- by close(), by `with`, and by garbage
    collection -- and a closed parent leaves no dangling child handle (checked against a recording fake of the library)."""
    import ctypes
    import gc
    calls = []

    class FakeLib:
        def __getattr__(self, name):
            return lambda h: calls.append((name, h.value))

    monkeypatch.setattr(capi, "lib", lambda: FakeLib())

    def make(cls, handle, fn, parent):
        o = cls.__new__(cls)
        o.h = ctypes.c_void_p(handle)
        o._own(fn, parent)
        return o

    ctx = make(capi.Context, 1, "mcmcgpu_destroy", None)
    m = make(capi.DeviceModel, 2, "mcmcgpu_model_destroy", ctx); m.ctx = ctx
    r1 = make(capi.DeviceRun, 3, "mcmcgpu_run_destroy", m); r1.model = m
    r2 = make(capi.DeviceRun, 4, "mcmcgpu_run_destroy", m); r2.model = m
    r2.close(); r2.close()                                   # idempotent
    assert calls == [("mcmcgpu_run_destroy", 4)] and r2.h is None
    ctx.close()                                              # closes the live run, then the model, then the context
    assert calls[1:] == [("mcmcgpu_run_destroy", 3), ("mcmcgpu_model_destroy", 2), ("mcmcgpu_destroy", 1)]
    assert r1.h is None and m.h is None and ctx.h is None
    # garbage collection: dropping every reference releases run -> model -> context, in that order
    del calls[:]
    ctx = make(capi.Context, 10, "mcmcgpu_destroy", None)
    m = make(capi.DeviceModel, 20, "mcmcgpu_model_destroy", ctx)
    r = make(capi.DeviceRun, 30, "mcmcgpu_run_destroy", m)
    del ctx, m
    gc.collect()
    assert calls == []                                       # the run still keeps its model and context alive
    del r
    gc.collect()
    assert calls == [("mcmcgpu_run_destroy", 30), ("mcmcgpu_model_destroy", 20), ("mcmcgpu_destroy", 10)]
    with make(capi.Context, 11, "mcmcgpu_destroy", None) as c2:
        make(capi.DeviceModel, 21, "mcmcgpu_model_destroy", c2)
    gc.collect()
    assert calls[-2:] == [("mcmcgpu_model_destroy", 21), ("mcmcgpu_destroy", 11)]
