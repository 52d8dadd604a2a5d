"""ctypes binding of the CPU oracle (oracle/mcmc_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs; never by the product package.  "parity unpinned": see mcmc_oracle.h.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libmcmc_oracle.so")

FAM = dict(normal_fn=0, normal_dsl=1, linear=2, logistic=3, probit=4, ou=5, abs_normal=6)
KIND = dict(RWM=0, MALA=1, HMC=2, HMCDA=3, RAM=4)


def build(force=False):
    src = os.path.join(_HERE, "mcmc_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libmcmc_oracle.so"])
    return _LIB


class _Model(C.Structure):
    _fields_ = [("family", C.c_int32), ("N", C.c_int64), ("d", C.c_int64),
                ("X", C.c_void_p), ("y", C.c_void_p), ("hyper", C.c_double * 4)]


class _Sampler(C.Structure):
    _fields_ = [("kind", C.c_int32), ("scale", C.c_double), ("nleaps", C.c_int32),
                ("rate", C.c_double), ("len", C.c_double), ("shrinkage", C.c_double),
                ("t0", C.c_double), ("step", C.c_double), ("max_leaps", C.c_int64),
                ("tuner_on", C.c_int32), ("adapt_step", C.c_int32), ("max_step", C.c_int32),
                ("target_path", C.c_double), ("target_rate", C.c_double), ("force_eps", C.c_void_p), ("rb_out", C.c_void_p),
                ("start_step", C.c_int64), ("da_state", C.c_void_p)]


class _Range(C.Structure):
    _fields_ = [("first", C.c_int64), ("step", C.c_int64), ("last", C.c_int64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        _lib.orc_eval.restype = C.c_double
        _lib.orc_eval.argtypes = [C.POINTER(_Model), dp]
        _lib.orc_evalallg.restype = C.c_double
        _lib.orc_evalallg.argtypes = [C.POINTER(_Model), dp, dp]
        _lib.orc_log_ndtr.restype = C.c_double
        _lib.orc_log_ndtr.argtypes = [C.c_double]
        _lib.orc_run_chain.restype = C.c_int32
        _lib.orc_run_chain.argtypes = [C.POINTER(_Model), C.POINTER(_Sampler), C.POINTER(_Range), dp, dp, dp, dp,
                                       dp, dp, C.POINTER(C.c_uint8), dp, dp, C.POINTER(C.c_int64),
                                       C.POINTER(C.c_int64)]
        for f in ("orc_mean", "orc_mcvar_iid"):
            getattr(_lib, f).restype = C.c_double
            getattr(_lib, f).argtypes = [dp, C.c_int64]
        for f in ("orc_mcvar_bm", "orc_mcvar_imse", "orc_mcvar_ipse"):
            getattr(_lib, f).restype = C.c_double
            getattr(_lib, f).argtypes = [dp, C.c_int64, C.c_int64]
        _lib.orc_acov.restype = None
        _lib.orc_acov.argtypes = [dp, C.c_int64, C.c_int64, dp]
        _lib.orc_philox4x32_10.restype = None
        _lib.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        _lib.orc_draw_normals.restype = None
        _lib.orc_draw_normals.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int64, dp]
        _lib.orc_draw_uniform.restype = C.c_double
        _lib.orc_draw_uniform.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
    return _lib


def set_fast_baseline(on):
    """row-blocked logistic evaluation for the timed CPU baseline (bench.py); never used by parity tests"""
    lib().orc_set_fast_baseline.argtypes = [C.c_int]
    lib().orc_set_fast_baseline.restype = None
    lib().orc_set_fast_baseline(1 if on else 0)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


class Model:
    """family + data; X is N x d (any layout, copied to column-major), y length N."""

    def __init__(self, family, d, X=None, y=None, hyper=()):
        self.family = family
        self.d = int(d)
        self.X = None if X is None else np.asfortranarray(X, dtype=np.float64)
        self.y = None if y is None else np.ascontiguousarray(y, dtype=np.float64)
        self.N = 0 if self.y is None and self.X is None else (len(self.y) if self.y is not None else self.X.shape[0])
        self.hyper = tuple(float(h) for h in hyper)
        m = _Model()
        m.family = FAM[family]
        m.N = self.N
        m.d = self.d
        m.X = self.X.ctypes.data if self.X is not None else None
        m.y = self.y.ctypes.data if self.y is not None else None
        for i, h in enumerate(self.hyper):
            m.hyper[i] = h
        self._c = m

    def eval(self, beta):
        b = np.ascontiguousarray(beta, dtype=np.float64)
        return lib().orc_eval(C.byref(self._c), _dp(b))

    def evalallg(self, beta):
        b = np.ascontiguousarray(beta, dtype=np.float64)
        g = np.empty(self.d)
        lt = lib().orc_evalallg(C.byref(self._c), _dp(b), _dp(g))
        return lt, g


def sampler(kind, scale=1.0, nleaps=10, rate=0.65, len=2.0, shrinkage=0.05, t0=10.0, step=0.75, max_leaps=0,
            tuner=None, force_eps=None, rb_out=None, start_step=0, da_state=None):
    s = _Sampler()
    s.start_step = int(start_step)
    if da_state is not None:
        s._da = np.ascontiguousarray(da_state, dtype=np.float64)      # (leapStep, dualLeapStep, dualH); keep alive
        assert s._da.shape == (3,)
        s.da_state = s._da.ctypes.data
    if rb_out is not None:
        assert rb_out.flags["C_CONTIGUOUS"] and rb_out.dtype == np.float64
        s._rb = rb_out
        s.rb_out = rb_out.ctypes.data
    if force_eps is not None:
        s._keep = np.ascontiguousarray(force_eps, dtype=np.float64)  # keep alive
        s.force_eps = s._keep.ctypes.data
    s.kind = KIND[kind]
    s.scale = scale
    s.nleaps = nleaps
    s.rate, s.len, s.shrinkage, s.t0, s.step, s.max_leaps = rate, len, shrinkage, t0, step, max_leaps
    if tuner is not None:
        s.tuner_on = 1
        s.adapt_step = tuner.get("adapt_step", 100)
        s.max_step = tuner.get("max_step", 200)
        s.target_path = tuner.get("target_path", 1.0)
        s.target_rate = tuner["target_rate"]
    return s


def range_length(first, step, last):
    return 0 if (step < 1 or last < first) else (last - first) // step + 1


def run_chain(model, smp, rng, init, scale=None, normals=None, uniforms=None):
    """rng = (first, step, last).  normals: (last+1, d) array [row i = step i; row 0 = pre-loop draw];
    uniforms: (last+1,).  Returns dict(samples (S,d), grads (S,d), accept (S,), logtarget (S,), eps, nleaps,
    n_grad_evals, rc)."""
    first, step, last = rng
    d = model.d
    S = range_length(first, step, last)
    r = _Range(first, step, last)
    init = np.ascontiguousarray(init, dtype=np.float64)
    scale = np.ones(d) if scale is None else np.ascontiguousarray(np.broadcast_to(scale, (d,)), dtype=np.float64)
    normals = np.ascontiguousarray(np.asarray(normals, dtype=np.float64)[:max(last, 0) + 1])
    uniforms = np.ascontiguousarray(np.asarray(uniforms, dtype=np.float64)[:max(last, 0) + 1])
    assert normals.shape == (max(last, 0) + 1, d) and uniforms.shape == (max(last, 0) + 1,)
    samples = np.full((S, d), np.nan)
    grads = np.full((S, d), np.nan)
    accept = np.zeros(S, dtype=np.uint8)
    lt = np.full(S, np.nan)
    eps = np.full(S, np.nan)
    nl = np.zeros(S, dtype=np.int64)
    nev = C.c_int64(0)
    rc = lib().orc_run_chain(C.byref(model._c), C.byref(smp), C.byref(r), _dp(init), _dp(scale), _dp(normals),
                             _dp(uniforms), _dp(samples), _dp(grads), accept.ctypes.data_as(C.POINTER(C.c_uint8)),
                             _dp(lt), _dp(eps), nl.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(nev))
    return dict(samples=samples, grads=grads, accept=accept, logtarget=lt, eps=eps, nleaps=nl,
                n_grad_evals=nev.value, rc=rc)


def _arrays(models, samplers):
    ms = (_Model * len(models))(*[m._c for m in models])
    ss = (_Sampler * len(samplers))(*samplers)
    return ms, ss


def run_seqmc(models, samplers, steps, burnin, trigger, particles, normals, uniforms, res_uniforms):
    """particles (npart, d); normals (steps, nt, npart, d); uniforms / res_uniforms (steps, nt, npart)."""
    nt, d = len(models), models[0].d
    particles = np.ascontiguousarray(particles, dtype=np.float64)
    npart = particles.shape[0]
    normals = np.ascontiguousarray(normals, dtype=np.float64); uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
    res_uniforms = np.ascontiguousarray(res_uniforms, dtype=np.float64)
    assert normals.shape == (steps, nt, npart, d) and uniforms.shape == (steps, nt, npart) == res_uniforms.shape
    S = (steps - burnin) * npart
    samples = np.full((max(S, 0), d), np.nan); weights = np.full(max(S, 0), np.nan)
    nres = C.c_int64(0)
    ms, ss = _arrays(models, samplers)
    L = lib()
    L.orc_run_seqmc.restype = C.c_int32
    rc = L.orc_run_seqmc(ms, ss, C.c_int32(nt), C.c_int64(steps), C.c_int64(burnin), C.c_double(trigger), C.c_int64(npart),
                         _dp(particles), _dp(normals), _dp(uniforms), _dp(res_uniforms), _dp(samples), _dp(weights), C.byref(nres))
    return dict(samples=samples, weights=weights, n_resamples=nres.value, rc=rc)


def run_serialtemp(models, samplers, steps, burnin, swap_period, inits, normals, uniforms, u_pick, u_swap):
    """inits (nt, d); normals (steps+2, d); uniforms (steps+2,); u_pick, u_swap (steps+1,)."""
    nt, d = len(models), models[0].d
    inits = np.ascontiguousarray(inits, dtype=np.float64)
    normals = np.ascontiguousarray(normals, dtype=np.float64); uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
    u_pick = np.ascontiguousarray(u_pick, dtype=np.float64); u_swap = np.ascontiguousarray(u_swap, dtype=np.float64)
    assert inits.shape == (nt, d) and normals.shape == (steps + 2, d) and uniforms.shape == (steps + 2,)
    assert u_pick.shape == (steps + 1,) == u_swap.shape
    S = steps - burnin
    samples = np.full((max(S, 0), d), np.nan); at = np.zeros(max(S, 0), dtype=np.int32)
    ms, ss = _arrays(models, samplers)
    L = lib()
    L.orc_run_serialtemp.restype = C.c_int32
    rc = L.orc_run_serialtemp(ms, ss, C.c_int32(nt), C.c_int64(steps), C.c_int64(burnin), C.c_int64(swap_period), _dp(inits),
                              _dp(normals), _dp(uniforms), _dp(u_pick), _dp(u_swap), _dp(samples),
                              at.ctypes.data_as(C.POINTER(C.c_int32)))
    return dict(samples=samples, at=at, rc=rc)


def mean(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    return lib().orc_mean(_dp(x), len(x))


def mcvar(x, vtype="imse", maxlag=None, batchlen=100):
    x = np.ascontiguousarray(x, dtype=np.float64)
    n = len(x)
    if vtype == "iid":
        return lib().orc_mcvar_iid(_dp(x), n)
    if vtype == "bm":
        return lib().orc_mcvar_bm(_dp(x), n, batchlen)
    ml = n - 1 if maxlag is None else maxlag
    f = lib().orc_mcvar_imse if vtype == "imse" else lib().orc_mcvar_ipse
    return f(_dp(x), n, ml)


def ess(x, vtype="imse", **kw):  # ess.jl:6-10
    return len(x) * mcvar(x, "iid") / mcvar(x, vtype, **kw)


def actime(x, vtype="imse", **kw):  # ess.jl:15-19
    return mcvar(x, vtype, **kw) / mcvar(x, "iid")


def acov(x, maxlag):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty(maxlag + 1)
    lib().orc_acov(_dp(x), len(x), maxlag, _dp(out))
    return out


def zv(x, grad, order=1):
    """x, grad: (S, d).  Returns (zvchain (S, d), a (k, d)) like linearZv / quadraticZv (zv.jl:8-66)."""
    x = np.ascontiguousarray(x, dtype=np.float64); g = np.ascontiguousarray(grad, dtype=np.float64)
    S, d = x.shape
    k = d if order == 1 else d * (d + 3) // 2
    out, a = np.empty((S, d)), np.empty((k, d))
    L = lib()
    L.orc_zv.restype = C.c_int32
    rc = L.orc_zv(_dp(x), _dp(g), C.c_int64(S), C.c_int64(d), C.c_int32(order), _dp(out), _dp(a))
    if rc != 0:
        raise RuntimeError(f"orc_zv rc={rc}")
    return out, a


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return tuple(o)


def draw_normals(seed, chain, step, d):
    z = np.empty(d)
    lib().orc_draw_normals(seed, chain, step, d, _dp(z))
    return z


def draw_uniform(seed, chain, step):
    return lib().orc_draw_uniform(seed, chain, step)
