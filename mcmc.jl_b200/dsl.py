"""Front-end recogniser for the model DSL of MCMC.jl.

The reference turns `quote ... ~ ... end` into Julia code through ReverseDiffSource (src/dsl/expr_funcs.jl:8-36,
modelparser.jl:39-104).  Arbitrary expressions cannot run on the GPU, so this module *recognises* the model shapes
of the reference's README and examples/ and maps them onto the built-in likelihood families; anything else is
refused with the list of shapes it knows.  `model(text, gradient=..., <param>=init, <data>=array, ...)` keeps the
reference's calling convention: keyword arguments that are sampled (`~`) without being assigned are parameters
(expr_funcs.jl:76-91 modelVars), the rest are data the expression refers to.
"""
import re

import numpy as np

_NUM = r"[-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+)"
_ID = r"[A-Za-z_]\w*"


def _statements(text):
    out = []
    for line in text.replace(";", "\n").splitlines():
        line = line.split("#")[0].strip()
        if not line or line in ("quote", "begin", "end"):
            continue
        out.append(re.sub(r"\s+", "", line))
    return out


def _m(pattern, s):
    return re.fullmatch(pattern, s)


SHAPES = """recognised model shapes:
  v ~ Normal(mu, sigma)                                                      (README.md:67-72)
  y = abs(x); y ~ Normal(mu, sigma)                                          (README.md:253-259, the SeqMC ladder)
  vars ~ Normal(0, s); resid = Y - X * vars; resid ~ Normal(0, s2)           (examples/linear_regression.jl:14-18)
  vars ~ Normal(0, s); prob = 1 / (1. + exp(-X * vars)); Y ~ Bernoulli(prob) (examples/logistic_regression.jl:16-20;
                                                                              exp(X * vars) as in test/test_syntax.jl:18)
  tau ~ Uniform(0,a); sigma ~ Uniform(0,b); mu ~ Uniform(0,c); fac = exp(- 1. / tau);
  resid = x[2:end] - x[1:end-1] * fac - mu * (1. - fac); resid ~ Normal(0, sigma)   (examples/ornstein.jl:19-27)"""


def recognise(text, kwargs):
    """-> dict(family, init, pmap, X, y, hyper)"""
    st = _statements(text)
    kw = dict(kwargs)

    def need(name):
        if name not in kw:
            raise ValueError(f"the expression refers to `{name}`: pass it as a keyword argument")
        return kw[name]

    # --- single Normal statement
    if len(st) == 1:
        m = _m(rf"(?P<v>{_ID})~Normal\((?P<mu>{_NUM}),(?P<sd>{_NUM})\)", st[0])
        if m:
            v0 = np.atleast_1d(np.asarray(need(m["v"]), dtype=np.float64))
            return dict(family="normal_dsl", init=v0, pmap={m["v"]: (1, tuple(np.shape(need(m["v"]))))}, X=None, y=None,
                        hyper=(float(m["mu"]), float(m["sd"])))
    # --- folded Normal of the SeqMC example (README.md:253-259)
    if len(st) == 2:
        a = _m(rf"(?P<y>{_ID})=abs\((?P<x>{_ID})\)", st[0])
        if a:
            n = _m(rf"{a['y']}~Normal\((?P<mu>{_NUM}),(?P<sd>{_NUM})\)", st[1])
            if n:
                v0 = np.atleast_1d(np.asarray(need(a["x"]), dtype=np.float64))
                return dict(family="abs_normal", init=v0, pmap={a["x"]: (1, tuple(np.shape(need(a["x"]))))}, X=None, y=None,
                            hyper=(float(n["mu"]), float(n["sd"])))
    # --- regressions
    if len(st) == 3:
        prior = _m(rf"(?P<b>{_ID})~Normal\(0(?:\.0*)?,(?P<sd>{_NUM})\)", st[0])
        if prior:
            b = prior["b"]
            lin = _m(rf"(?P<r>{_ID})=(?P<Y>{_ID})-(?P<X>{_ID})\*{b}", st[1])
            if lin:
                lik = _m(rf"{lin['r']}~Normal\(0(?:\.0*)?,(?P<sd>{_NUM})\)", st[2])
                if lik:
                    v0 = np.asarray(need(b), dtype=np.float64)
                    return dict(family="linear", init=v0, pmap={b: (1, v0.shape)}, X=np.asarray(need(lin["X"]), dtype=np.float64),
                                y=np.asarray(need(lin["Y"]), dtype=np.float64), hyper=(float(prior["sd"]), float(lik["sd"])))
            lg = _m(rf"(?P<p>{_ID})=1/\(1\.?0*\+exp\((?P<sign>-?)(?P<X>{_ID})\*{b}\)\)", st[1])
            if lg:
                lik = _m(rf"(?P<Y>{_ID})~Bernoulli\({lg['p']}\)", st[2])
                if lik:
                    v0 = np.asarray(need(b), dtype=np.float64)
                    return dict(family="logistic", init=v0, pmap={b: (1, v0.shape)}, X=np.asarray(need(lg["X"]), dtype=np.float64),
                                y=np.asarray(need(lik["Y"]), dtype=np.float64),
                                hyper=(float(prior["sd"]), -1.0 if lg["sign"] == "-" else 1.0))
    # --- Ornstein-Uhlenbeck
    if len(st) == 6:
        u = [_m(rf"(?P<v>{_ID})~Uniform\(0(?:\.0*)?,(?P<hi>{_NUM})\)", s) for s in st[:3]]
        if all(u):
            tau, sig, mu = (x["v"] for x in u)
            f = _m(rf"(?P<f>{_ID})=exp\(-1\.?0*/{tau}\)", st[3])
            if f:
                r = _m(rf"(?P<r>{_ID})=(?P<x>{_ID})\[2:end\]-(?P=x)\[1:end-1\]\*{f['f']}-{mu}\*\(1\.?0*-{f['f']}\)", st[4])
                if r and _m(rf"{r['r']}~Normal\(0(?:\.0*)?,{sig}\)", st[5]):
                    names = [k for k in kw if k in (tau, sig, mu)]
                    if names != [tau, sig, mu]:
                        raise ValueError(f"give the parameters in the order of the model: {tau}=, {sig}=, {mu}=")
                    v0 = np.array([kw[tau], kw[sig], kw[mu]], dtype=np.float64)
                    return dict(family="ou", init=v0, pmap={tau: (1, ()), sig: (2, ()), mu: (3, ())}, X=None,
                                y=np.asarray(need(r["x"]), dtype=np.float64), hyper=tuple(float(x["hi"]) for x in u))
    raise NotImplementedError("model expression not recognised; " + SHAPES)
