import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=True)["cases"].item()


@pytest.fixture(scope="session")
def O():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def capi():
    import mcmc_jl_b200  # noqa: F401
    from mcmc_jl_b200 import _capi
    return _capi


@pytest.fixture(scope="session")
def ctx(capi):
    """GPU context: fails loudly (no skip, no fallback) when the library or the device is missing."""
    c = capi.Context(0)
    yield c
    c.close()


def make_regression(fam, N, d, seed):
    r = np.random.default_rng(seed)
    X = np.concatenate([np.ones((N, 1)), r.standard_normal((N, d - 1))], axis=1) if d > 1 else np.ones((N, 1))
    b0 = r.standard_normal(d) / np.sqrt(d)
    eta = X @ b0
    if fam == "linear":
        y, hy = eta + r.standard_normal(N), (1.0, 1.0)
    elif fam == "logistic":
        y, hy = (r.random(N) < 1 / (1 + np.exp(-eta))).astype(float), (1.0, -1.0)
    else:
        from scipy.special import ndtr
        y, hy = (r.random(N) < ndtr(eta)).astype(float), (10.0,)
    return X, y, hy, b0


def ou_series(T, seed):
    r = np.random.default_rng(seed)
    x = np.empty(T)
    x[0] = 1.0
    for i in range(1, T):
        x[i] = x[i - 1] * np.exp(-1 / 20) + 10 * (1 - np.exp(-1 / 20)) + 0.1 * r.standard_normal()
    return x
