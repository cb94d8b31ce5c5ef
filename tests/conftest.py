import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


@pytest.fixture(scope="session")
def dp():
    import dpomp_b200

    return dpomp_b200


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle

    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def built_lib():
    """libdpomp.so built in-tree (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as ge

    ge.build()
    import dpomp_b200

    return dpomp_b200._capi.lib()


def load_case(dp, name):
    """(model, observations, private model, theta) of the named config."""
    import numpy as np

    cases = {
        "sis_pooley": ("SIS", [100, 1], "pooley.csv", [0.003, 0.1]),
        "sir_c2": ("SIR", [100, 1, 0], "sir_c2.csv", [0.003, 0.1]),
        "sir_dense": ("SIR", [1000, 10, 0], "sir_dense.csv", [0.0003, 0.1]),
        "seir_c3": ("SEIR", [100, 0, 1, 0], "seir_c3.csv", [0.005, 0.2, 0.1]),
        "lotka_c4": ("LOTKA", [70, 70], "lotka_c4.csv", [0.5, 0.0025, 0.3]),
    }
    mname, ic, csv, theta = cases[name]
    model = dp.generate_model(mname, ic)
    y = dp.get_observations(os.path.join(GOLDEN, csv))
    return model, y, dp.get_private_model(model, y), np.asarray(theta, dtype=np.float64)


def pure_death_case(dp):
    """Pure-death process (SIS with theta_1 = 0) with five Gaussian observations and its EXACT log-likelihood by the forward
    recursion over the 61 hidden states (binomial thinning between observations).  Returns (hmm, theta, exact log-lik)."""
    import numpy as np
    from scipy import stats

    model = dp.generate_model("SIS", [40, 60])
    ys = [47, 36, 29, 22, 18]
    y = [dp.Observation(5.0 * (k + 1), 1, 1.0, [0, v]) for k, v in enumerate(ys)]
    gam, sigma, i0 = 0.05, 2.0, 60
    p = np.exp(-gam * 5.0)
    states = np.arange(i0 + 1)
    trans = np.array([stats.binom.pmf(states, k, p) for k in states])  # trans[k, j] = P(j survivors | k)
    alpha = np.zeros(i0 + 1)
    alpha[i0] = 1.0
    ll = 0.0
    for v in ys:
        alpha = alpha @ trans
        alpha = alpha * np.exp(np.log(1.0 / (np.sqrt(2 * np.pi) * sigma)) - (v - states) ** 2 / (2 * sigma * sigma))
        ll += np.log(alpha.sum())
        alpha /= alpha.sum()
    return model, y, dp.get_private_model(model, y), np.array([0.0, gam]), float(ll)


def death_rate_case(dp, nodes: int = 2001):
    """One-parameter pure-death model (rate theta * I, prior U(0, 0.2)) with the observations of pure_death_case, and the
    EXACT likelihood on a grid of death rates (forward recursion over the 61 hidden states).  Returns a dict with the model,
    the observations, the grid `g`, the likelihood `lik` on it, and by trapezoid quadrature the exact -ln p(y) (`bme`),
    posterior mean (`mean`) and standard deviation (`sd`)."""
    import numpy as np
    from scipy import stats

    def rf(out, p, x):
        out[0] = p[0] * x[1]
    model = dp.generate_custom_model("DEATH", rf, [40, 60], [[1, -1]], prior=dp.UniformProduct([0.0], [0.2]))
    ys = [47, 36, 29, 22, 18]
    y = [dp.Observation(5.0 * (k + 1), 1, 1.0, [0, v]) for k, v in enumerate(ys)]
    states, sigma = np.arange(61), 2.0

    def exact_ll(gam):
        trans = stats.binom.pmf(states[None, :], states[:, None], np.exp(-gam * 5.0))
        alpha = np.zeros(61)
        alpha[60] = 1.0
        ll = 0.0
        for v in ys:
            alpha = (alpha @ trans) * np.exp(np.log(1.0 / (np.sqrt(2 * np.pi) * sigma)) - (v - states) ** 2 / (2 * sigma * sigma))
            ll += np.log(alpha.sum())
            alpha /= alpha.sum()
        return ll
    g = np.linspace(0.0, 0.2, nodes)
    lik = np.exp(np.array([exact_ll(v) for v in g]))
    z = np.trapezoid(lik, g)
    mean = np.trapezoid(lik * g, g) / z
    return dict(model=model, y=y, hmm=dp.get_private_model(model, y), cm=dp.compile_model(model, y), g=g, lik=lik,
                bme=float(-np.log(z / 0.2)), mean=float(mean), sd=float(np.sqrt(np.trapezoid(lik * (g - mean) ** 2, g) / z)))
