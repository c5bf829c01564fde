"""GPU parity of the covariance builders, through the C ABI, against (a) the committed golden
matrices produced by the reference's own compiled source and (b) the oracle on fresh seeded
inputs.  Bar: every entry within 1e-12 relative (BASELINE.json north_star), exact zeros exact."""
import numpy as np
import pytest

import cocons_b200 as cb
from conftest import relerr, theta_dict
from oracle import cov

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _gpu(case):
    th = theta_dict(case["theta6"])
    if "locs_pred" in case:
        return cb.cov_rns_pred(th, case["locs"], case["locs_pred"], case["X"], case["X_pred"], case["limits"])
    if "limits" in case:
        return cb.cov_rns(th, case["locs"], case["X"], case["limits"])
    return cb.cov_rns_classic(th, case["locs"], case["X"])


def test_golden_cases(cov_cases):
    report = {}
    for name, case in cov_cases.items():
        got = _gpu(case)
        assert got.shape == case["out"].shape
        report[name] = relerr(got, case["out"])
    bad = {k: v for k, v in report.items() if not v < TOL}
    assert not bad, "relative error above 1e-12: %s (all: %s)" % (bad, report)
    print("max rel err per golden case:", {k: "%.1e" % v for k, v in report.items()})


def test_quirks_survive_on_the_device(cov_cases):
    S = _gpu(cov_cases["degenerate_nu1_fixed"])
    assert S[3, 50] == S[3, 3] and S[50, 3] == S[3, 3]  # App. B-1
    S = _gpu(cov_cases["general_duplicates"])
    assert S[5, 90] == S[5, 5] and S[90, 5] == S[5, 5] and S[17, 100] == S[17, 17]  # App. B-2
    assert np.array_equal(S, S.T)


@pytest.mark.parametrize("n,p,seed", [(1, 1, 0), (2, 2, 1), (127, 3, 2), (128, 3, 3), (129, 4, 4), (700, 5, 5)])
def test_against_oracle_on_seeded_inputs(n, p, seed):
    rng = np.random.default_rng(seed)
    locs = rng.uniform(-1, 1, (n, 2))
    X = np.column_stack([np.ones(n), rng.standard_normal((n, p - 1))])
    th = {k: 0.25 * rng.standard_normal(p) for k in cov.ASPECTS}
    th["scale"][0] = -1.4
    th["nugget"][0] = -3.0
    for lim in ([0.5, 2.5], [1.5, 1.5], [0.3, 0.9]):
        ref = cov.cov_rns(th, locs, X, lim)
        got = cb.cov_rns(th, locs, X, lim)
        assert relerr(got, ref) < TOL, (lim, relerr(got, ref))
        assert np.array_equal(got, got.T)
    th2 = dict(th)
    th2["smooth"] = np.zeros(p)
    for lim in ([0.5, 0.5], [1.5, 1.5], [2.5, 2.5], [1.0, 1.0]):
        assert relerr(cb.cov_rns(th2, locs, X, lim), cov.cov_rns(th2, locs, X, lim)) < TOL, lim
    assert relerr(cb.cov_rns_classic(th, locs, X), cov.cov_rns_classic(th, locs, X)) < TOL
    m = max(1, n // 3)
    lp = rng.uniform(-1, 1, (m, 2))
    lp[0] = locs[n // 2]
    Xp = np.column_stack([np.ones(m), rng.standard_normal((m, p - 1))])
    got = cb.cov_rns_pred(th, locs, lp, X, Xp, [0.5, 2.5])
    assert got.shape == (m, n)
    assert relerr(got, cov.cov_rns_pred(th, locs, lp, X, Xp, [0.5, 2.5])) < TOL


def test_argument_errors():
    th = {k: np.zeros(2) for k in cov.ASPECTS}
    bad = dict(th)
    del bad["tilt"]
    with pytest.raises(KeyError):  # missing aspect name is an error, as in src/cocons_full.cpp:47-54
        cb.cov_rns(bad, np.zeros((3, 2)), np.ones((3, 2)), [0.5, 0.5])
    extra = dict(th, mean=np.zeros(2))  # extra names are tolerated (getCovMatrix passes "mean")
    out = cb.cov_rns(extra, np.array([[0.0, 0], [1, 0], [0, 1]]), np.column_stack([np.ones(3), np.arange(3.0)]),
                     [0.5, 0.5])
    assert out.shape == (3, 3)
    # integer inputs are coerced like Rcpp does
    out2 = cb.cov_rns(extra, np.array([[0, 0], [1, 0], [0, 1]]), np.column_stack([np.ones(3), np.arange(3)]),
                      [0.5, 0.5])
    assert np.array_equal(out, out2)
