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


def _cond_tol(th, locs_i, X_i, locs_j, X_j, lim, classic=False):
    """Entry-wise tolerance: 1e-12 relative, widened where the ENTRY ITSELF is ill-conditioned.

    The per-site links go through libm (exp, sin, cos), where CUDA and glibc legitimately differ by an
    ulp.  An entry is ~exp(-Q), so a relative change d of Q moves it by Q d; and Q inherits the
    cancellation of det = s11 s22 - s12^2 (near-singular local kernels: tilt close to 0 or pi) and of the
    quadratic form.  tol = 1e-12 + 4 ulp * max(1, Q) * (kappa_det + kappa_quad), all evaluated here in
    plain numpy from the mathematical formulas."""
    def site(X):
        sje = np.array(th["scale"], dtype=float).copy()
        sje[0] = 0.0
        t = np.pi / (1 + np.exp(-(X @ th["tilt"])))
        eta = X @ th["smooth"]
        nu = np.exp(eta) if classic else (lim[1] - lim[0]) / (1 + np.exp(-eta)) + lim[0]
        return t, np.exp(2 * (X @ sje)), np.exp(X @ th["aniso"]), nu
    ti, ri, ai, nui = site(np.asarray(X_i, dtype=float))
    tj, rj, aj, nuj = site(np.asarray(X_j, dtype=float))
    s11 = (ri[:, None] + rj[None, :]) / 2
    s22 = ((ri * ai ** 2)[:, None] + (rj * aj ** 2)[None, :]) / 2
    s12 = ((ri * ai * np.cos(ti))[:, None] + (rj * aj * np.cos(tj))[None, :]) / 2
    det = s11 * s22 - s12 ** 2
    dx = locs_i[:, 0][:, None] - locs_j[:, 0][None, :]
    dy = locs_i[:, 1][:, None] - locs_j[:, 1][None, :]
    t1, t2, t3 = s22 * dx * dx, s11 * dy * dy, 2 * s12 * dx * dy
    quad = t1 + t2 - t3
    nu = (nui[:, None] + nuj[None, :]) / 2 if classic else np.sqrt(nui[:, None] * nuj[None, :])
    with np.errstate(divide="ignore", invalid="ignore"):
        Q = np.sqrt(8 * nu / (np.exp(2 * th["scale"][0]) * det)) * np.sqrt(np.abs(quad))
        kappa = s11 * s22 / np.abs(det) + (np.abs(t1) + np.abs(t2) + np.abs(t3)) / np.abs(quad)
    amp = np.where(np.isfinite(Q) & np.isfinite(kappa), np.maximum(1.0, Q) * kappa, 1.0)
    return 1e-12 + 4 * 2.2e-16 * amp


def _check(got, ref, tol):
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.where(ref != 0, np.abs(got - ref) / np.abs(ref), np.where(got == 0, 0.0, np.inf))
    # entries the reference itself builds from a subnormal exp(-Q) (Q > 708) carry no significant digits
    dust = np.abs(ref) < 1e-290
    rel = np.where(dust, np.where(np.abs(got) < 1e-280, 0.0, np.inf), rel)
    bad = rel > tol
    assert not bad.any(), (float(rel[bad].max()), float(tol[bad].min()), int(bad.sum()))
    return float((rel / tol).max()), float(rel.max())


def test_randomised_sweep_typical_models_flat_tolerance():
    """30 random models in the range fitted models live in (link coefficients within +-0.4, ranges such
    that Q stays below ~100): EVERY entry within a flat 1e-12, no conditioning allowance."""
    rng = np.random.default_rng(777)
    worst = 0.0
    for trial in range(30):
        n = int(rng.integers(20, 200))
        p = int(rng.integers(1, 6))
        locs = rng.uniform(-1, 1, (n, 2))
        X = np.column_stack([np.ones(n), rng.standard_normal((n, p - 1))])
        th = {k: rng.uniform(-0.4, 0.4, p) for k in cov.ASPECTS}
        th["scale"][0] = rng.uniform(-2.2, 0.5)
        th["nugget"][0] = rng.choice([-np.inf, -4.0, -1.0])
        if np.isneginf(th["nugget"][0]):
            th["nugget"][1:] = 0.0
        lim = [0.5, 2.5] if rng.random() < 0.7 else [float(rng.choice([0.5, 1.5, 2.5]))] * 2
        if lim[0] == lim[1]:
            th["smooth"] = np.zeros(p)
        e = relerr(cb.cov_rns(th, locs, X, lim), cov.cov_rns(th, locs, X, lim))
        m = int(rng.integers(1, 50))
        lp = rng.uniform(-1, 1, (m, 2))
        Xp = np.column_stack([np.ones(m), rng.standard_normal((m, p - 1))])
        e = max(e, relerr(cb.cov_rns_pred(th, locs, lp, X, Xp, lim), cov.cov_rns_pred(th, locs, lp, X, Xp, lim)))
        e = max(e, relerr(cb.cov_rns_classic(th, locs, X), cov.cov_rns_classic(th, locs, X)))
        worst = max(worst, e)
        assert e < TOL, (trial, n, p, lim, e)
    print("typical-model sweep: worst relative error %.2e" % worst)


def test_randomised_parameter_sweep_against_oracle():
    """40 random models (design width, link coefficients, smoothness limits, ranges spanning the Temme /
    CF2 / Hankel / >=706 bands, -Inf nuggets, duplicated sites, tilts up to the degenerate 0 / pi ends):
    every entry within 1e-12 relative, or within the entry's own 1-ulp conditioning where that is larger."""
    rng = np.random.default_rng(424242)
    worst = worst_plain = 0.0
    n_tight = n_all = 0
    for trial in range(40):
        n = int(rng.integers(20, 160))
        p = int(rng.integers(1, 6))
        locs = rng.uniform(-1, 1, (n, 2)) * rng.choice([0.05, 1.0, 20.0])
        X = np.column_stack([np.ones(n), rng.standard_normal((n, p - 1))])
        th = {k: rng.uniform(-0.6, 0.6, p) for k in cov.ASPECTS}
        th["scale"][0] = rng.uniform(-5.0, 1.5)
        th["tilt"] = rng.uniform(-2.5, 2.5, p)
        th["nugget"][0] = rng.choice([-np.inf, -6.0, -2.0, 0.5])
        if np.isneginf(th["nugget"][0]):
            th["nugget"][1:] = 0.0
        lo = rng.uniform(0.1, 1.5)
        lim = [lo, lo + rng.choice([0.0, 0.3, 1.0, 3.5])]
        if rng.random() < 0.3:
            th["smooth"] = np.zeros(p)
            lim = [rng.choice([0.5, 1.5, 2.5, 0.8]), 0.0]
            lim[1] = lim[0]
        if rng.random() < 0.3 and n > 4:
            locs[n - 1] = locs[1]
            locs[n // 2] = locs[0]
        m = int(rng.integers(1, 40))
        lp = rng.uniform(-1, 1, (m, 2)) * np.abs(locs).max()
        lp[0] = locs[0]
        Xp = np.column_stack([np.ones(m), rng.standard_normal((m, p - 1))])
        thc = dict(th, smooth=rng.uniform(-0.5, 0.8, p))
        for got, ref, tol in (
                (cb.cov_rns(th, locs, X, lim), cov.cov_rns(th, locs, X, lim), _cond_tol(th, locs, X, locs, X, lim)),
                (cb.cov_rns_pred(th, locs, lp, X, Xp, lim), cov.cov_rns_pred(th, locs, lp, X, Xp, lim),
                 _cond_tol(th, lp, Xp, locs, X, lim)),
                (cb.cov_rns_classic(thc, locs, X), cov.cov_rns_classic(thc, locs, X),
                 _cond_tol(thc, locs, X, locs, X, lim, classic=True))):
            ratio, plain = _check(got, ref, tol)
            worst, worst_plain = max(worst, ratio), max(worst_plain, plain)
            n_tight += int((tol <= 2e-12).sum())
            n_all += tol.size
    print("sweep: worst error / tolerance %.2f, worst plain relative error %.2e, %.1f%% of %d entries held to <= 2e-12"
          % (worst, worst_plain, 100.0 * n_tight / n_all, n_all))
