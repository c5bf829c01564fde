"""GPU parity of the three objectives (GetNeg2loglikelihood / Profile / REML), the Cholesky
factor and the DMMA trailing-update kernel.  Bar: -2 loglik within 1e-8 relative of the oracle
(BASELINE.json north_star); the tests assert 1e-9 and print what was reached."""
import numpy as np
import pytest

import cocons_b200 as cb
from cocons_b200 import _lib
from conftest import case_design, relerr
from oracle import cov, rmirror

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _values(c, locs, X, z, ctx=None):
    n, p = c["n"], c["p"]
    lam = c["lambda"]
    out = {}
    if "ml" in c["values"]:
        out["ml"] = cb.GetNeg2loglikelihood(c["theta"], c["par_pos"], locs, X, c["limits"], z, n, lam, ctx=ctx)
    ppm = dict(c["par_pos"])
    ppm["mean"] = np.zeros(p, dtype=bool)
    th = c["theta"][p:]
    if "profile" in c["values"]:
        if ctx is not None:
            ctx.set_xbetas(X)
        out["profile"] = cb.GetNeg2loglikelihoodProfile(th, ppm, locs, X, c["limits"], z, n, X, lam, ctx=ctx)
    if "reml" in c["values"]:
        zc = cb.reml_contrasts(X, z)
        if ctx is not None:
            ctx.set_z(zc)
        out["reml"] = cb.GetNeg2loglikelihoodREML(th, ppm, locs, X, X, c["limits"], zc, n, lam, ctx=ctx)
        if ctx is not None:
            ctx.set_z(z)
    return out


SMALL = ["holes1500_nu15", "holes1500_general", "holes1500_general_pen", "holes777_ragged", "holesbm1000_r10",
         "stripes2000_p4"]


@pytest.mark.parametrize("name", SMALL)
def test_objectives_one_shot_host_buffers(name, n2ll_cases, datasets):
    c = n2ll_cases[name]
    locs, X, z = case_design(c, datasets)
    got = _values(c, locs, X, z)
    errs = {k: abs(got[k] - c["values"][k]) / abs(c["values"][k]) for k in got}
    print(name, {k: "%.2e" % v for k, v in errs.items()})
    assert all(v < TOL for v in errs.values()), errs


@pytest.mark.parametrize("name", ["holes1500_general", "holesbm1000_r10"])
def test_objectives_resident_context(name, n2ll_cases, datasets):
    c = n2ll_cases[name]
    locs, X, z = case_design(c, datasets)
    with cb.DenseLikelihood(locs, X, z) as ctx:
        got = _values(c, locs, X, z, ctx=ctx)
        again = _values(c, locs, X, z, ctx=ctx)
        t = ctx.timings()
    assert got == again  # deterministic, bit for bit
    assert t["total_ms"] > 0
    for k in got:
        assert abs(got[k] - c["values"][k]) < TOL * abs(c["values"][k]), (k, got[k], c["values"][k])


def test_profile_betas(n2ll_cases, datasets):
    c = n2ll_cases["holes1500_general"]
    locs, X, z = case_design(c, datasets)
    p = c["p"]
    ppm = dict(c["par_pos"])
    ppm["mean"] = np.zeros(p, dtype=bool)
    tl = cb.getModelLists(c["theta"][p:], ppm, "diff")
    with cb.DenseLikelihood(locs, X, z) as ctx:
        ctx.set_xbetas(X)
        ctx.factor(tl, c["limits"])
        betas = ctx.profile_betas(_lib.PROFILE)
    assert relerr(betas, np.array(c["values"]["betas"])) < 1e-8


def test_not_positive_definite_follows_the_safe_logic(n2ll_cases, datasets):
    c = n2ll_cases["holes300_notpd"]
    locs, X, z = case_design(c, datasets)
    v = cb.GetNeg2loglikelihood(c["theta"], c["par_pos"], locs, X, c["limits"], z, c["n"], c["lambda"], safe=True)
    assert v == 1e6 == c["values"]["ml"]  # R/neg2loglikelihood.R:202-206
    with pytest.raises(ArithmeticError, match="Cholesky error"):
        cb.GetNeg2loglikelihood(c["theta"], c["par_pos"], locs, X, c["limits"], z, c["n"], c["lambda"], safe=False)
    # NaN in the inputs must fail the factorisation too
    Xn = X.copy()
    Xn[5, 1] = np.nan
    assert cb.GetNeg2loglikelihood(c["theta"], c["par_pos"], locs, Xn, [1.5, 1.5], z, c["n"], c["lambda"]) == 1e6


@pytest.mark.parametrize("name", ["holes_full_nu15", "holes_full_general", "stripes_full_general"])
def test_full_size_configs_against_committed_goldens(name, n2ll_cases, datasets):
    """BASELINE.json configs[0] (holes, n = 5570) and configs[1] (stripes, n = 11977)."""
    if name not in n2ll_cases:
        pytest.skip("golden value not generated")
    c = n2ll_cases[name]
    locs, X, z = case_design(c, datasets)
    with cb.DenseLikelihood(locs, X, z) as ctx:
        got = _values(c, locs, X, z, ctx=ctx)
    errs = {k: abs(got[k] - c["values"][k]) / abs(c["values"][k]) for k in got}
    print(name, {k: "%.2e" % v for k, v in errs.items()})
    assert all(v < TOL for v in errs.values()), errs


@pytest.mark.parametrize("n", [100, 128, 517, 1290])
def test_factor_reconstructs_sigma(n, datasets):
    H = datasets["holes_training"]
    X = cb.getScale(np.column_stack([np.ones(n), H[:n, 2], H[:n, 3]]))["std.covs"]
    locs = H[:n, :2]
    th = {"std.dev": np.array([0.2, 0.15, 0.1]), "scale": np.array([-1.6, 0.2, -0.15]),
          "aniso": np.array([0.1, 0.2, -0.1]), "tilt": np.array([0.3, -0.2, 0.1]),
          "smooth": np.array([0.2, 0.3, -0.2]), "nugget": np.array([-4, 0.1, 0.1])}
    with cb.DenseLikelihood(locs, X, H[:n, 4]) as ctx:
        ctx.factor(th, [0.5, 2.5])
        L, perm = ctx.get_factor()
    assert sorted(perm.tolist()) == list(range(n))
    S = cov.cov_rns(th, locs, X, [0.5, 2.5])[np.ix_(perm, perm)]
    assert np.all(np.triu(L, 1) == 0)
    resid = np.abs(L @ L.T - S).max() / np.abs(S).max()
    assert resid < 1e-13, resid
    ref = np.linalg.cholesky(S)
    assert relerr(np.diag(L), np.diag(ref)) < 1e-10


def test_permutation_invariance_at_scale():
    """Size-independent property at a size the oracle would need minutes for: the objective does
    not depend on the order the sites are given in, and duplicated evaluation is bit-stable."""
    rng = np.random.default_rng(20261018)
    n, p = 6000, 5
    locs = rng.uniform(-1, 1, (n, 2))
    c1, c2 = (locs[:, 0] + 1) / 2, (locs[:, 1] + 1) / 2
    X = cb.getScale(np.column_stack([np.ones(n), c1, c2, c1 * c2,
                                     0.5 + 0.5 * np.sin(np.pi * locs[:, 0]) * np.cos(np.pi * locs[:, 1])]))["std.covs"]
    z = rng.standard_normal(n)
    tl = {"mean": np.zeros(p), "std.dev": np.array([0.2, 0.15, 0.10, -0.05, 0.05]),
          "scale": np.array([-1.6, 0.2, -0.15, 0.1, -0.1]), "aniso": np.array([0.1, 0.2, -0.1, 0.05, 0]),
          "tilt": np.array([0.3, -0.2, 0.1, 0.1, -0.1]), "smooth": np.array([0.2, 0.3, -0.2, 0.1, 0.1]),
          "nugget": np.array([-4, 0.1, 0.1, 0, 0])}
    vals = []
    for perm in (np.arange(n), rng.permutation(n)):
        with cb.DenseLikelihood(locs[perm], X[perm], z[perm]) as ctx:
            t = ctx.terms(_lib.ML, tl, [0.5, 2.5], tl["mean"])
            vals.append(2 * t["logdet"] + t["quad"][0])
    assert abs(vals[0] - vals[1]) < 1e-10 * abs(vals[0]), vals


def test_dmma_trailing_update_kernel_runs_at_speed():
    ms = _lib.ctypes.c_double()
    n, k = 8192, 512
    _lib.check(_lib.lib().cocons_bench_syrk(0, n, k, 3, _lib.ctypes.byref(ms)))
    flops = (n / 128) * (n / 128 + 1) / 2 * 2 * 128 * 128 * k
    tflops = flops / (ms.value * 1e-3) / 1e12
    print("SYRK n=%d k=%d: %.3f ms, %.1f TFLOP/s" % (n, k, ms.value, tflops))
    assert tflops > 5.0


def test_pooled_evaluations_are_bit_identical_to_single_context(datasets):
    """Several evaluations submitted from host threads (DenseLikelihoodPool, the batched finite-difference
    gradient of cocoOptim / getHessian) must give exactly the values a single context gives, although their
    kernel chains overlap on the device (round 1 had to serialise them; see tests/test_gpu_repro.py)."""
    H = datasets["holes_training"]
    n = 3000
    X = cb.getScale(np.column_stack([np.ones(n), H[:n, 2], H[:n, 3]]))["std.covs"]
    base = {"mean": np.zeros(3), "std.dev": np.array([0.2, 0.15, 0.1]), "scale": np.array([-1.6, 0.2, -0.15]),
            "aniso": np.array([0.1, 0.2, -0.1]), "tilt": np.array([0.3, -0.2, 0.1]),
            "smooth": np.array([0.2, 0.3, -0.2]), "nugget": np.array([-4, 0.1, 0.1])}
    pts = []
    for k in range(16):
        t = {a: v.copy() for a, v in base.items()}
        t["scale"][0] += 1e-4 * k
        pts.append(t)

    def f(ctx, t):
        r = ctx.terms(_lib.ML, t, [0.5, 2.5], t["mean"])
        return r["logdet"], float(r["quad"][0])

    with cb.DenseLikelihood(H[:n, :2], X, H[:n, 4]) as ctx:
        ref = [f(ctx, t) for t in pts]
    with cb.DenseLikelihoodPool(H[:n, :2], X, H[:n, 4], size=4) as pool:
        for _ in range(2):
            assert pool.map(f, pts) == ref
