"""GPU parity of the sparse (tapered) model, through the C ABI:
  * cov_rns_taper / cov_rns_taper_pred entries against the goldens produced by the reference's own compiled
    src/cocons_taper.cpp and against the oracle on fresh seeded inputs - every entry within 1e-12 relative;
  * GetNeg2loglikelihoodTaper / ...TaperProfile and the sparse cocoPredict against the committed values
    (reference entries + LAPACK on the dense expansion in place of spam's sparse Cholesky) - 1e-8 relative;
  * size-independent properties at sizes no CPU oracle reaches in seconds."""
import numpy as np
import pytest

import cocons_b200 as cb
from conftest import relerr, theta_dict
from oracle import cov, rmirror

pytestmark = pytest.mark.gpu
TOL = 1e-12
PP = {"mean": np.ones(3, dtype=bool), "std.dev": np.ones(3, dtype=bool), "scale": np.ones(3, dtype=bool), "aniso": 0.0,
      "tilt": 0.0, "smooth": np.ones(3, dtype=bool), "nugget": np.ones(3, dtype=bool)}


def _gpu_entries(case):
    th = theta_dict(case["theta6"])
    if "locs_pred" in case:
        return cb.cov_rns_taper_pred(th, case["locs"], case["locs_pred"], case["X"], case["X_pred"],
                                     case["colindices"], case["rowpointers"], case["limits"])
    return cb.cov_rns_taper(th, case["locs"], case["X"], case["colindices"], case["rowpointers"], case["limits"])


def test_golden_entries(taper_cases):
    report = {}
    for name, case in taper_cases.items():
        if name == "obj":
            continue
        got = _gpu_entries(case)
        assert got.shape == case["out"].shape
        report[name] = relerr(got, case["out"])
    bad = {k: v for k, v in report.items() if not v < TOL}
    assert not bad, "relative error above 1e-12: %s (all: %s)" % (bad, report)
    print("max rel err per taper golden case:", {k: "%.1e" % v for k, v in report.items()})


def test_quirks_survive_on_the_device(taper_cases):
    c = taper_cases["taper_degenerate_nu1"]
    got = _gpu_entries(c)
    rp = c["rowpointers"].astype(np.int64)
    for i in (0, 17, 499):
        row = got[rp[i] - 1:rp[i + 1] - 1]
        assert np.all(row == row[0])
    c = taper_cases["taper_pred_general"]
    got = _gpu_entries(c)
    rp, ci = c["rowpointers"].astype(np.int64), c["colindices"].astype(np.int64)
    k = rp[3] - 1 + np.flatnonzero(ci[rp[3] - 1:rp[4] - 1] == 11)[0]  # prediction site 3 sits on training site 10
    assert abs(got[k] - c["out"][k]) <= 1e-15 * c["out"][k]


@pytest.mark.parametrize("n,p,delta,seed", [(1, 1, 0.5, 0), (2, 2, 3.0, 1), (129, 3, 0.4, 2), (1500, 4, 0.15, 3)])
def test_entries_against_oracle_on_seeded_inputs(n, p, delta, seed):
    rng = np.random.default_rng(seed)
    locs = rng.uniform(-1, 1, (n, 2))
    X = np.column_stack([np.ones(n), rng.standard_normal((n, p - 1))])
    th = {k: 0.25 * rng.standard_normal(p) for k in cov.ASPECTS}
    th["scale"][0], th["nugget"][0] = -2.0, -3.0
    sp = cb.nearest_dist(locs, delta=delta)
    for lim in ([0.5, 2.5], [0.3, 0.9]):
        ref = cov.cov_rns_taper(th, locs, X, sp.colindices, sp.rowpointers, lim)
        got = cb.cov_rns_taper(th, locs, X, sp.colindices, sp.rowpointers, lim)
        assert relerr(got, ref) < TOL, (lim, relerr(got, ref))
    th0 = dict(th, smooth=np.zeros(p))
    for lim in ([0.5, 0.5], [1.5, 1.5], [2.5, 2.5], [1.0, 1.0]):
        ref = cov.cov_rns_taper(th0, locs, X, sp.colindices, sp.rowpointers, lim)
        assert relerr(cb.cov_rns_taper(th0, locs, X, sp.colindices, sp.rowpointers, lim), ref) < TOL, lim
    m = 37
    lp = rng.uniform(-1, 1, (m, 2))
    lp[5] = locs[0]
    Xp = np.column_stack([np.ones(m), rng.standard_normal((m, p - 1))])
    spp = cb.nearest_dist(lp, locs, delta=delta)
    ref = cov.cov_rns_taper_pred(th, locs, lp, X, Xp, spp.colindices, spp.rowpointers, [0.5, 2.5])
    got = cb.cov_rns_taper_pred(th, locs, lp, X, Xp, spp.colindices, spp.rowpointers, [0.5, 2.5])
    assert relerr(got, ref) < TOL


def _obj_setup(taper_cases):
    o, c = taper_cases["obj"], taper_cases["taper_general"]
    delta = float(o["delta"])
    ref_taper = cb.cov_wend1(cb.nearest_dist(c["locs"], delta=delta), (delta, 1))
    return o, c, delta, ref_taper


def test_objectives_match_goldens(taper_cases):
    o, c, delta, ref_taper = _obj_setup(taper_cases)
    n, lim, z = len(c["locs"]), [0.5, 2.5], o["z"]
    v = cb.GetNeg2loglikelihoodTaper(o["theta"], PP, ref_taper, c["locs"], c["X"], lim, None, z, n, (0, 0, 0))
    assert abs(v - float(o["ml"])) < 1e-9 * abs(v), (v, float(o["ml"]))
    v = cb.GetNeg2loglikelihoodTaper(o["theta"], PP, ref_taper, c["locs"], c["X"], lim, None, z, n, (0.05, 0.02, 0.3))
    assert abs(v - float(o["ml_pen"])) < 1e-9 * abs(v)
    ppp = dict(PP)
    ppp["std.dev"] = np.array([False, True, True])
    with cb.DenseLikelihood(c["locs"], c["X"], z) as ctx:  # resident context, several evaluations
        ctx.set_taper(ref_taper)
        for _ in range(2):
            v = cb.GetNeg2loglikelihoodTaperProfile(o["theta_profile"], ppp, ref_taper, None, None, lim, None, None, n,
                                                    (0, 0, 0), ctx=ctx)
            assert abs(v - float(o["profile"])) < 1e-9 * abs(v), (v, float(o["profile"]))
            v = cb.GetNeg2loglikelihoodTaper(o["theta"], PP, ref_taper, None, None, lim, None, None, n, (0, 0, 0), ctx=ctx)
            assert abs(v - float(o["ml"])) < 1e-9 * abs(v)
        # the dense objective on the same context is unaffected by the attached taper
        dense = cb.GetNeg2loglikelihood(o["theta"], PP, None, None, lim, None, n, (0, 0, 0), ctx=ctx)
        want = rmirror.neg2loglik(o["theta"], PP, c["locs"], c["X"], lim, z, n, (0, 0, 0))
        assert abs(dense - want) < 1e-9 * abs(want)


def test_not_positive_definite_and_malformed_patterns(taper_cases):
    o, c, delta, ref_taper = _obj_setup(taper_cases)
    n, lim = len(c["locs"]), [0.5, 2.5]
    bad = cb.spam(o["notpd_taper"], o["notpd_colindices"], o["notpd_rowpointers"], (n, n))
    v = cb.GetNeg2loglikelihoodTaper(o["theta"], PP, bad, c["locs"], c["X"], lim, None, o["z"], n, (0, 0, 0))
    assert v == 1e6 == float(o["notpd"])
    with pytest.raises(ArithmeticError, match="Cholesky error"):
        cb.GetNeg2loglikelihoodTaper(o["theta"], PP, bad, c["locs"], c["X"], lim, None, o["z"], n, (0, 0, 0), safe=False)
    th = theta_dict(c["theta6"])
    rp = c["rowpointers"].copy()
    rp[-1] += 1
    with pytest.raises(cb.CoconsError, match="malformed pattern"):
        cb.cov_rns_taper(th, c["locs"], c["X"], c["colindices"], rp, lim)
    ci = c["colindices"].copy()
    ci[7] = n + 1
    with pytest.raises(cb.CoconsError, match="outside"):
        cb.cov_rns_taper(th, c["locs"], c["X"], ci, c["rowpointers"], lim)
    with cb.DenseLikelihood(c["locs"], c["X"], o["z"]) as ctx:
        with pytest.raises(cb.CoconsError, match="set_taper"):
            ctx.terms_taper(th, lim, np.zeros(3))
        ctx.set_taper(ref_taper)
        ctx.factor_taper(th, lim)
        with pytest.raises(cb.CoconsError, match="tapered model"):
            ctx.predict(c["locs"][:5], c["X"][:5], np.zeros(n))


def _sparse_coco(taper_cases, datasets):
    o, c, delta, _ = _obj_setup(taper_cases)
    H = datasets["holes_training"]
    n = len(c["locs"])
    data = {"cov_x": H[:n, 2], "cov_y": H[:n, 3]}
    f = "~ 1 + cov_x + cov_y"
    obj = cb.coco("sparse", data, c["locs"], o["z"], {"mean": f, "std.dev": f, "scale": f, "smooth": f, "nugget": f},
                  info={"smooth.limits": [0.5, 2.5], "taper": cb.cov_wend1, "delta": delta})
    sc = cb.getScale(cb.getDesignMatrix(obj.model_list, obj.data)["model.matrix"])
    obj.output = {"par": o["theta"]}
    obj.info.update({"mean.vector": sc["mean.vector"], "sd.vector": sc["sd.vector"]})
    return obj, o, c


def test_sparse_predict_matches_golden(taper_cases, datasets):
    obj, o, c = _sparse_coco(taper_cases, datasets)
    p = taper_cases["taper_pred_general"]
    HT = datasets["holes_test"]
    m = len(p["locs_pred"])
    new = {"cov_x": HT[:m, 2], "cov_y": HT[:m, 3]}
    out = cb.cocoPredict(obj, new, p["locs_pred"], type="pred")
    assert relerr(out["stochastic"], o["pred_stochastic"]) < 1e-8
    assert relerr(out["sd.pred"], o["pred_sd"]) < 1e-8
    mean_only = cb.cocoPredict(obj, new, p["locs_pred"], type="mean")
    assert np.array_equal(mean_only["stochastic"], out["stochastic"]) and "sd.pred" not in mean_only


def test_sparse_getcovmatrix_factor_and_sim(taper_cases, datasets):
    obj, o, c = _sparse_coco(taper_cases, datasets)
    S = cb.getCovMatrix(obj)
    assert isinstance(S, cb.spam)
    tl = cb.getModelLists(o["theta"], PP, "diff")
    d, ci, rp = rmirror.nearest_dist(c["locs"], delta=float(o["delta"]))
    want = rmirror.tapered_matrix(tl, rmirror.cov_wend1(d, (float(o["delta"]), 1)), ci, rp, c["locs"], c["X"], [0.5, 2.5])
    dense = S.toarray()
    assert relerr(np.tril(dense), np.tril(want)) < TOL
    # the factor kept on the device is a Cholesky factor of that matrix (in the context's site order)
    with cb.DenseLikelihood(c["locs"], c["X"], o["z"]) as ctx:
        ctx.set_taper(cb.cov_wend1(cb.nearest_dist(c["locs"], delta=float(o["delta"])), (float(o["delta"]), 1)))
        ctx.factor_taper(tl, [0.5, 2.5])
        L, perm = ctx.get_factor()
        assert np.allclose(L @ L.T, want[np.ix_(perm, perm)], rtol=0, atol=1e-12 * np.abs(want).max())
        eps = np.random.default_rng(0).standard_normal((len(perm), 2))
        draws = ctx.sim(eps)
        assert np.allclose(draws[perm], L @ eps[perm], rtol=1e-11, atol=1e-13)
    sim = cb.cocoSim(obj, n=2, seed=3)
    assert sim.shape == (len(perm), 2) and np.all(np.isfinite(sim))


def test_full_pattern_with_a_flat_taper_is_the_dense_objective():
    """delta so large that the Wendland taper rounds to 1 on every pair and the pattern is full: the tapered
    objective must equal the dense one of the same isotropic model (two different assembly kernels)."""
    rng = np.random.default_rng(4)
    n = 1800
    locs = rng.uniform(-1, 1, (n, 2))
    X = cb.getScale(np.column_stack([np.ones(n), locs[:, 0], locs[:, 1]]))["std.covs"]
    z = rng.standard_normal(n)
    theta = np.array([0.1, 0.3, -0.2, -1.4, 0.35, -0.05, 1.8, -0.05, 0.25, 0.2, 0.3, -0.2, -4, 0.1, 0.1])
    ref_taper = cb.cov_wend1(cb.nearest_dist(locs, delta=1e9), (1e9, 1))
    assert ref_taper.entries.shape[0] == n * n and np.all(np.abs(ref_taper.entries - 1.0) < 1e-15)
    ref_taper.entries[:] = 1.0
    a = cb.GetNeg2loglikelihoodTaper(theta, PP, ref_taper, locs, X, [0.5, 2.5], None, z, n, (0, 0, 0))
    b = cb.GetNeg2loglikelihood(theta, PP, locs, X, [0.5, 2.5], z, n, (0, 0, 0))
    assert abs(a - b) < 1e-9 * abs(b), (a, b)


def test_diagonal_pattern_at_scale():
    """Size-independent property at n = 30 000 (a 7.2 GB matrix): with a taper radius below the smallest
    site spacing only the diagonal is stored, Sigma = diag(E(std.dev) + E(nugget)) and the objective has a
    closed form the host evaluates in numpy."""
    import bench
    n = 30000
    locs, X, z = bench.synthetic(n)
    tl = dict(bench.THETA)
    sp = cb.nearest_dist(locs, delta=1e-7)
    assert sp.entries.shape[0] == n
    with cb.DenseLikelihood(locs, X, z) as ctx:
        ctx.set_taper(cb.cov_wend1(sp, (1e-7, 1)))
        t = ctx.terms_taper(tl, bench.LIMITS, tl["mean"])
    dv = 1 / np.exp(-(X @ tl["std.dev"])) + 1 / np.exp(-(X @ tl["nugget"]))
    resid = z.reshape(-1) - X @ tl["mean"]
    assert abs(t["logdet"] - 0.5 * np.sum(np.log(dv))) < 1e-11 * abs(0.5 * np.sum(np.log(dv)))
    assert abs(t["quad"][0] - np.sum(resid ** 2 / dv)) < 1e-11 * np.sum(resid ** 2 / dv)


def test_sparse_smoke_sequence_of_the_reference_test_script(datasets):
    """tests/coco_test.R:130-200 in miniature: sparse coco object (nu = 1.5, Wendland-1 taper), "ml" and "pml"
    fits, the tapered covariance matrix is positive definite, predictions hold no NaN; plus the identity
    that ties the pml recovery (R/optim.R:590-662) together: at the recovered parameters the full tapered
    objective equals the profiled one at the optimum."""
    H, T = datasets["holes_training"][:100], datasets["holes_test"][:60]
    data = {"x": H[:, 0], "y": H[:, 1], "cov_x": H[:, 2], "cov_y": H[:, 3]}
    ml = {"mean": 0, "std.dev": "~ 1 + cov_x", "scale": "~ 1", "aniso": 0, "tilt": 0, "smooth": 1.5,
          "nugget": -np.inf}
    info = {"taper": cb.cov_wend1, "delta": 0.4}
    obj = cb.coco("sparse", data, H[:, :2], H[:, 4], ml, info=dict(info))
    assert 0 < cb.getDensityFromDelta(obj, 0.2) < cb.getDensityFromDelta(obj, 0.4) < 1
    with pytest.raises(ValueError, match="only for sparse"):
        cb.getDensityFromDelta(cb.coco("dense", data, H[:, :2], H[:, 4], ml), 0.2)
    bounds = {"theta_init": np.array([0.0, 0.0, -1.0]), "theta_lower": np.array([-4.0, -2.0, -6.0]),
              "theta_upper": np.array([4.0, 2.0, 4.0])}
    fit = cb.cocoOptim(obj, bounds, optim_control={"maxiter": 25})
    assert fit.output["value"] < 1e6 and np.all(np.isfinite(fit.output["par"]))
    cmat = cb.getCovMatrix(fit).toarray()
    cmat = np.tril(cmat) + np.tril(cmat, -1).T
    assert cmat.shape == (100, 100) and np.all(np.linalg.eigvalsh(cmat) > 0)
    newdata = {"x": T[:, 0], "y": T[:, 1], "cov_x": T[:, 2], "cov_y": T[:, 3]}
    pr = cb.cocoPredict(fit, newdata, T[:, :2], type="pred")
    assert not np.any(np.isnan(pr["stochastic"])) and np.all(pr["sd.pred"] > 0)
    # the fit did not end above its starting value
    dm = cb.getDesignMatrix(obj.model_list, obj.data)
    X = cb.getScale(dm["model.matrix"])["std.covs"]
    ref_taper = cb.cov_wend1(cb.nearest_dist(H[:, :2], delta=0.4), (0.4, 1))
    f0 = cb.GetNeg2loglikelihoodTaper(bounds["theta_init"], dm["par.pos"], ref_taper, H[:, :2], X, [1.5, 1.5], None,
                                      H[:, 4], 100, (0, 0, 0))
    assert fit.output["value"] <= f0 + 1e-9 * abs(f0)
    # pml: the global variance is profiled out and recovered
    obj2 = cb.coco("sparse", data, H[:, :2], H[:, 4], ml, info=dict(info))
    fit2 = cb.cocoOptim(obj2, bounds, optim_type="pml", optim_control={"maxiter": 25})
    assert fit2.output["par"].shape == (3,) and np.all(np.isfinite(fit2.output["par"]))
    full = cb.GetNeg2loglikelihoodTaper(fit2.output["par"], dm["par.pos"], ref_taper, H[:, :2], X, [1.5, 1.5], None,
                                        H[:, 4], 100, (0, 0, 0))
    assert abs(full - fit2.output["value"]) < 1e-8 * abs(full), (full, fit2.output["value"])
    assert fit2.output["value"] <= fit.output["value"] + 1e-6 * abs(fit.output["value"])
