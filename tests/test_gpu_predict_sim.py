"""GPU parity of cocoPredict / cocoSim / getCovMatrix / cocoOptim on the kept device factor,
against the literal restatement of R/predict.R, R/sim.R (oracle/rmirror.py), and the smoke
properties tests/coco_test.R asserts."""
import numpy as np
import pytest

import cocons_b200 as cb
from conftest import relerr
from oracle import rmirror

pytestmark = pytest.mark.gpu

TL = {"mean": np.array([0.1, 0.3, -0.2]), "std.dev": np.array([0.2, 0.15, 0.1]),
      "scale": np.array([-1.6, 0.2, -0.15]), "aniso": np.array([0.1, 0.2, -0.1]),
      "tilt": np.array([0.3, -0.2, 0.1]), "smooth": np.array([0.2, 0.3, -0.2]), "nugget": np.array([-4, 0.1, 0.1])}
LIM = [0.5, 2.5]


def _setup(datasets, n, m):
    H, T = datasets["holes_training"], datasets["holes_test"]
    sc = cb.getScale(np.column_stack([np.ones(n), H[:n, 2], H[:n, 3]]))
    Xp = cb.getScale(np.column_stack([np.ones(m), T[:m, 2], T[:m, 3]]), sc["mean.vector"], sc["sd.vector"])["std.covs"]
    return H[:n, :2], sc["std.covs"], H[:n, 4], T[:m, :2].copy(), Xp


@pytest.mark.parametrize("n,m", [(600, 150), (1000, 430), (5570, 430)])  # the last: BASELINE configs[0] at full size
def test_predict_matches_reference_route(datasets, n, m):
    locs, X, z, lp, Xp = _setup(datasets, n, m)
    lp[7] = locs[11]  # a prediction site sitting on a training site
    ref = rmirror.predict(TL, locs, lp, X, Xp, LIM, z, type="pred")
    resid = z - X @ TL["mean"]
    with cb.DenseLikelihood(locs, X, z) as ctx:
        ctx.factor(TL, LIM)
        sto, expl = ctx.predict(lp, Xp, resid)
    assert relerr(sto, ref["stochastic"]) < 1e-8
    u = np.exp(Xp @ TL["std.dev"]) + np.exp(Xp @ TL["nugget"]) - expl
    neg = u < 1e-10
    u[neg] = np.abs(u[neg])
    assert np.max(np.abs(np.sqrt(u) - ref["sd.pred"]) / ref["sd.pred"]) < 1e-7


def test_predict_after_fixed_smoothness_factor(datasets):
    """cov_rns takes the nu = 1.5 closed form, cov_rns_pred the Bessel branch."""
    locs, X, z, lp, Xp = _setup(datasets, 500, 100)
    tl = dict(TL, aniso=np.zeros(3), tilt=np.zeros(3), smooth=np.zeros(3), nugget=np.array([-3.0, 0, 0]))
    ref = rmirror.predict(tl, locs, lp, X, Xp, [1.5, 1.5], z, type="mean")
    with cb.DenseLikelihood(locs, X, z) as ctx:
        ctx.factor(tl, [1.5, 1.5])
        sto, _ = ctx.predict(lp, Xp, z - X @ tl["mean"], want_explained=False)
    assert relerr(sto, ref["stochastic"]) < 1e-8


@pytest.mark.parametrize("typ", ["diff", "classic"])
def test_marginal_simulation_is_L_times_eps(datasets, typ):
    locs, X, z, _, _ = _setup(datasets, 700, 10)
    tl = dict(TL)
    if typ == "classic":
        tl["smooth"] = np.array([0.1, 0.2, -0.1])
    eps = np.random.default_rng(4).standard_normal((700, 3))
    ref = rmirror.sim_marginal(tl, locs, X, LIM, eps, type=typ)
    with cb.DenseLikelihood(locs, X, z) as ctx:
        ctx.factor(tl, LIM, type=typ)
        draws = ctx.sim(eps) + (X @ tl["mean"])[:, None]
    # same Sigma, same eps; the site ordering inside the context differs from LAPACK's, so the
    # draws agree in distribution, not entry-wise: compare through Sigma^-1/2-free invariants
    S = (rmirror._cov.cov_rns_classic(tl, locs, X) if typ == "classic" else rmirror._cov.cov_rns(tl, locs, X, LIM))
    d_ref = ref - (X @ tl["mean"])[:, None]
    d_got = draws - (X @ tl["mean"])[:, None]
    R = rmirror.r_chol(S)
    # |L^-1 d|^2 must equal |eps|^2 column by column for any valid square root of Sigma
    for dd in (d_ref, d_got):
        y = rmirror._fwd(R, dd)
        assert np.allclose((y * y).sum(axis=0), (eps * eps).sum(axis=0), rtol=1e-9)


def test_marginal_simulation_entrywise_in_sorted_order(datasets):
    """With eps permuted the same way the context permutes the sites, L eps is entry-wise checkable."""
    locs, X, z, _, _ = _setup(datasets, 400, 10)
    eps = np.random.default_rng(9).standard_normal((400, 2))
    with cb.DenseLikelihood(locs, X, z) as ctx:
        ctx.factor(TL, LIM)
        L, perm = ctx.get_factor()
        got = ctx.sim(eps)
    want = np.empty_like(got)
    want[perm] = L @ eps[perm]
    assert relerr(got, want) < 1e-11


def test_conditional_simulation(datasets):
    locs, X, z, lp, Xp = _setup(datasets, 500, 128)
    eps = np.random.default_rng(5).standard_normal((128, 2))
    with cb.DenseLikelihood(locs, X, z) as ctx:
        ctx.factor(TL, LIM)
        got = ctx.sim_cond(lp, Xp, eps)
    S = rmirror._cov.cov_rns(TL, locs, X, LIM)
    C = rmirror._cov.cov_rns_pred(TL, locs, lp, X, Xp, LIM)
    Su = rmirror._cov.cov_rns(TL, lp, Xp, LIM)
    schur = Su - C @ np.linalg.solve(S, C.T)
    want = np.linalg.cholesky(schur) @ eps  # t(t(eps) %*% chol(.)) with chol upper
    assert relerr(got, want) < 1e-7


def test_smoke_sequence_of_the_reference_test_script(datasets):
    """tests/coco_test.R:16-46 in miniature: fit 50 holes points (nu = 1.5), covariance matrix is
    50 x 50 and positive definite, predictions have systematic part exactly 0 and no NaN."""
    H, T = datasets["holes_training"][:50], datasets["holes_test"][:50]
    data = {"x": H[:, 0], "y": H[:, 1], "cov_x": H[:, 2], "cov_y": H[:, 3]}
    ml = {"mean": 0, "std.dev": "~ 1", "scale": "~ 1", "aniso": 0, "tilt": 0, "smooth": 1.5, "nugget": -np.inf}
    obj = cb.coco("dense", data, H[:, :2], H[:, 4], ml)
    bounds = {"theta_init": np.array([0.0, -1.0]), "theta_lower": np.array([-4.0, -6.0]),
              "theta_upper": np.array([4.0, 4.0])}
    fit = cb.cocoOptim(obj, bounds, optim_control={"maxiter": 30})
    assert np.isfinite(fit.output["value"]) and fit.output["value"] < 1e6
    cmat = cb.getCovMatrix(fit)
    assert cmat.shape == (50, 50) and np.all(np.linalg.eigvalsh(cmat) > 0)
    newdata = {"x": T[:, 0], "y": T[:, 1], "cov_x": T[:, 2], "cov_y": T[:, 3]}
    pr = cb.cocoPredict(fit, newdata, T[:, :2], type="pred")
    assert np.all(pr["systematic"] == 0) and not np.any(np.isnan(pr["stochastic"]))
    assert np.all(pr["sd.pred"] > 0)
    sims = cb.cocoSim(fit, n=2, seed=1)
    assert sims.shape == (50, 2) and np.all(np.isfinite(sims))


def test_pml_fit_recovers_betas(datasets):
    """tests/coco_test.R:48-72 in miniature (pml with a mean model)."""
    H = datasets["holes_training"][:60]
    data = {"x": H[:, 0], "y": H[:, 1], "cov_x": H[:, 2], "cov_y": H[:, 3]}
    ml = {"mean": "~ 1 + cov_x + cov_y", "std.dev": "~ 1", "scale": "~ 1", "aniso": 0, "tilt": 0, "smooth": 1.5,
          "nugget": -np.inf}
    obj = cb.coco("dense", data, H[:, :2], H[:, 4], ml)
    bounds = {"theta_init": np.array([0, 0, 0, 0.0, -1.0]), "theta_lower": np.array([-9, -9, -9, -4.0, -6.0]),
              "theta_upper": np.array([9, 9, 9, 4.0, 4.0])}
    fit = cb.cocoOptim(obj, bounds, optim_type="pml", optim_control={"maxiter": 20})
    assert fit.output["par"].shape == (5,) and np.all(np.isfinite(fit.output["par"]))
    # the recovered betas are the GLS estimate under the fitted covariance
    S = cb.getCovMatrix(fit)
    X = cb.getScale(fit)["std.covs"]
    Si = np.linalg.inv(S)
    gls = np.linalg.solve(X.T @ Si @ X, X.T @ Si @ H[:, 4])
    assert np.allclose(fit.output["par"][:3], gls, rtol=1e-7, atol=1e-9)


def test_hessian_mirrors_the_reference_scheme(datasets):
    """getHessian (R/getFunctions.R:925-1034): same forward-difference formula, evaluated on the GPU pool;
    compared with the identical formula driven by the CPU oracle's objective."""
    H0 = datasets["holes_training"][:80]
    data = {"x": H0[:, 0], "y": H0[:, 1], "cov_x": H0[:, 2], "cov_y": H0[:, 3]}
    ml = {"mean": 0, "std.dev": "~ 1 + cov_x", "scale": "~ 1", "aniso": 0, "tilt": 0, "smooth": 1.5,
          "nugget": -np.inf}
    obj = cb.coco("dense", data, H0[:, :2], H0[:, 4], ml)
    dm = cb.getDesignMatrix(obj.model_list, obj.data)
    sc = cb.getScale(dm["model.matrix"])
    par = np.array([0.3, 0.1, -1.2])
    lam = (0.0, 0.0, 0.0)
    f = lambda th: rmirror.neg2loglik(th, dm["par.pos"], H0[:, :2], sc["std.covs"], [1.5, 1.5], H0[:, 4], 80, lam)  # noqa: E731
    obj.output = {"par": par, "value": f(par)}
    obj.info.update({"mean.vector": sc["mean.vector"], "sd.vector": sc["sd.vector"], "optim.type": "ml"})
    Hg = cb.getHessian(obj)
    eps = np.finfo(float).eps ** 0.25
    p = 3
    Hr = np.zeros((p, p))
    for j in range(p):
        for i in range(j, p):
            e1, e2 = np.eye(p)[j] * eps, np.eye(p)[i] * eps
            Hr[j, i] = 0.5 * (f(par + e1 + e2) - f(par + e1) - f(par + e2) + f(par)) / eps ** 2
    Hr = Hr + Hr.T
    Hr[np.diag_indices(p)] /= 2
    assert np.array_equal(Hg, Hg.T)
    # both are differences of O(1e2) values divided by eps^2 = 1.5e-8: agreement to ~1e-5 of |f|/eps^2 scale
    assert np.max(np.abs(Hg - Hr)) < 1e-3 * max(1.0, np.max(np.abs(Hr)))


def test_penalised_two_step_fit_prunes_and_refits(datasets):
    """R/optim.R:127-230: with lambda.Sigma / lambda.betas > 0 the 'ml' fit is penalised, small coefficients are
    pruned at sparse.point and the pruned model is refitted with lambda = (0, 0, lambda.reg).  A covariate that
    is pure noise for the standard deviation must leave the model under a strong penalty, and the second fit
    must be a plain ML fit of the model that is left."""
    H = datasets["holes_training"][:120]
    rng = np.random.default_rng(5)
    noise = rng.standard_normal(120)
    data = {"x": H[:, 0], "y": H[:, 1], "cov_x": H[:, 2], "noise": noise}
    ml = {"mean": 0, "std.dev": "~ 1 + noise", "scale": "~ 1", "aniso": 0, "tilt": 0, "smooth": 1.5, "nugget": -np.inf}
    obj = cb.coco("dense", data, H[:, :2], H[:, 4], ml, info={"lambda.Sigma": 5.0, "sparse.point": 1e-3})
    bounds = {"theta_init": np.array([0.0, 0.0, -1.0]), "theta_lower": np.array([-4.0, -2.0, -6.0]),
              "theta_upper": np.array([4.0, 2.0, 4.0])}
    fit = cb.cocoOptim(obj, bounds, optim_control={"maxiter": 60})
    assert fit.model_list["std.dev"].replace(" ", "") == "~1", fit.model_list
    assert len(fit.output["par"]) == 2 and len(fit.info["boundaries"]["theta_init"]) == 2
    assert abs(fit.info["first.step"]["par"][1]) <= 1e-3
    # the refit is the unpenalised objective of the pruned model at its optimum
    dm = cb.getDesignMatrix(fit.model_list, fit.data)
    X = cb.getScale(dm["model.matrix"])["std.covs"]
    v = cb.GetNeg2loglikelihood(fit.output["par"], dm["par.pos"], fit.locs, X, fit.info["smooth.limits"], fit.z,
                                120, (0.0, 0.0, 0.0))
    assert abs(v - fit.output["value"]) <= 1e-9 * abs(v)
