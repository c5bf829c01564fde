"""Host-side mirror of the reference's R interface for the dense-likelihood path.

The reference's host language is R, which is not installed here; the R glue a
maintainer would drop into the package is in cocons_b200/rglue/ (see
INTEGRATION.md).  This module offers the same functions, names and argument
order in Python on top of the same C ABI so that the path can be driven and
tested without R.  Everything numeric of order n^2 or n^3 happens in
libcocons_b200.so on the GPU; what stays here is what stays in R in the
reference: parameter packing (getModelLists), standardisation (getScale), the
penalty (.cocons.getPen), the `safe`/1e6 logic and the n log(2 pi) constant.

Reference lines are cited per function.  R lists become dicts, formulas become
strings such as "~ 1 + cov_x + cov_y", data.frames become dicts of columns (or
pandas DataFrames).
"""
import numpy as np

from . import _lib
from ._lib import NotPositiveDefinite, CoconsError  # noqa: F401  (re-exported)

DICTIONARY = ("mean", "std.dev", "scale", "aniso", "tilt", "smooth", "nugget")  # R/profile.R:5-7


# --------------------------------------------------------------------------
# FFI-level functions (R/RcppExports.R:10-46)
# --------------------------------------------------------------------------
def sumsmoothlone(x, lambda_, alpha=1e6):
    """R/RcppExports.R:10-12 -> src/cocons_full.cpp:12-30."""
    x = np.ascontiguousarray(np.atleast_1d(np.asarray(x, dtype=np.float64)))
    return _lib.lib().cocons_sumsmoothlone(_lib.ptr(x), x.shape[0], float(lambda_), float(alpha))


def cov_rns(theta, locs, x_covariates, smooth_limits):
    """R/RcppExports.R:21-23 -> src/cocons_full.cpp:40-321.  Returns the n x n matrix."""
    locs, X = _lib.fmat(locs), _lib.fmat(x_covariates)
    n, p = X.shape
    th = _lib.pack_theta(theta, p)
    lim = np.ascontiguousarray(np.asarray(smooth_limits, dtype=np.float64))
    out = np.empty((n, n), order="F")
    _lib.check(_lib.lib().cocons_cov_rns(n, p, _lib.ptr(locs), _lib.ptr(X), _lib.ptr(th), _lib.ptr(lim),
                                         _lib.ptr(out)))
    return out


def cov_rns_pred(theta, locs, locs_pred, x_covariates, x_covariates_pred, smooth_limits):
    """R/RcppExports.R:34-36 -> src/cocons_full.cpp:334-471.  Returns m x n (prediction sites are rows)."""
    locs, lp = _lib.fmat(locs), _lib.fmat(locs_pred)
    X, Xp = _lib.fmat(x_covariates), _lib.fmat(x_covariates_pred)
    n, p = X.shape
    m = Xp.shape[0]
    th = _lib.pack_theta(theta, p)
    lim = np.ascontiguousarray(np.asarray(smooth_limits, dtype=np.float64))
    out = np.empty((m, n), order="F")
    _lib.check(_lib.lib().cocons_cov_rns_pred(n, m, p, _lib.ptr(locs), _lib.ptr(lp), _lib.ptr(X), _lib.ptr(Xp),
                                              _lib.ptr(th), _lib.ptr(lim), _lib.ptr(out)))
    return out


def cov_rns_classic(theta, locs, x_covariates):
    """R/RcppExports.R:44-46 -> src/cocons_full.cpp:480-594."""
    locs, X = _lib.fmat(locs), _lib.fmat(x_covariates)
    n, p = X.shape
    th = _lib.pack_theta(theta, p)
    out = np.empty((n, n), order="F")
    _lib.check(_lib.lib().cocons_cov_rns_classic(n, p, _lib.ptr(locs), _lib.ptr(X), _lib.ptr(th), _lib.ptr(out)))
    return out


# --------------------------------------------------------------------------
# parameter packing that stays on the host (R/getFunctions.R)
# --------------------------------------------------------------------------
def is_formula(x):
    """R/isFunctions.R:10-12."""
    return isinstance(x, str) and x.strip().startswith("~")


def _terms(formula):
    rhs = formula.strip()[1:]
    intercept, labels = True, []
    for tok in rhs.replace("-", "+-").split("+"):
        tok = tok.strip()
        if tok in ("", "1"):
            continue
        if tok in ("0", "-1"):
            intercept = False
            continue
        labels.append(tok)
    return intercept, labels


def _columns(data):
    if hasattr(data, "columns") and hasattr(data, "__getitem__") and not isinstance(data, dict):
        return {c: np.asarray(data[c], dtype=np.float64) for c in data.columns}
    return {k: np.asarray(v, dtype=np.float64) for k, v in data.items()}


def getDesignMatrix(model_list, data):
    """R/getFunctions.R:450-555 for main-effect formulas: the union model matrix (intercept +
    covariates in order of first appearance) and, per aspect, a logical index (free) or the
    fixed value."""
    cols = _columns(data)
    n = len(next(iter(cols.values())))
    formulas = [(k, v) for k, v in model_list.items() if is_formula(v)]
    if not formulas:
        raise ValueError("No formula detected")
    labels, any_intercept = [], False
    for _, f in formulas:
        ic, ls = _terms(f)
        any_intercept = any_intercept or ic
        for l in ls:
            if l not in labels:
                labels.append(l)
    names = (["(Intercept)"] if (any_intercept or not labels) else []) + labels
    mm = np.empty((n, len(names)), order="F")
    for j, nm in enumerate(names):
        mm[:, j] = 1.0 if nm == "(Intercept)" else cols[nm]
    par_pos = {}
    for k, v in model_list.items():
        if not is_formula(v):
            par_pos[k] = float(np.atleast_1d(v)[0])
            continue
        ic, ls = _terms(v)
        pos = np.array([nm in ls for nm in names], dtype=bool)
        if ic and names[0] == "(Intercept)":
            pos[0] = True
        par_pos[k] = pos
    return {"model.matrix": mm, "par.pos": par_pos, "colnames": names}


def getScale(x, mean_vector=None, sd_vector=None):
    """R/getFunctions.R:376-436: centre and scale columns 2..p (sd with n-1); column 1 untouched."""
    if isinstance(x, coco):
        x = getDesignMatrix(x.model_list, x.data)["model.matrix"]
    x = np.array(x, dtype=np.float64, order="F", copy=True)
    if mean_vector is None:
        mean_vector = x.mean(axis=0)
        mean_vector[0] = 0.0
    if sd_vector is None:
        sd_vector = x.std(axis=0, ddof=1)
        sd_vector[0] = 1.0
    for k in range(1, x.shape[1]):
        x[:, k] = (x[:, k] - mean_vector[k]) / sd_vector[k]
    return {"std.covs": x, "mean.vector": np.asarray(mean_vector), "sd.vector": np.asarray(sd_vector)}


def getModelLists(theta, par_pos, type="diff"):
    """R/getFunctions.R:570-616: theta vector -> one length-p vector per aspect; fixed aspects put
    their constant in slot 1; with type "diff", where std.dev and scale are both free at column k,
    std.dev_k = (a+b)/2 and scale_k = (a-b)/2."""
    theta = np.atleast_1d(np.asarray(theta, dtype=np.float64))
    free = {k: isinstance(v, np.ndarray) and v.dtype == bool for k, v in par_pos.items()}
    length = max((len(v) if free[k] else 1) for k, v in par_pos.items())
    pars, used = {}, 0
    for name, pos in par_pos.items():
        vec = np.zeros(length)
        if free[name]:
            k = int(pos.sum())
            vec[np.flatnonzero(pos)] = theta[used:used + k]
            used += k
        else:
            vec[0] = float(np.atleast_1d(pos)[0])
        pars[name] = vec
    if type == "classic":
        return pars
    out = {k: v.copy() for k, v in pars.items()}
    if free.get("std.dev") and free.get("scale"):
        both = par_pos["std.dev"] & par_pos["scale"]
        out["std.dev"][both] = (pars["std.dev"][both] + pars["scale"][both]) / 2
        out["scale"][both] = (pars["std.dev"][both] - pars["scale"][both]) / 2
    return out


def _getPen(n, lambda_, theta_list, smooth_limits):
    """.cocons.getPen, R/checkFunctions.R:474-492."""
    names = list(theta_list.keys())
    summ = lambda_[2] * np.exp(theta_list["scale"][0]) * np.sqrt(
        (smooth_limits[1] - smooth_limits[0]) / (1 + np.exp(-theta_list["smooth"][0])) + smooth_limits[0]
    ) + sumsmoothlone(theta_list[names[0]][1:], lambda_[1])
    for ii in range(1, 6):
        summ = summ + sumsmoothlone(theta_list[names[ii]][1:], lambda_[0])
    return 2 * n * summ


# --------------------------------------------------------------------------
# device-resident likelihood context
# --------------------------------------------------------------------------
class DenseLikelihood:
    """locs / x_covariates / z resident on one GPU; evaluates the three objectives and keeps the
    Cholesky factor for prediction and simulation.  Wraps cocons_ctx_* (include/cocons_b200.h)."""

    def __init__(self, locs, x_covariates, z, device=0, stream=None):
        self.locs, self.X = _lib.fmat(locs), _lib.fmat(x_covariates)
        self.n, self.p = self.X.shape
        self.z = _lib.fmat(z, rows=self.n)
        self.r = self.z.shape[1]
        self._h = _lib._vp()
        _lib.check(_lib.lib().cocons_ctx_create(int(device), self.n, self.p, self.r, _lib.ptr(self.locs),
                                                _lib.ptr(self.X), _lib.ptr(self.z), stream, self._h))
        self.q = 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().cocons_ctx_destroy(self._h)
            self._h = _lib._vp()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_z(self, z):
        z = _lib.fmat(z, rows=self.n)
        assert z.shape == (self.n, self.r)
        _lib.check(_lib.lib().cocons_ctx_set_z(self._h, _lib.ptr(z)))

    def set_xbetas(self, x_betas):
        xb = _lib.fmat(x_betas, rows=self.n)
        self.q = xb.shape[1]
        _lib.check(_lib.lib().cocons_ctx_set_xbetas(self._h, self.q, _lib.ptr(xb)))

    def terms(self, kind, theta_list, smooth_limits, mean=None):
        """One evaluation; returns dict(logdet, quad[r], logdet_w, rank).  Raises NotPositiveDefinite."""
        th = _lib.pack_theta(theta_list, self.p)
        lim = np.ascontiguousarray(np.asarray(smooth_limits, dtype=np.float64))
        mean_v = None if mean is None else np.ascontiguousarray(np.asarray(mean, dtype=np.float64))
        logdet, ldw = _lib.ctypes.c_double(), _lib.ctypes.c_double()
        rank = _lib.ctypes.c_int()
        quad = np.empty(self.r)
        _lib.check(_lib.lib().cocons_n2ll(self._h, int(kind), _lib.ptr(th), _lib.ptr(lim), _lib.ptr(mean_v),
                                          _lib.ctypes.byref(logdet), _lib.ptr(quad), _lib.ctypes.byref(ldw),
                                          _lib.ctypes.byref(rank)))
        return {"logdet": logdet.value, "quad": quad, "logdet_w": ldw.value, "rank": rank.value}

    def profile_betas(self, kind):
        k = self.q if kind == _lib.PROFILE else self.p
        out = np.empty(k)
        _lib.check(_lib.lib().cocons_profile_betas(self._h, int(kind), _lib.ptr(out)))
        return out

    def factor(self, theta_list, smooth_limits=None, type="diff"):
        th = _lib.pack_theta(theta_list, self.p)
        lim = None if smooth_limits is None else np.ascontiguousarray(np.asarray(smooth_limits, dtype=np.float64))
        par = _lib.PAR_CLASSIC if type == "classic" else _lib.PAR_DIFF
        _lib.check(_lib.lib().cocons_factor(self._h, par, _lib.ptr(th), _lib.ptr(lim)))

    def predict(self, locs_pred, x_covariates_pred, resid, want_explained=True):
        lp, Xp = _lib.fmat(locs_pred), _lib.fmat(x_covariates_pred)
        m = Xp.shape[0]
        resid = np.ascontiguousarray(np.asarray(resid, dtype=np.float64))
        sto = np.empty(m)
        expl = np.empty(m) if want_explained else None
        _lib.check(_lib.lib().cocons_predict(self._h, m, _lib.ptr(lp), _lib.ptr(Xp), _lib.ptr(resid), _lib.ptr(sto),
                                             _lib.ptr(expl)))
        return sto, expl

    def sim(self, eps):
        eps = _lib.fmat(eps, rows=self.n)
        out = np.empty_like(eps, order="F")
        _lib.check(_lib.lib().cocons_sim(self._h, eps.shape[1], _lib.ptr(eps), _lib.ptr(out)))
        return out

    def sim_cond(self, locs_pred, x_covariates_pred, eps):
        lp, Xp = _lib.fmat(locs_pred), _lib.fmat(x_covariates_pred)
        m = Xp.shape[0]
        eps = _lib.fmat(eps, rows=m)
        out = np.empty_like(eps, order="F")
        _lib.check(_lib.lib().cocons_sim_cond(self._h, m, _lib.ptr(lp), _lib.ptr(Xp), eps.shape[1], _lib.ptr(eps),
                                              _lib.ptr(out)))
        return out

    def get_factor(self):
        L = np.empty((self.n, self.n), order="F")
        perm = np.empty(self.n, dtype=np.int64)
        _lib.check(_lib.lib().cocons_ctx_get_factor(self._h, _lib.ptr(L), perm.ctypes.data_as(_lib._lp)))
        return L, perm

    def timings(self):
        ms = np.empty(4)
        _lib.lib().cocons_ctx_timings(self._h, _lib.ptr(ms))
        kms, kfl = _lib.ctypes.c_double(), _lib.ctypes.c_double()
        _lib.lib().cocons_ctx_kernel_timing(self._h, _lib.ctypes.byref(kms), _lib.ctypes.byref(kfl))
        return {"assembly_ms": ms[0], "factor_ms": ms[1], "solve_ms": ms[2], "total_ms": ms[3],
                "kernel_ms": kms.value, "kernel_flops": kfl.value}


# --------------------------------------------------------------------------
# objectives (R/neg2loglikelihood.R), one-shot: host buffers in, scalar out
# --------------------------------------------------------------------------
def _one_shot(kind, theta_list, locs, x_covariates, smooth_limits, z, n, x_betas=None):
    locs, X = _lib.fmat(locs), _lib.fmat(x_covariates)
    z = _lib.fmat(z, rows=n)
    p, r = X.shape[1], z.shape[1]
    xb = None if x_betas is None else _lib.fmat(x_betas, rows=n)
    q = 0 if xb is None else xb.shape[1]
    th = _lib.pack_theta(theta_list, p)
    lim = np.ascontiguousarray(np.asarray(smooth_limits, dtype=np.float64))
    mean = np.ascontiguousarray(np.asarray(theta_list["mean"], dtype=np.float64))
    logdet, ldw = _lib.ctypes.c_double(), _lib.ctypes.c_double()
    rank = _lib.ctypes.c_int()
    quad = np.empty(r)
    _lib.check(_lib.lib().cocons_neg2loglik_dense(int(kind), n, p, r, q, _lib.ptr(locs), _lib.ptr(X), _lib.ptr(z),
                                                  _lib.ptr(xb), _lib.ptr(th), _lib.ptr(lim), _lib.ptr(mean),
                                                  _lib.ctypes.byref(logdet), _lib.ptr(quad), _lib.ctypes.byref(ldw),
                                                  _lib.ctypes.byref(rank)))
    return {"logdet": logdet.value, "quad": quad, "logdet_w": ldw.value, "rank": rank.value}


def _combine(kind, t, n, r, lambda_, theta_list, smooth_limits):
    if kind == _lib.REML:  # R/neg2loglikelihood.R:283-289
        p = t["rank"]
        total = sum((n - p) * np.log(2 * np.pi) + 2 * t["logdet"] + 2 * t["logdet_w"] + qd for qd in t["quad"])
        return total + _getPen((n - p) * r, lambda_, theta_list, smooth_limits)
    total = sum(n * np.log(2 * np.pi) + 2 * t["logdet"] + qd for qd in t["quad"])  # :155-160, :212-218
    return total + _getPen(n * r, lambda_, theta_list, smooth_limits)


def _objective(kind, theta, par_pos, locs, x_covariates, smooth_limits, z, n, lambda_, safe, x_betas=None, ctx=None):
    theta_list = getModelLists(theta, par_pos, "diff")
    try:
        if ctx is not None:
            t = ctx.terms(kind, theta_list, smooth_limits, theta_list["mean"])
            r = ctx.r
        else:
            t = _one_shot(kind, theta_list, locs, x_covariates, smooth_limits, z, n, x_betas)
            r = len(t["quad"])
    except NotPositiveDefinite:
        if safe:
            return 1e6  # R/neg2loglikelihood.R:138-142, 202-206, 261-265
        raise ArithmeticError("Cholesky error")
    return _combine(kind, t, n, r, lambda_, theta_list, smooth_limits)


def GetNeg2loglikelihood(theta, par_pos, locs, x_covariates, smooth_limits, z, n, lambda_, safe=True, ctx=None):
    """R/neg2loglikelihood.R:183-222.  `ctx` (a DenseLikelihood) keeps the data on the device between
    calls; without it every call uploads locs / x_covariates / z, as the R closure would."""
    return _objective(_lib.ML, theta, par_pos, locs, x_covariates, smooth_limits, z, n, lambda_, safe, None, ctx)


def GetNeg2loglikelihoodProfile(theta, par_pos, locs, x_covariates, smooth_limits, z, n, x_betas, lambda_, safe=True,
                                ctx=None):
    """R/neg2loglikelihood.R:127-165."""
    return _objective(_lib.PROFILE, theta, par_pos, locs, x_covariates, smooth_limits, z, n, lambda_, safe, x_betas,
                      ctx)


def GetNeg2loglikelihoodREML(theta, par_pos, locs, x_covariates, x_betas, smooth_limits, z, n, lambda_, safe=True,
                             ctx=None):
    """R/neg2loglikelihood.R:241-291 (x_betas is accepted and unused there too; z are the contrasts of
    R/optim.R:311, see reml_contrasts())."""
    return _objective(_lib.REML, theta, par_pos, locs, x_covariates, smooth_limits, z, n, lambda_, safe, None, ctx)


def reml_contrasts(mod_DM, z):
    """R/optim.R:311 without forming the n x n projector: z - X (X'X)^-1 X' z."""
    X = np.asarray(mod_DM, dtype=np.float64)
    z = np.asarray(z, dtype=np.float64).reshape(X.shape[0], -1)
    return z - X @ np.linalg.solve(X.T @ X, X.T @ z)


# --------------------------------------------------------------------------
# coco objects and the user-facing verbs
# --------------------------------------------------------------------------
class coco:
    """S4 class `coco` (R/methods.R:17-25) built by coco() (R/cocons.R:84-175)."""

    def __init__(self, type, data, locs, z, model_list, info=None, output=None):
        if type not in ("dense", "sparse"):
            raise ValueError("type must be 'dense' or 'sparse'")
        if type == "sparse":
            raise NotImplementedError("the tapered (sparse) model is outside this build's scope (SURVEY.md §8f N3)")
        self.type = type
        self.data = _columns(data)
        self.locs = np.asfortranarray(np.asarray(locs, dtype=np.float64))
        z = np.asarray(z, dtype=np.float64)
        self.z = np.asfortranarray(z.reshape(-1, 1) if z.ndim == 1 else z)
        ml = dict(model_list)
        ml.setdefault("mean", 0)
        ml.setdefault("aniso", 0)
        ml.setdefault("tilt", 0)
        ml.setdefault("smooth", 0.5)
        ml.setdefault("nugget", -np.inf)
        self.model_list = {k: ml[k] for k in DICTIONARY}  # :132
        info = dict(info or {})
        info.setdefault("lambda.reg", 0)
        info.setdefault("lambda.betas", 0)
        info.setdefault("lambda.Sigma", 0)
        if not is_formula(self.model_list["smooth"]):  # :157-162
            s = float(np.atleast_1d(self.model_list["smooth"])[0])
            info["smooth.limits"] = np.array([s, s])
        if "smooth.limits" not in info:
            raise ValueError("info['smooth.limits'] is required when smooth is a formula")
        info["smooth.limits"] = np.asarray(info["smooth.limits"], dtype=np.float64)
        self.info = info
        self.output = dict(output or {})


def getCovMatrix(coco_object):
    """R/getFunctions.R:35-52 (dense, type 'global')."""
    x_covs = getScale(coco_object)["std.covs"]
    par_pos = getDesignMatrix(coco_object.model_list, coco_object.data)["par.pos"]
    theta_list = getModelLists(coco_object.output["par"], par_pos, "diff")
    return cov_rns(theta_list, coco_object.locs, x_covs, coco_object.info["smooth.limits"])


def fd_value_and_grad(fn, theta, lower, upper, ndeps, forward=False, group=None, batch=None):
    """What optimParallel does per L-BFGS-B iteration (R/optim.R:157,256,321; options at
    R/profile.R:9-16): fn(theta) and its finite-difference neighbours (2p central, or p forward)
    evaluated as ONE batch of independent points.  With torch.distributed initialised (one process
    per GPU) the batch is dealt to the ranks by distributed.fan_out(); every rank gets every value,
    so all ranks take the same optimiser step."""
    from .distributed import fan_out
    theta = np.asarray(theta, dtype=np.float64)
    p = theta.shape[0]
    pts, spans = [theta.copy()], []
    for i in range(p):
        up, dn = theta.copy(), theta.copy()
        up[i] = min(theta[i] + ndeps, upper[i])
        dn[i] = theta[i] if forward else max(theta[i] - ndeps, lower[i])
        if up[i] == dn[i]:  # pinned at a degenerate bound
            dn[i] = max(theta[i] - ndeps, lower[i])
        pts.append(up)
        if not forward:
            pts.append(dn)
        spans.append((up[i], dn[i]))
    if batch is not None:  # several evaluations in flight on this GPU (DenseLikelihoodPool.map)
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            world, rank = dist.get_world_size(group), dist.get_rank(group)
            mine = batch(pts[rank::world])
            it = iter(mine)
            vals = fan_out(pts, lambda _pt: next(it), group=group)
        else:
            vals = batch(pts)
    else:
        vals = fan_out(pts, fn, group=group)
    f0 = vals[0]
    grad = np.empty(p)
    for i in range(p):
        up_i, dn_i = spans[i]
        if forward:
            grad[i] = (vals[1 + i] - f0) / (up_i - dn_i) if up_i != dn_i else 0.0
        else:
            grad[i] = (vals[1 + 2 * i] - vals[2 + 2 * i]) / (up_i - dn_i) if up_i != dn_i else 0.0
    return f0, grad


def cocoOptim(coco_object, boundaries, ncores="auto", safe=True, optim_type="ml", optim_control=None, device=0,
              forward=False):
    """R/optim.R:65-365, dense branch.  L-BFGS-B (scipy) stands in for optimParallel's optimiser; its
    gradient is the same batched finite-difference scheme (fd_value_and_grad), whose independent
    objective evaluations run on the device-resident context of each rank - across all GPUs when
    launched with one process per GPU - instead of on forked R workers.
    `boundaries` = dict(theta_init, theta_lower, theta_upper)."""
    from scipy.optimize import minimize

    dm = getDesignMatrix(coco_object.model_list, coco_object.data)
    sc = getScale(dm["model.matrix"])
    mod_DM = sc["std.covs"]
    n = coco_object.z.shape[0]
    lim = coco_object.info["smooth.limits"]
    lam = (0.0, 0.0, coco_object.info["lambda.reg"])
    if optim_type == "ml":
        lam = (coco_object.info["lambda.Sigma"], coco_object.info["lambda.betas"], coco_object.info["lambda.reg"])
    optim_type = optim_type.lower()
    ctrl = {"maxiter": 500, "ftol": 1e-8, "maxcor": 100}  # R/profile.R:9-16 (factr, maxit, lmm)
    ctrl.update(optim_control or {})
    ndeps = ctrl.pop("ndeps", np.finfo(float).eps ** 0.25)
    init = np.asarray(boundaries["theta_init"], dtype=np.float64)
    lower = np.asarray(boundaries["theta_lower"], dtype=np.float64)
    upper = np.asarray(boundaries["theta_upper"], dtype=np.float64)
    par_pos = dm["par.pos"]
    z = coco_object.z
    x_betas = None
    if optim_type in ("pml", "reml"):
        if not isinstance(par_pos["mean"], np.ndarray):
            raise ValueError("Profile ML or Restricted ML only available when considering covariates in the mean.")
        x_betas = mod_DM[:, par_pos["mean"]]
        nb = int(par_pos["mean"].sum())
        init, lower, upper = init[nb:], lower[nb:], upper[nb:]
        par_pos = dict(par_pos)
        par_pos["mean"] = np.zeros_like(dm["par.pos"]["mean"])
        if optim_type == "reml":
            z = reml_contrasts(mod_DM, z)
    # `ncores` keeps its meaning of "objective evaluations in flight" (R/optim.R:80-86): here they are
    # contexts on one GPU; "auto" = as many as fit comfortably, none extra once one evaluation fills the GPU
    if ncores == "auto":
        ncores = max(1, min(8, int(4e9 // (8 * n * n))))
    with DenseLikelihoodPool(coco_object.locs, mod_DM, z, size=int(ncores), device=device) as pool:
        ctx = pool.ctxs[0]
        if optim_type == "pml":
            pool.set_xbetas(x_betas)
        kind = {"ml": _lib.ML, "pml": _lib.PROFILE, "reml": _lib.REML}[optim_type]

        def fn_on(c, theta):
            return _objective(kind, theta, par_pos, None, None, lim, None, n, lam, safe, ctx=c)

        def fn(theta):
            return fn_on(ctx, theta)

        res = minimize(lambda th: fd_value_and_grad(fn, th, lower, upper, ndeps, forward=forward,
                                                    batch=lambda pts: pool.map(fn_on, pts)),
                       init, jac=True, method="L-BFGS-B", bounds=list(zip(lower, upper)), options=ctrl)
        par = res.x
        if optim_type in ("pml", "reml"):  # R/optim.R:326-345
            theta_list = getModelLists(par, par_pos, "diff")
            ctx.set_z(coco_object.z)
            ctx.set_xbetas(x_betas)
            ctx.factor(theta_list, lim)
            betas = ctx.profile_betas(_lib.PROFILE)
            par = np.concatenate([betas, par])
    coco_object.output = {"par": par, "value": res.fun, "counts": res.nfev, "convergence": res.status,
                          "message": res.message}
    coco_object.info.update({"mean.vector": sc["mean.vector"], "sd.vector": sc["sd.vector"],
                             "optim.type": optim_type, "safe": safe, "boundaries": boundaries})
    return coco_object


def getHessian(coco_object, ncores="auto", eps=np.finfo(float).eps ** 0.25, device=0):
    """R/getFunctions.R:925-1034, dense branch: forward-difference Hessian of the NEGATIVE log-likelihood
    at the fitted parameters, H[j,i] = 0.5 (f11 - f01 - f10 + f00) / eps^2 with f = -2 loglik (for "pml"
    fits on the full likelihood, as there).  The reference farms three evaluations per index pair over a
    PSOCK cluster; here the p + p(p+1)/2 distinct points are evaluated as one batch on the GPU(s)."""
    if coco_object.info.get("optim.type") == "reml":
        raise NotImplementedError("reml hessian not implemented yet.")  # as the reference
    par = np.asarray(coco_object.output["par"], dtype=np.float64)
    p = par.shape[0]
    dm = getDesignMatrix(coco_object.model_list, coco_object.data)
    X = getScale(dm["model.matrix"], coco_object.info["mean.vector"], coco_object.info["sd.vector"])["std.covs"]
    n = coco_object.z.shape[0]
    lim = coco_object.info["smooth.limits"]
    lam = (coco_object.info["lambda.Sigma"], coco_object.info["lambda.betas"], coco_object.info["lambda.reg"])
    if ncores == "auto":
        ncores = max(1, min(8, int(4e9 // (8 * n * n))))
    singles = [par + eps * np.eye(p)[j] for j in range(p)]
    pairs = [(j, i) for j in range(p) for i in range(j, p)]
    doubles = [par + eps * (np.eye(p)[j] + np.eye(p)[i]) for j, i in pairs]
    with DenseLikelihoodPool(coco_object.locs, X, coco_object.z, size=int(ncores), device=device) as pool:
        def fn_on(c, theta):
            return _objective(_lib.ML, theta, dm["par.pos"], None, None, lim, None, n, lam, True, ctx=c)
        vals = pool.map(fn_on, [par] + singles + doubles)
    f00 = vals[0] if coco_object.info.get("optim.type") == "pml" else coco_object.output["value"]
    f1 = vals[1:1 + p]
    H = np.zeros((p, p))
    for (j, i), f11 in zip(pairs, vals[1 + p:]):
        H[j, i] = 0.5 * ((f11 - f1[j] - f1[i] + f00) / (eps * eps))
    H = H + H.T
    H[np.diag_indices(p)] /= 2
    return H


def cocoPredict(coco_object, newdataset, newlocs, type="mean", index_pred=0, device=0):
    """R/predict.R:84-188, dense branch: kriging mean and (type "pred") standard deviation.  The
    reference solves with LU (`solve`); here the Cholesky factor on the device is used."""
    if not coco_object.output:
        raise ValueError("coco object has not yet been fitted.")
    dm = getDesignMatrix(coco_object.model_list, coco_object.data)
    eff = getModelLists(coco_object.output["par"], dm["par.pos"], "diff")
    X_std = getScale(dm["model.matrix"], coco_object.info["mean.vector"], coco_object.info["sd.vector"])["std.covs"]
    dmp = getDesignMatrix(coco_object.model_list, newdataset)
    X_pred = getScale(dmp["model.matrix"], coco_object.info["mean.vector"], coco_object.info["sd.vector"])["std.covs"]
    systematic = X_pred @ eff["mean"]
    resid = coco_object.z[:, index_pred] - X_std @ eff["mean"]
    with DenseLikelihood(coco_object.locs, X_std, coco_object.z, device=device) as ctx:
        ctx.factor(eff, coco_object.info["smooth.limits"])
        sto, expl = ctx.predict(newlocs, X_pred, resid, want_explained=(type == "pred"))
    out = {"systematic": systematic, "stochastic": sto}
    if type == "pred":  # :170-183
        u = 1 / np.exp(-(X_pred @ eff["std.dev"])) + np.exp(X_pred @ eff["nugget"]) - expl
        neg = u < 1e-10
        u[neg] = np.abs(u[neg])
        out["sd.pred"] = np.sqrt(u)
    return out


def cocoSim(coco_object, pars=None, n=1, seed=None, standardize=True, type="classic", sim_type=None, cond_info=None,
            device=0, eps=None):
    """R/sim.R:52-175, dense branch.  The N(0,1) draws come from numpy's legacy generator seeded
    with `seed` (R's come from set.seed/rnorm; pass `eps` to supply R's own draws)."""
    if pars is None:
        pars, type = coco_object.output["par"], "diff"
    dm = getDesignMatrix(coco_object.model_list, coco_object.data)
    rng = np.random.RandomState(seed)
    if sim_type == "cond":
        std_coco = getScale(dm["model.matrix"])["std.covs"]  # freshly standardised, as at :76
        dmp = getDesignMatrix(coco_object.model_list, cond_info["newdataset"])
        std_pred = getScale(dmp["model.matrix"], coco_object.info["mean.vector"],
                            coco_object.info["sd.vector"])["std.covs"]
        to_pass = getModelLists(coco_object.output["par"], dmp["par.pos"], "diff")
        newlocs = np.asarray(cond_info["newlocs"], dtype=np.float64)
        m = newlocs.shape[0]
        if eps is None:
            eps = rng.standard_normal((m, n))
        with DenseLikelihood(coco_object.locs, std_coco, coco_object.z, device=device) as ctx:
            ctx.factor(to_pass, coco_object.info["smooth.limits"])
            draws = ctx.sim_cond(newlocs, std_pred, eps)
        step_one = cocoPredict(coco_object, cond_info["newdataset"], newlocs, "mean", device=device)
        return draws + (step_one["systematic"] + step_one["stochastic"])[:, None]
    if standardize:
        std_coco = getScale(dm["model.matrix"])["std.covs"]
    else:
        k = dm["model.matrix"].shape[1]
        std_coco = getScale(dm["model.matrix"], np.zeros(k), np.ones(k))["std.covs"]
    theta_to_fit = getModelLists(pars, dm["par.pos"], type)
    if not is_formula(coco_object.model_list["smooth"]):  # :141-145
        theta_to_fit["smooth"][0] = np.log(coco_object.info["smooth.limits"][0])
    nsites = std_coco.shape[0]
    if eps is None:
        eps = rng.standard_normal((nsites, n))
    with DenseLikelihood(coco_object.locs, std_coco, coco_object.z, device=device) as ctx:
        ctx.factor(theta_to_fit, coco_object.info["smooth.limits"], type=type)
        draws = ctx.sim(eps)
    return draws + (std_coco @ theta_to_fit["mean"])[:, None]


# --------------------------------------------------------------------------
# several evaluations in flight on ONE GPU
# --------------------------------------------------------------------------
class DenseLikelihoodPool:
    """`size` device-resident contexts over the same data, each with its own streams, driven by host
    threads (the C ABI releases the GIL).  Below n ~ 10 000 one evaluation is a latency-bound chain of
    small kernels that leaves most of the 148 SMs idle; running the optimiser's independent
    finite-difference points (R/optim.R:157,256,321) side by side fills them."""

    def __init__(self, locs, x_covariates, z, size=4, device=0):
        from concurrent.futures import ThreadPoolExecutor
        self.ctxs = [DenseLikelihood(locs, x_covariates, z, device=device) for _ in range(int(size))]
        self.exec = ThreadPoolExecutor(max_workers=len(self.ctxs))
        self.n, self.p, self.r = self.ctxs[0].n, self.ctxs[0].p, self.ctxs[0].r

    def close(self):
        self.exec.shutdown(wait=True)
        for c in self.ctxs:
            c.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_xbetas(self, xb):
        for c in self.ctxs:
            c.set_xbetas(xb)

    def set_z(self, z):
        for c in self.ctxs:
            c.set_z(z)

    def map(self, fn, points):
        """[fn(ctx, point) for point in points], evaluated len(ctxs) at a time."""
        import queue
        free = queue.SimpleQueue()
        for c in self.ctxs:
            free.put(c)

        def run(pt):
            c = free.get()
            try:
                return fn(c, pt)
            finally:
                free.put(c)

        return list(self.exec.map(run, points))
