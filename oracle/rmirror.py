"""TEST INFRASTRUCTURE ONLY (oracle/): numpy/LAPACK restatement of the R side of
the dense-likelihood path, written literally - including the reference's
explicit `chol2inv` + projector route - so that the *reference's* rounding
behaviour is what the CUDA path is compared with.

R's base LAPACK is replaced by scipy's (OpenBLAS): chol -> dpotrf('U'),
forwardsolve/backsolve -> dtrtrs, chol2inv -> dpotri, solve -> dgesv.
Parity status: see oracle/cov_oracle.cpp (the covariance is pinned against the
reference's own compiled source; the algebra below has no reference-side golden
values - the reference's tests hold none - and is cross-checked in
tests/test_oracle.py through the identities of SURVEY.md §8c(v)).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import
this module.
"""
import numpy as np
import scipy.linalg as sla

from . import cov as _cov

ASPECT_ORDER = ("mean", "std.dev", "scale", "aniso", "tilt", "smooth", "nugget")


class CholeskyError(Exception):
    pass


def get_scale(x, mean_vector=None, sd_vector=None):
    """R/getFunctions.R:376-436 (matrix branch): centre/scale columns 2..p, sd with n-1."""
    x = np.array(x, dtype=np.float64, order="F", copy=True)
    if mean_vector is None:
        mean_vector = x.mean(axis=0)
        mean_vector[0] = 0.0
    if sd_vector is None:
        sd_vector = x.std(axis=0, ddof=1) if x.shape[0] > 1 else np.full(x.shape[1], np.nan)
        sd_vector[0] = 1.0
    for k in range(1, x.shape[1]):
        x[:, k] = (x[:, k] - mean_vector[k]) / sd_vector[k]
    return {"std.covs": x, "mean.vector": np.asarray(mean_vector), "sd.vector": np.asarray(sd_vector)}


def get_model_lists(theta, par_pos, type="diff"):
    """R/getFunctions.R:570-616.  par_pos: dict aspect -> bool array (free) or number (fixed)."""
    theta = np.asarray(theta, dtype=np.float64)
    is_free = {k: isinstance(v, np.ndarray) and v.dtype == bool for k, v in par_pos.items()}
    length = max(len(v) if is_free[k] else 1 for k, v in par_pos.items())
    out, used = {}, 0
    for name, pos in par_pos.items():
        vec = np.zeros(length)
        if not is_free[name]:
            vec[0] = float(np.atleast_1d(pos)[0])
        else:
            k = int(pos.sum())
            vec[np.flatnonzero(pos)] = theta[used:used + k]
            used += k
        out[name] = vec
    if type == "classic":
        return out
    res = {k: v.copy() for k, v in out.items()}
    if is_free["std.dev"] and is_free["scale"]:
        for i in range(length):
            if par_pos["std.dev"][i] and par_pos["scale"][i]:
                res["std.dev"][i] = (out["std.dev"][i] + out["scale"][i]) / 2
                res["scale"][i] = (out["std.dev"][i] - out["scale"][i]) / 2
    return res


def get_pen(n, lam, theta_list, smooth_limits):
    """R/checkFunctions.R:474-492 (theta_list in ASPECT_ORDER: [[1]] is mean, 2..6 the next five)."""
    names = list(theta_list.keys())
    summ = lam[2] * np.exp(theta_list["scale"][0]) * np.sqrt(
        (smooth_limits[1] - smooth_limits[0]) / (1 + np.exp(-theta_list["smooth"][0])) + smooth_limits[0]
    ) + _cov.sumsmoothlone(theta_list[names[0]][1:], lam[1])
    for ii in range(1, 6):
        summ = summ + _cov.sumsmoothlone(theta_list[names[ii]][1:], lam[0])
    return 2 * n * summ


def r_chol(S):
    """base::chol: upper R with S = R'R, error on a non-positive leading minor."""
    try:
        return sla.cholesky(S, lower=False, check_finite=True)
    except (sla.LinAlgError, ValueError) as e:
        raise CholeskyError(str(e))


def _fwd(R, b):  # forwardsolve(R, b, transpose=TRUE, upper.tri=TRUE): R' y = b
    return sla.solve_triangular(R, b, trans="T", lower=False, check_finite=False)


def _back(R, b):  # backsolve(R, b): R x = b
    return sla.solve_triangular(R, b, trans="N", lower=False, check_finite=False)


def _chol2inv(R):
    inv, info = sla.lapack.dpotri(R, lower=0)
    assert info == 0
    iu = np.triu_indices_from(inv, 1)
    inv[(iu[1], iu[0])] = inv[iu]
    return inv


def neg2loglik(theta, par_pos, locs, x_covariates, smooth_limits, z, n, lam, safe=True, cov_kind="restatement",
               Sigma=None):
    """GetNeg2loglikelihood, R/neg2loglikelihood.R:183-222.  (Sigma: reuse an already assembled cov_rns.)"""
    tl = get_model_lists(theta, par_pos, "diff")
    S = _cov.cov_rns(tl, locs, x_covariates, smooth_limits, kind=cov_kind) if Sigma is None else Sigma
    try:
        R = r_chol(S)
    except CholeskyError:
        if safe:
            return 1e6
        raise CholeskyError("Cholesky error")
    logdet = np.sum(np.log(np.diag(R)))
    trend = np.asarray(x_covariates) @ tl["mean"]
    z = np.asarray(z, dtype=np.float64).reshape(n, -1)
    total = 0.0
    for c in range(z.shape[1]):
        y = _fwd(R, z[:, c] - trend)
        total += n * np.log(2 * np.pi) + 2 * logdet + float(y @ y)
    return total + get_pen(n * z.shape[1], lam, tl, smooth_limits)


def _projector(R, X):
    V = _back(R, _fwd(R, X))
    W = X.T @ V
    Sinv = _chol2inv(R)
    P = Sinv - V @ np.linalg.solve(W, V.T)
    return P, W


def neg2loglik_profile(theta, par_pos, locs, x_covariates, smooth_limits, z, n, x_betas, lam, safe=True,
                       cov_kind="restatement", Sigma=None):
    """GetNeg2loglikelihoodProfile, R/neg2loglikelihood.R:127-165."""
    tl = get_model_lists(theta, par_pos, "diff")
    S = _cov.cov_rns(tl, locs, x_covariates, smooth_limits, kind=cov_kind) if Sigma is None else Sigma
    try:
        R = r_chol(S)
    except CholeskyError:
        if safe:
            return 1e6
        raise CholeskyError("Cholesky error")
    P, _ = _projector(R, np.asarray(x_betas, dtype=np.float64).reshape(n, -1))
    logdet = np.sum(np.log(np.diag(R)))
    z = np.asarray(z, dtype=np.float64).reshape(n, -1)
    total = 0.0
    for c in range(z.shape[1]):
        total += n * np.log(2 * np.pi) + 2 * logdet + float(z[:, c] @ (P @ z[:, c]))
    return total + get_pen(n * z.shape[1], lam, tl, smooth_limits)


def neg2loglik_reml(theta, par_pos, locs, x_covariates, x_betas, smooth_limits, z, n, lam, safe=True,
                    cov_kind="restatement", Sigma=None):
    """GetNeg2loglikelihoodREML, R/neg2loglikelihood.R:241-291 (x_betas is accepted and unused, as there)."""
    tl = get_model_lists(theta, par_pos, "diff")
    S = _cov.cov_rns(tl, locs, x_covariates, smooth_limits, kind=cov_kind) if Sigma is None else Sigma
    try:
        R = r_chol(S)
    except CholeskyError:
        if safe:
            return 1e6
        raise CholeskyError("Cholesky error")
    X = np.asarray(x_covariates, dtype=np.float64)
    logdet = np.sum(np.log(np.diag(R)))
    p = r_qr_rank(X)
    P, W = _projector(R, X)
    cholW = r_chol(W)
    z = np.asarray(z, dtype=np.float64).reshape(n, -1)
    total = 0.0
    for c in range(z.shape[1]):
        total += (n - p) * np.log(2 * np.pi) + 2 * logdet + 2 * np.sum(np.log(np.diag(cholW))) + float(
            z[:, c] @ (P @ z[:, c]))
    return total + get_pen((n - p) * z.shape[1], lam, tl, smooth_limits)


def r_qr_rank(X, tol=1e-7):
    """qr(X)$rank - LINPACK dqrdc2 limited pivoting: a column whose residual norm falls
    below tol x its original norm is moved to the end (R's src/appl/dqrdc2.f)."""
    A = np.array(X, dtype=np.float64, order="F", copy=True)
    n, p = A.shape
    orig = np.linalg.norm(A, axis=0)
    orig[orig == 0] = 1.0
    rank, k = p, 0
    order = list(range(p))
    while k < rank:
        while k < rank and np.linalg.norm(A[k:, k]) < tol * orig[order[k]]:
            A[:, k:] = np.roll(A[:, k:], -1, axis=1)
            order = order[:k] + order[k + 1:] + [order[k]]
            rank -= 1
        if k >= rank:
            break
        v = A[k:, k].copy()
        nrm = np.linalg.norm(v)
        if nrm != 0:
            v[0] += np.copysign(nrm, v[0] if v[0] != 0 else 1.0)
            v /= np.linalg.norm(v)
            A[k:, k:] -= 2 * np.outer(v, v @ A[k:, k:])
        k += 1
    return rank


def reml_contrast(mod_DM, z):
    """R/optim.R:311 - z pre-multiplied by I - X (X'X)^-1 X' (formed explicitly there)."""
    X = np.asarray(mod_DM, dtype=np.float64)
    n = X.shape[0]
    Pm = np.eye(n) - X @ np.linalg.solve(X.T @ X, X.T)
    return Pm @ np.asarray(z, dtype=np.float64).reshape(n, -1)


def profile_betas(theta_list, locs, x_covariates, smooth_limits, x_betas, z, cov_kind="restatement"):
    """R/optim.R:326-343 - beta recovery after pml/reml."""
    S = _cov.cov_rns(theta_list, locs, x_covariates, smooth_limits, kind=cov_kind)
    L = r_chol(S)
    Xb = np.asarray(x_betas, dtype=np.float64).reshape(S.shape[0], -1)
    V = _back(L, _fwd(L, Xb))
    W = Xb.T @ V
    z = np.asarray(z, dtype=np.float64).reshape(S.shape[0], -1)
    return (np.linalg.solve(W, V.T) @ z.sum(axis=1)) / z.shape[1]


def predict(theta_list, locs, newlocs, X_std, X_pred_std, smooth_limits, z_col, type="mean", cov_kind="restatement"):
    """cocoPredict dense branch, R/predict.R:136-187 (LU solve, abs() of tiny negative variances)."""
    S = _cov.cov_rns(theta_list, locs, X_std, smooth_limits, kind=cov_kind)
    C = _cov.cov_rns_pred(theta_list, locs, newlocs, X_std, X_pred_std, smooth_limits, kind=cov_kind)
    inv_cov = np.linalg.solve(S, C.T)
    systematic = X_pred_std @ theta_list["mean"]
    resid = np.asarray(z_col, dtype=np.float64) - X_std @ theta_list["mean"]
    out = {"systematic": systematic, "stochastic": resid @ inv_cov}
    if type == "pred":
        u = 1 / np.exp(-(X_pred_std @ theta_list["std.dev"])) + np.exp(X_pred_std @ theta_list["nugget"])
        u = u - np.sum(C * inv_cov.T, axis=1)
        neg = u < 1e-10
        u[neg] = np.abs(u[neg])
        out["sd.pred"] = np.sqrt(u)
    return out


def sim_marginal(theta_list, locs, X_std, smooth_limits, eps, type="diff", cov_kind="restatement"):
    """cocoSim dense marginal branch, R/sim.R:147-172 with the N(0,1) draws `eps` (n x k) supplied."""
    if type == "classic":
        S = _cov.cov_rns_classic(theta_list, locs, X_std, kind=cov_kind)
    else:
        S = _cov.cov_rns(theta_list, locs, X_std, smooth_limits, kind=cov_kind)
    R = r_chol(S)
    mu = X_std @ theta_list["mean"]
    eps = np.asarray(eps, dtype=np.float64).reshape(S.shape[0], -1)
    return (eps.T @ R + mu[None, :]).T


def sim_conditional(theta_list, locs, newlocs, X_std, X_pred_std, smooth_limits, z_col, eps, cov_kind="restatement"):
    """cocoSim conditional branch, R/sim.R:87-121 (Schur complement, then predictive mean + L' eps)."""
    S = _cov.cov_rns(theta_list, locs, X_std, smooth_limits, kind=cov_kind)
    C = _cov.cov_rns_pred(theta_list, locs, newlocs, X_std, X_pred_std, smooth_limits, kind=cov_kind)
    Su = _cov.cov_rns(theta_list, newlocs, X_pred_std, smooth_limits, kind=cov_kind)
    L = r_chol(Su - C @ np.linalg.solve(S, C.T))
    pr = predict(theta_list, locs, newlocs, X_std, X_pred_std, smooth_limits, z_col, "mean", cov_kind)
    mu = pr["systematic"] + pr["stochastic"]
    eps = np.asarray(eps, dtype=np.float64).reshape(C.shape[0], -1)
    return (eps.T @ L + mu[None, :]).T


# ---- sparse (tapered) model ---------------------------------------------------------------
# spam (>= 2.9.1, DESCRIPTION:21) is third-party and absent: its sparse matrices are (entries,
# colindices, rowpointers) triples here, nearest.dist is a brute-force distance threshold, cov.wend1
# is its documented formula, and its sparse Cholesky / solve are replaced by LAPACK on the dense
# expansion - determinant, quadratic forms and solutions do not depend on spam's fill-reducing
# permutation, so the values are those of the reference up to rounding.
def nearest_dist(x, y=None, delta=1.0):
    """spam::nearest.dist(..., upper = NULL): Euclidean distances <= delta, zero distances and the diagonal kept.
    Returns (distances, colindices, rowpointers), 1-based CSR with ascending columns."""
    x = np.asarray(x, dtype=np.float64)
    y = x if y is None else np.asarray(y, dtype=np.float64)
    ent, ci, rp = [], [], [1]
    for i in range(x.shape[0]):
        d = np.sqrt((x[i, 0] - y[:, 0]) ** 2 + (x[i, 1] - y[:, 1]) ** 2)
        j = np.flatnonzero(d <= delta)
        ent.append(d[j]), ci.append(j + 1)
        rp.append(rp[-1] + len(j))
    return np.concatenate(ent), np.concatenate(ci), np.asarray(rp)


def cov_wend1(h, theta):
    """spam::cov.wend1(h, theta = c(range, sill)): sill (1 - d)^4_+ (1 + 4 d), d = h / range."""
    d = np.asarray(h, dtype=np.float64) / theta[0]
    return theta[1] * np.where(d < 1, (1 - d) ** 4 * (1 + 4 * d), 0.0)


def csr_to_dense(entries, colindices, rowpointers, ncol):
    nrow = len(rowpointers) - 1
    out = np.zeros((nrow, ncol))
    rows = np.repeat(np.arange(nrow), np.diff(rowpointers))
    out[rows, np.asarray(colindices, dtype=np.int64) - 1] = entries
    return out


def tapered_matrix(theta_list, taper, colindices, rowpointers, locs, x_covariates, smooth_limits,
                   cov_kind="restatement"):
    """ref_taper@entries * cov_rns_taper(...) (R/neg2loglikelihood.R:26-31), expanded to a dense n x n matrix.
    spam's Cholesky reads one triangle of the (ulp-level asymmetric) matrix; the lower one is used here."""
    ent = taper * _cov.cov_rns_taper(theta_list, locs, x_covariates, colindices, rowpointers, smooth_limits,
                                     kind=cov_kind)
    S = np.tril(csr_to_dense(ent, colindices, rowpointers, len(rowpointers) - 1))
    return S + np.tril(S, -1).T


def _taper_chol_terms(tl, taper, colindices, rowpointers, locs, x_covariates, smooth_limits, z, n, cov_kind):
    S = tapered_matrix(tl, taper, colindices, rowpointers, locs, x_covariates, smooth_limits, cov_kind)
    R = r_chol(S)
    logdet = np.sum(np.log(np.diag(R)))  # determinant.spam.chol.NgPeyton(cholS)$modulus
    trend = np.asarray(x_covariates) @ tl["mean"]
    z = np.asarray(z, dtype=np.float64).reshape(n, -1)
    quads = []
    for c in range(z.shape[1]):
        y = _fwd(R, z[:, c] - trend)
        quads.append(float(y @ y))
    return logdet, quads


def neg2loglik_taper(theta, par_pos, taper, colindices, rowpointers, locs, x_covariates, smooth_limits, z, n, lam,
                     safe=True, cov_kind="restatement"):
    """GetNeg2loglikelihoodTaper, R/neg2loglikelihood.R:20-53."""
    tl = get_model_lists(theta, par_pos, "diff")
    try:
        logdet, quads = _taper_chol_terms(tl, taper, colindices, rowpointers, locs, x_covariates, smooth_limits, z,
                                          n, cov_kind)
    except CholeskyError:
        if safe:
            return 1e6
        raise CholeskyError("Cholesky error")
    total = sum(n * np.log(2 * np.pi) + 2 * logdet + q for q in quads)
    return total + get_pen(n * len(quads), lam, tl, smooth_limits)


def neg2loglik_taper_profile(theta, par_pos, taper, colindices, rowpointers, locs, x_covariates, smooth_limits, z,
                             n, lam, safe=True, cov_kind="restatement"):
    """GetNeg2loglikelihoodTaperProfile, R/neg2loglikelihood.R:73-108."""
    tl = get_model_lists(theta, par_pos, "diff")
    tl["std.dev"][0] = 0.0  # :78
    try:
        logdet, quads = _taper_chol_terms(tl, taper, colindices, rowpointers, locs, x_covariates, smooth_limits, z,
                                          n, cov_kind)
    except CholeskyError:
        if safe:
            return 1e6
        raise CholeskyError("Cholesky error")
    r = len(quads)
    sum_in = float(np.sum(quads))
    return (r * n * np.log(2 * np.pi) + r * n + r * 2 * logdet + r * n * np.log(sum_in / (r * n)) +
            get_pen(n * r, lam, tl, smooth_limits))


def predict_taper(theta_list, delta, locs, newlocs, X_std, X_pred_std, smooth_limits, z_col, type="mean",
                  taper_fn=cov_wend1, cov_kind="restatement"):
    """cocoPredict sparse branch, R/predict.R:219-288."""
    n, m = len(locs), len(newlocs)
    d, ci, rp = nearest_dist(locs, delta=delta)
    S = tapered_matrix(theta_list, taper_fn(d, (delta, 1)), ci, rp, locs, X_std, smooth_limits, cov_kind)
    dp, cip, rpp = nearest_dist(newlocs, locs, delta=delta)
    ent = taper_fn(dp, (delta, 1)) * _cov.cov_rns_taper_pred(theta_list, locs, newlocs, X_std, X_pred_std, cip, rpp,
                                                             smooth_limits, kind=cov_kind)
    C = csr_to_dense(ent, cip, rpp, n)
    inv_cov = np.linalg.solve(S, C.T)
    systematic = X_pred_std @ theta_list["mean"]
    resid = np.asarray(z_col, dtype=np.float64) - X_std @ theta_list["mean"]
    out = {"systematic": systematic, "stochastic": resid @ inv_cov}
    if type == "pred":
        u = 1 / np.exp(-(X_pred_std @ theta_list["std.dev"])) + np.exp(X_pred_std @ theta_list["nugget"])
        u = u - np.sum(C * inv_cov.T, axis=1)
        neg = u < 1e-10
        u[neg] = np.abs(u[neg])
        out["sd.pred"] = np.sqrt(u)
    return out
