# R side of the drop-in: these definitions replace R/RcppExports.R (the four dense entry points
# and sumsmoothlone keep their names, argument names and order) and the dense halves of
# R/neg2loglikelihood.R.  Everything else in the package (coco(), cocoOptim(), getModelLists(),
# getScale(), .cocons.getPen(), ...) stays as it is and keeps calling these.
#
# Not executable in this repository's build image (no R there); see INTEGRATION.md.

# ---- R/RcppExports.R:10-46 ---------------------------------------------------------------
sumsmoothlone <- function(x, lambda, alpha = 1e6) {
  .Call(`_cocons_sumsmoothlone`, x, lambda, alpha)
}

cov_rns <- function(theta, locs, x_covariates, smooth_limits) {
  .Call(`_cocons_cov_rns`, theta, locs, x_covariates, smooth_limits)
}

cov_rns_pred <- function(theta, locs, locs_pred, x_covariates, x_covariates_pred, smooth_limits) {
  .Call(`_cocons_cov_rns_pred`, theta, locs, locs_pred, x_covariates, x_covariates_pred, smooth_limits)
}

cov_rns_classic <- function(theta, locs, x_covariates) {
  .Call(`_cocons_cov_rns_classic`, theta, locs, x_covariates)
}

# ---- fused objectives --------------------------------------------------------------------
# kind: 0 ML, 1 profile, 2 REML (include/cocons_b200.h).  The device returns
# c(status, logdet, logdet_w, rank, quad_1..quad_r); n*log(2*pi), the penalty and the `safe`
# logic stay here, exactly as in R/neg2loglikelihood.R.
.cocons.n2ll.device <- function(kind, theta_list, locs, x_covariates, smooth.limits, z, x_betas = NULL) {
  tryCatch(
    .Call(`_cocons_n2ll_dense`, as.integer(kind), theta_list[-1], locs, as.matrix(x_covariates),
          smooth.limits, as.matrix(z), if (is.null(x_betas)) NULL else as.matrix(x_betas), theta_list$mean),
    error = function(e) e)
}

.cocons.chol.failed <- function(out, safe) {
  # status > 0: the leading minor of that order is not positive definite - what base::chol
  # reports as an error at R/neg2loglikelihood.R:200
  if (inherits(out, "error")) stop(out)
  if (out[1] > 0) {
    if (safe) return(TRUE)
    stop("Cholesky error")
  }
  FALSE
}

# R/neg2loglikelihood.R:183-222
GetNeg2loglikelihood <- function(theta, par.pos, locs, x_covariates, smooth.limits, z, n, lambda, safe = TRUE) {
  theta_list <- cocons::getModelLists(theta = theta, par.pos = par.pos, type = "diff")
  out <- .cocons.n2ll.device(0L, theta_list, locs, x_covariates, smooth.limits, z)
  if (.cocons.chol.failed(out, safe)) return(1e+06)
  logdet <- out[2]
  quad <- out[-(1:4)]
  sum_logliks <- sum(n * log(2 * pi) + 2 * logdet + quad)
  sum_logliks + .cocons.getPen(n * dim(z)[2], lambda, theta_list, smooth.limits)
}

# R/neg2loglikelihood.R:127-165
GetNeg2loglikelihoodProfile <- function(theta, par.pos, locs, x_covariates, smooth.limits, z, n, x_betas,
                                        lambda, safe = TRUE) {
  theta_list <- cocons::getModelLists(theta = theta, par.pos = par.pos, type = "diff")
  out <- .cocons.n2ll.device(1L, theta_list, locs, x_covariates, smooth.limits, z, x_betas)
  if (.cocons.chol.failed(out, safe)) return(1e+06)
  sum_logliks <- sum(n * log(2 * pi) + 2 * out[2] + out[-(1:4)])
  sum_logliks + .cocons.getPen(n * dim(z)[2], lambda, theta_list, smooth.limits)
}

# R/neg2loglikelihood.R:241-291
GetNeg2loglikelihoodREML <- function(theta, par.pos, locs, x_covariates, x_betas, smooth.limits, z, n,
                                     lambda, safe = TRUE) {
  theta_list <- cocons::getModelLists(theta = theta, par.pos = par.pos, type = "diff")
  out <- .cocons.n2ll.device(2L, theta_list, locs, x_covariates, smooth.limits, z)
  if (.cocons.chol.failed(out, safe)) return(1e+06)
  p <- out[4]
  sum_logliks <- sum((n - p) * log(2 * pi) + 2 * out[2] + 2 * out[3] + out[-(1:4)])
  sum_logliks + .cocons.getPen((n - p) * dim(z)[2], lambda, theta_list, smooth.limits)
}

# ---- factor reuse for cocoPredict / cocoSim ----------------------------------------------
# Replace R/predict.R:136-159 by
#   ctx  <- .Call(`_cocons_ctx_new`, coco.object@locs, X_std$std.covs, coco.object@z, 0L)
#   .Call(`_cocons_ctx_factor`, ctx, 0L, adjusted_eff_values[-1], ncol(X_std$std.covs), smooth.limits)
#   pr   <- .Call(`_cocons_ctx_predict`, ctx, newlocs, X_pred_std$std.covs, coco.resid)
#   stochastic_part <- pr[[1]];  rowSums(cov_pred * t(inv_cov)) == pr[[2]]        (:159, :173)
# and R/sim.R:162-172 by
#   draws <- .Call(`_cocons_ctx_sim`, ctx, iiderrors)        # == t(t(iiderrors) %*% cholS)
# with iiderrors still produced by set.seed()/rnorm() in R so that seeds reproduce.
