/* .Call glue between the R package and libcocons_b200 - the hand-written replacement for the
 * Rcpp-generated src/RcppExports.cpp of the reference (its wrappers :16-103, the registration
 * table :105-113 and R_init_cocons :115-118).  Plain C on R's C API, no Rcpp.
 *
 * Build inside the R package:  R CMD SHLIB cocons_glue.c -L<dir> -lcocons_b200 -o cocons.so
 * (see INTEGRATION.md).  R is not available in the build image of this repository, so here the
 * file is only syntax-checked against rglue/stub/ (tests/test_host.py); every numeric path it
 * calls is exercised through the same C ABI from Python.
 *
 * Conventions kept from the reference:
 *   - `theta` is a named list; aspects are looked up BY NAME ("std.dev","scale","aniso","tilt",
 *     "smooth","nugget", src/cocons_full.cpp:47-54), extra names such as "mean" are ignored and
 *     a missing name is an error;
 *   - integer matrices / vectors are silently coerced to double, as Rcpp's input_parameter<> does
 *     (src/RcppExports.cpp:34-36);
 *   - results are freshly allocated REALSXP matrices with a `dim` attribute and no dimnames;
 *     inputs are never written to.
 * Errors: a negative status from the library becomes an R error (raised after every temporary
 * has been released); a positive status (matrix not positive definite) is handed back to the R
 * caller, which applies the reference's own `safe` / 1e6 logic (R/neg2loglikelihood.R:200-206).
 * CUDA is initialised lazily on the first call in each process, never at load time, so forked
 * optimParallel workers (R/optim.R:117-121) each get their own context.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>

#include "../../include/cocons_b200.h"

static const char* const kAspects[6] = {"std.dev", "scale", "aniso", "tilt", "smooth", "nugget"};

/* coerce to REALSXP (caller PROTECTs) */
static SEXP as_real(SEXP x) { return TYPEOF(x) == REALSXP ? x : Rf_coerceVector(x, REALSXP); }

static SEXP list_get(SEXP list, const char* name) {
  SEXP names = Rf_getAttrib(list, R_NamesSymbol);
  if (TYPEOF(list) != VECSXP || Rf_isNull(names)) Rf_error("theta must be a named list");
  for (R_xlen_t i = 0; i < XLENGTH(list); ++i)
    if (strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return VECTOR_ELT(list, i);
  Rf_error("Index out of bounds: [index='%s'].", name);
}

/* named list -> 6 x p block (malloc'd; caller frees), all aspects must have length p */
static double* pack_theta(SEXP theta, int p) {
  double* out = (double*)malloc(sizeof(double) * 6 * (size_t)p);
  if (!out) Rf_error("out of memory");
  for (int a = 0; a < 6; ++a) {
    SEXP v = PROTECT(as_real(list_get(theta, kAspects[a])));
    if (LENGTH(v) != p) {
      free(out);
      UNPROTECT(1);
      Rf_error("theta$%s has length %d, expected %d", kAspects[a], LENGTH(v), p);
    }
    memcpy(out + (size_t)a * p, REAL(v), sizeof(double) * (size_t)p);
    UNPROTECT(1);
  }
  return out;
}

static void raise(int status) {
  if (status < 0) Rf_error("cocons_b200: %s", cocons_last_error());
}

/* ---- the reference's six registered entry points -------------------------------------- */

SEXP _cocons_sumsmoothlone(SEXP xS, SEXP lambdaS, SEXP alphaS) {
  SEXP x = PROTECT(as_real(xS));
  double v = cocons_sumsmoothlone(REAL(x), XLENGTH(x), Rf_asReal(lambdaS), Rf_asReal(alphaS));
  UNPROTECT(1);
  return Rf_ScalarReal(v);
}

SEXP _cocons_cov_rns(SEXP thetaS, SEXP locsS, SEXP xS, SEXP limS) {
  SEXP locs = PROTECT(as_real(locsS)), x = PROTECT(as_real(xS)), lim = PROTECT(as_real(limS));
  const int n = Rf_nrows(locs), p = Rf_ncols(x);
  double* th = pack_theta(thetaS, p);
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, n, n));
  int rc = cocons_cov_rns(n, p, REAL(locs), REAL(x), th, REAL(lim), REAL(out));
  free(th);
  UNPROTECT(4);
  raise(rc);
  return out;
}

SEXP _cocons_cov_rns_pred(SEXP thetaS, SEXP locsS, SEXP locsPredS, SEXP xS, SEXP xPredS, SEXP limS) {
  SEXP locs = PROTECT(as_real(locsS)), lp = PROTECT(as_real(locsPredS));
  SEXP x = PROTECT(as_real(xS)), xp = PROTECT(as_real(xPredS)), lim = PROTECT(as_real(limS));
  const int n = Rf_nrows(locs), m = Rf_nrows(lp), p = Rf_ncols(x);
  double* th = pack_theta(thetaS, p);
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, m, n));
  int rc = cocons_cov_rns_pred(n, m, p, REAL(locs), REAL(lp), REAL(x), REAL(xp), th, REAL(lim), REAL(out));
  free(th);
  UNPROTECT(6);
  raise(rc);
  return out;
}

SEXP _cocons_cov_rns_classic(SEXP thetaS, SEXP locsS, SEXP xS) {
  SEXP locs = PROTECT(as_real(locsS)), x = PROTECT(as_real(xS));
  const int n = Rf_nrows(locs), p = Rf_ncols(x);
  double* th = pack_theta(thetaS, p);
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, n, n));
  int rc = cocons_cov_rns_classic(n, p, REAL(locs), REAL(x), th, REAL(out));
  free(th);
  UNPROTECT(3);
  raise(rc);
  return out;
}

/* spam's colindices / rowpointers slots are INTSXP; the reference's Rcpp wrappers accept any numeric
 * vector for them (src/RcppExports.cpp:81-82, 99-100), so doubles are coerced too */
static SEXP as_int(SEXP x) { return TYPEOF(x) == INTSXP ? x : Rf_coerceVector(x, INTSXP); }

SEXP _cocons_cov_rns_taper_pred(SEXP thetaS, SEXP locsS, SEXP locsPredS, SEXP xS, SEXP xPredS, SEXP colS, SEXP rowS,
                                SEXP limS) {
  SEXP locs = PROTECT(as_real(locsS)), lp = PROTECT(as_real(locsPredS));
  SEXP x = PROTECT(as_real(xS)), xp = PROTECT(as_real(xPredS)), lim = PROTECT(as_real(limS));
  SEXP col = PROTECT(as_int(colS)), row = PROTECT(as_int(rowS));
  const int n = Rf_nrows(locs), m = Rf_nrows(lp), p = Rf_ncols(x);
  double* th = pack_theta(thetaS, p);
  SEXP out = PROTECT(Rf_allocVector(REALSXP, XLENGTH(col)));
  int rc = (LENGTH(row) == m + 1)
               ? cocons_cov_rns_taper_pred(n, m, p, REAL(locs), REAL(lp), REAL(x), REAL(xp), th, REAL(lim),
                                           INTEGER(col), INTEGER(row), XLENGTH(col), REAL(out))
               : -100;
  free(th);
  UNPROTECT(8);
  if (rc == -100) Rf_error("cov_rns_taper_pred: rowpointers must have nrow(locs_pred) + 1 elements");
  raise(rc);
  return out;
}

SEXP _cocons_cov_rns_taper(SEXP thetaS, SEXP locsS, SEXP xS, SEXP colS, SEXP rowS, SEXP limS) {
  SEXP locs = PROTECT(as_real(locsS)), x = PROTECT(as_real(xS)), lim = PROTECT(as_real(limS));
  SEXP col = PROTECT(as_int(colS)), row = PROTECT(as_int(rowS));
  const int n = Rf_nrows(locs), p = Rf_ncols(x);
  double* th = pack_theta(thetaS, p);
  SEXP out = PROTECT(Rf_allocVector(REALSXP, XLENGTH(col)));
  int rc = (LENGTH(row) == n + 1) ? cocons_cov_rns_taper(n, p, REAL(locs), REAL(x), th, REAL(lim), INTEGER(col),
                                                         INTEGER(row), XLENGTH(col), REAL(out))
                                  : -100;
  free(th);
  UNPROTECT(6);
  if (rc == -100) Rf_error("cov_rns_taper: rowpointers must have nrow(locs) + 1 elements");
  raise(rc);
  return out;
}

/* ---- fused objective: what GetNeg2loglikelihood{,Profile,REML} call --------------------- */

/* returns c(status, logdet, logdet_w, rank, quad_1..quad_r) */
SEXP _cocons_n2ll_dense(SEXP kindS, SEXP thetaS, SEXP locsS, SEXP xS, SEXP limS, SEXP zS, SEXP xbS, SEXP meanS) {
  SEXP locs = PROTECT(as_real(locsS)), x = PROTECT(as_real(xS)), lim = PROTECT(as_real(limS));
  SEXP z = PROTECT(as_real(zS)), mean = PROTECT(as_real(meanS));
  SEXP xb = PROTECT(Rf_isNull(xbS) ? xbS : as_real(xbS));
  const int n = Rf_nrows(locs), p = Rf_ncols(x);
  const int r = Rf_isMatrix(z) ? Rf_ncols(z) : 1;
  const int q = Rf_isNull(xb) ? 0 : (Rf_isMatrix(xb) ? Rf_ncols(xb) : 1);
  double* th = pack_theta(thetaS, p);
  SEXP out = PROTECT(Rf_allocVector(REALSXP, 4 + r));
  double logdet = R_NaReal, ldw = 0.0;
  int rank = 0;
  int rc = cocons_neg2loglik_dense(Rf_asInteger(kindS), n, p, r, q, REAL(locs), REAL(x), REAL(z),
                                   q ? REAL(xb) : NULL, th, REAL(lim), LENGTH(mean) == p ? REAL(mean) : NULL, &logdet,
                                   REAL(out) + 4, &ldw, &rank);
  free(th);
  REAL(out)[0] = rc, REAL(out)[1] = logdet, REAL(out)[2] = ldw, REAL(out)[3] = rank;
  UNPROTECT(7);
  raise(rc);
  return out;
}

/* ---- device-resident context behind an external pointer --------------------------------- */

static void ctx_finalizer(SEXP ptr) {
  cocons_ctx* c = (cocons_ctx*)R_ExternalPtrAddr(ptr);
  if (c) cocons_ctx_destroy(c);
  R_ClearExternalPtr(ptr);
}

static cocons_ctx* ctx_of(SEXP ptr) {
  cocons_ctx* c = (cocons_ctx*)R_ExternalPtrAddr(ptr);
  if (!c) Rf_error("cocons_b200: the context has been released");
  return c;
}

SEXP _cocons_ctx_new(SEXP locsS, SEXP xS, SEXP zS, SEXP deviceS) {
  SEXP locs = PROTECT(as_real(locsS)), x = PROTECT(as_real(xS)), z = PROTECT(as_real(zS));
  const int n = Rf_nrows(locs), p = Rf_ncols(x), r = Rf_isMatrix(z) ? Rf_ncols(z) : 1;
  cocons_ctx* c = NULL;
  int rc = cocons_ctx_create(Rf_asInteger(deviceS), n, p, r, REAL(locs), REAL(x), REAL(z), NULL, &c);
  UNPROTECT(3);
  raise(rc);
  SEXP ptr = PROTECT(R_MakeExternalPtr(c, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(ptr, ctx_finalizer, TRUE);
  UNPROTECT(1);
  return ptr;
}

SEXP _cocons_ctx_free(SEXP ptr) {
  ctx_finalizer(ptr);
  return R_NilValue;
}

SEXP _cocons_ctx_set_z(SEXP ptr, SEXP zS) {
  SEXP z = PROTECT(as_real(zS));
  int rc = cocons_ctx_set_z(ctx_of(ptr), REAL(z));
  UNPROTECT(1);
  raise(rc);
  return R_NilValue;
}

SEXP _cocons_ctx_set_xbetas(SEXP ptr, SEXP xbS) {
  SEXP xb = PROTECT(as_real(xbS));
  int rc = cocons_ctx_set_xbetas(ctx_of(ptr), Rf_isMatrix(xb) ? Rf_ncols(xb) : 1, REAL(xb));
  UNPROTECT(1);
  raise(rc);
  return R_NilValue;
}

/* c(status, logdet, logdet_w, rank, quad...) on the resident data */
SEXP _cocons_ctx_n2ll(SEXP ptr, SEXP kindS, SEXP thetaS, SEXP pS, SEXP rS, SEXP limS, SEXP meanS) {
  SEXP lim = PROTECT(as_real(limS)), mean = PROTECT(as_real(meanS));
  const int p = Rf_asInteger(pS), r = Rf_asInteger(rS);
  double* th = pack_theta(thetaS, p);
  SEXP out = PROTECT(Rf_allocVector(REALSXP, 4 + r));
  double logdet = R_NaReal, ldw = 0.0;
  int rank = 0;
  int rc = cocons_n2ll(ctx_of(ptr), Rf_asInteger(kindS), th, REAL(lim), LENGTH(mean) == p ? REAL(mean) : NULL, &logdet,
                       REAL(out) + 4, &ldw, &rank);
  free(th);
  REAL(out)[0] = rc, REAL(out)[1] = logdet, REAL(out)[2] = ldw, REAL(out)[3] = rank;
  UNPROTECT(3);
  raise(rc);
  return out;
}

SEXP _cocons_ctx_factor(SEXP ptr, SEXP parS, SEXP thetaS, SEXP pS, SEXP limS) {
  SEXP lim = PROTECT(Rf_isNull(limS) ? limS : as_real(limS));
  double* th = pack_theta(thetaS, Rf_asInteger(pS));
  int rc = cocons_factor(ctx_of(ptr), Rf_asInteger(parS), th, Rf_isNull(lim) ? NULL : REAL(lim));
  free(th);
  UNPROTECT(1);
  raise(rc);
  return Rf_ScalarInteger(rc);
}

SEXP _cocons_ctx_profile_betas(SEXP ptr, SEXP kindS, SEXP qS) {
  SEXP out = PROTECT(Rf_allocVector(REALSXP, Rf_asInteger(qS)));
  int rc = cocons_profile_betas(ctx_of(ptr), Rf_asInteger(kindS), REAL(out));
  UNPROTECT(1);
  raise(rc);
  return out;
}

/* list(stochastic, explained) for cocoPredict (R/predict.R:150-173) */
SEXP _cocons_ctx_predict(SEXP ptr, SEXP locsPredS, SEXP xPredS, SEXP residS) {
  SEXP lp = PROTECT(as_real(locsPredS)), xp = PROTECT(as_real(xPredS)), resid = PROTECT(as_real(residS));
  const int m = Rf_nrows(lp);
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
  SEXP sto = PROTECT(Rf_allocVector(REALSXP, m)), expl = PROTECT(Rf_allocVector(REALSXP, m));
  int rc = cocons_predict(ctx_of(ptr), m, REAL(lp), REAL(xp), REAL(resid), REAL(sto), REAL(expl));
  SET_VECTOR_ELT(out, 0, sto);
  SET_VECTOR_ELT(out, 1, expl);
  UNPROTECT(6);
  raise(rc);
  return out;
}

/* ---- sparse (tapered) model on a resident context ----------------------------------------- */

/* attach ref_taper (its colindices / rowpointers / entries slots, R/optim.R:376-379) */
SEXP _cocons_ctx_set_taper(SEXP ptr, SEXP colS, SEXP rowS, SEXP entriesS) {
  SEXP col = PROTECT(as_int(colS)), row = PROTECT(as_int(rowS)), ent = PROTECT(as_real(entriesS));
  int rc = (XLENGTH(ent) == XLENGTH(col))
               ? cocons_ctx_set_taper(ctx_of(ptr), INTEGER(col), INTEGER(row), REAL(ent), XLENGTH(col))
               : -100;
  UNPROTECT(3);
  if (rc == -100) Rf_error("ctx_set_taper: entries and colindices differ in length");
  raise(rc);
  return R_NilValue;
}

/* c(status, logdet, quad...) of the tapered model (R/neg2loglikelihood.R:20-108) */
SEXP _cocons_ctx_n2ll_taper(SEXP ptr, SEXP thetaS, SEXP pS, SEXP rS, SEXP limS, SEXP meanS) {
  SEXP lim = PROTECT(as_real(limS)), mean = PROTECT(as_real(meanS));
  const int p = Rf_asInteger(pS), r = Rf_asInteger(rS);
  double* th = pack_theta(thetaS, p);
  SEXP out = PROTECT(Rf_allocVector(REALSXP, 2 + r));
  double logdet = R_NaReal;
  int rc = cocons_n2ll_taper(ctx_of(ptr), th, REAL(lim), LENGTH(mean) == p ? REAL(mean) : NULL, &logdet,
                             REAL(out) + 2);
  free(th);
  REAL(out)[0] = rc, REAL(out)[1] = logdet;
  UNPROTECT(3);
  raise(rc);
  return out;
}

SEXP _cocons_ctx_factor_taper(SEXP ptr, SEXP thetaS, SEXP pS, SEXP limS) {
  SEXP lim = PROTECT(as_real(limS));
  double* th = pack_theta(thetaS, Rf_asInteger(pS));
  int rc = cocons_factor_taper(ctx_of(ptr), th, REAL(lim));
  free(th);
  UNPROTECT(1);
  raise(rc);
  return Rf_ScalarInteger(rc);
}

/* list(stochastic, explained) for the sparse cocoPredict (R/predict.R:233-275); the pattern and entries are
 * those of pred_taper BEFORE the covariance is multiplied in */
SEXP _cocons_ctx_predict_taper(SEXP ptr, SEXP locsPredS, SEXP xPredS, SEXP colS, SEXP rowS, SEXP entriesS,
                               SEXP residS) {
  SEXP lp = PROTECT(as_real(locsPredS)), xp = PROTECT(as_real(xPredS)), resid = PROTECT(as_real(residS));
  SEXP col = PROTECT(as_int(colS)), row = PROTECT(as_int(rowS)), ent = PROTECT(as_real(entriesS));
  const int m = Rf_nrows(lp);
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
  SEXP sto = PROTECT(Rf_allocVector(REALSXP, m)), expl = PROTECT(Rf_allocVector(REALSXP, m));
  int rc = (LENGTH(row) == m + 1 && XLENGTH(ent) == XLENGTH(col))
               ? cocons_predict_taper(ctx_of(ptr), m, REAL(lp), REAL(xp), INTEGER(col), INTEGER(row), REAL(ent),
                                      XLENGTH(col), REAL(resid), REAL(sto), REAL(expl))
               : -100;
  SET_VECTOR_ELT(out, 0, sto);
  SET_VECTOR_ELT(out, 1, expl);
  UNPROTECT(9);
  if (rc == -100) Rf_error("ctx_predict_taper: inconsistent pattern");
  raise(rc);
  return out;
}

/* n x k draws L eps for cocoSim (R/sim.R:162-172); eps comes from R's own rnorm */
SEXP _cocons_ctx_sim(SEXP ptr, SEXP epsS) {
  SEXP eps = PROTECT(as_real(epsS));
  const int n = Rf_nrows(eps), k = Rf_ncols(eps);
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, n, k));
  int rc = cocons_sim(ctx_of(ptr), k, REAL(eps), REAL(out));
  UNPROTECT(2);
  raise(rc);
  return out;
}

/* m x k conditional draws (R/sim.R:87-121) */
SEXP _cocons_ctx_sim_cond(SEXP ptr, SEXP locsPredS, SEXP xPredS, SEXP epsS) {
  SEXP lp = PROTECT(as_real(locsPredS)), xp = PROTECT(as_real(xPredS)), eps = PROTECT(as_real(epsS));
  const int m = Rf_nrows(lp), k = Rf_ncols(eps);
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, m, k));
  int rc = cocons_sim_cond(ctx_of(ptr), m, REAL(lp), REAL(xp), k, REAL(eps), REAL(out));
  UNPROTECT(4);
  raise(rc);
  return out;
}

/* ---- registration (replaces src/RcppExports.cpp:105-118) -------------------------------- */

static const R_CallMethodDef CallEntries[] = {
    {"_cocons_sumsmoothlone", (DL_FUNC)&_cocons_sumsmoothlone, 3},
    {"_cocons_cov_rns", (DL_FUNC)&_cocons_cov_rns, 4},
    {"_cocons_cov_rns_pred", (DL_FUNC)&_cocons_cov_rns_pred, 6},
    {"_cocons_cov_rns_classic", (DL_FUNC)&_cocons_cov_rns_classic, 3},
    {"_cocons_cov_rns_taper_pred", (DL_FUNC)&_cocons_cov_rns_taper_pred, 8},
    {"_cocons_cov_rns_taper", (DL_FUNC)&_cocons_cov_rns_taper, 6},
    {"_cocons_n2ll_dense", (DL_FUNC)&_cocons_n2ll_dense, 8},
    {"_cocons_ctx_new", (DL_FUNC)&_cocons_ctx_new, 4},
    {"_cocons_ctx_free", (DL_FUNC)&_cocons_ctx_free, 1},
    {"_cocons_ctx_set_z", (DL_FUNC)&_cocons_ctx_set_z, 2},
    {"_cocons_ctx_set_xbetas", (DL_FUNC)&_cocons_ctx_set_xbetas, 2},
    {"_cocons_ctx_n2ll", (DL_FUNC)&_cocons_ctx_n2ll, 7},
    {"_cocons_ctx_factor", (DL_FUNC)&_cocons_ctx_factor, 5},
    {"_cocons_ctx_profile_betas", (DL_FUNC)&_cocons_ctx_profile_betas, 3},
    {"_cocons_ctx_predict", (DL_FUNC)&_cocons_ctx_predict, 4},
    {"_cocons_ctx_set_taper", (DL_FUNC)&_cocons_ctx_set_taper, 4},
    {"_cocons_ctx_n2ll_taper", (DL_FUNC)&_cocons_ctx_n2ll_taper, 6},
    {"_cocons_ctx_factor_taper", (DL_FUNC)&_cocons_ctx_factor_taper, 4},
    {"_cocons_ctx_predict_taper", (DL_FUNC)&_cocons_ctx_predict_taper, 7},
    {"_cocons_ctx_sim", (DL_FUNC)&_cocons_ctx_sim, 2},
    {"_cocons_ctx_sim_cond", (DL_FUNC)&_cocons_ctx_sim_cond, 4},
    {NULL, NULL, 0}};

void R_init_cocons(DllInfo* dll) {
  R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
