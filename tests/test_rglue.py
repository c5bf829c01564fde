"""The R boundary, EXECUTED: cocons_b200/rglue/cocons_glue.c (the hand-written replacement for the reference's
src/RcppExports.cpp) is compiled against a miniature of R's C API (tests/rmock/) and driven the way R drives it:
`.Call` by registered name and arity, named lists / matrices with a `dim` attribute in, fresh R objects out, R errors
out of Rf_error().  The miniature collects garbage at EVERY allocation, checks the PROTECT stack on return and counts
the glue's malloc / free - what gctorture / valgrind would check under real R, which this image does not have.

CPU tests: registration table (src/RcppExports.cpp:105-118), argument handling (lookup by name, coercion, shape
errors raised before a device is touched), the one host-only routine end to end, no CPU fallback; and the GPU tests'
.Call sequences with the glue linked against the HOST BUILD of the library's sources (tests/host_emul).
GPU tests (-m gpu): the same calls an R session would make, against the reference-made goldens."""
import numpy as np
import pytest

import cocons_b200 as cb
from cocons_b200 import _lib
from conftest import case_design, relerr, theta_dict
from oracle import cov
from rmock import ExtPtr, RCheckError, RError, RMock

HAVE_GPU = _lib.lib().cocons_device_count() > 0
ASPECTS = ("std.dev", "scale", "aniso", "tilt", "smooth", "nugget")


@pytest.fixture(scope="module")
def R(tmp_path_factory):
    r = RMock(tmp_path_factory.mktemp("rmock"))
    yield r
    r.release_all()


def _theta(p, **over):
    th = {k: np.zeros(p) for k in ASPECTS}
    th.update(over)
    return th


# ---- CPU ------------------------------------------------------------------------------------------------------------
def test_registration_table_is_the_references_plus_the_fused_entries(R):
    reg = R.routines()
    # src/RcppExports.cpp:105-113: the six names and arities an installed cocons resolves through .registration=TRUE
    for name, nargs in (("_cocons_sumsmoothlone", 3), ("_cocons_cov_rns", 4), ("_cocons_cov_rns_pred", 6),
                        ("_cocons_cov_rns_classic", 3), ("_cocons_cov_rns_taper_pred", 8), ("_cocons_cov_rns_taper", 6)):
        assert reg[name] == nargs
    assert {"_cocons_n2ll_dense", "_cocons_ctx_new", "_cocons_ctx_free", "_cocons_ctx_factor", "_cocons_ctx_predict",
            "_cocons_ctx_sim", "_cocons_ctx_sim_cond", "_cocons_ctx_n2ll"} <= set(reg)
    assert R.dynamic_symbols() == 0  # R_useDynamicSymbols(dll, FALSE), :117
    with pytest.raises(RCheckError, match="not available"):
        R.call("_cocons_no_such_routine")
    with pytest.raises(RCheckError, match="Incorrect number of arguments"):
        R.call("_cocons_cov_rns", _theta(1), np.zeros((2, 2)), np.ones((2, 1)))


def test_sumsmoothlone_through_dot_call(R):
    rng = np.random.default_rng(0)
    x = rng.standard_normal(12) * np.array([1, 1e-5, 1, 1e-6, 0, 1, 1, 1e-4, 1, 1e-3, 1e-7, 2])
    got = R.call("_cocons_sumsmoothlone", x, 0.3, 1e6)
    assert got.shape == (1,) and got[0] == cov.sumsmoothlone(x, 0.3)
    # integer input is coerced the way Rcpp's input_parameter<NumericVector> does (src/RcppExports.cpp:18-22)
    xi = np.array([3, -2, 0, 1], dtype=np.int32)
    assert R.call("_cocons_sumsmoothlone", xi, 2, 1e6)[0] == cov.sumsmoothlone(xi.astype(float), 2.0)
    assert R.call("_cocons_sumsmoothlone", np.zeros(0), 0.3, 1e6)[0] == 0.0


def test_theta_is_looked_up_by_name_and_every_error_precedes_the_device(R):
    locs, X, lim = np.zeros((3, 2)), np.ones((3, 2)), [0.5, 2.5]
    bad = _theta(2)
    del bad["tilt"]
    with pytest.raises(RError, match=r"Index out of bounds: \[index='tilt'\]"):  # Rcpp's message for a missing name
        R.call("_cocons_cov_rns", bad, locs, X, lim)
    with pytest.raises(RError, match="named list"):
        R.call("_cocons_cov_rns", np.zeros(12), locs, X, lim)
    with pytest.raises(RError, match=r"theta\$smooth has length 3, expected 2"):
        R.call("_cocons_cov_rns", _theta(2, smooth=np.zeros(3)), locs, X, lim)
    # shapes the reference never checks (Rcpp's operator() is unchecked): here they are R errors, not stray reads
    for args, msg in (
            ((_theta(2), np.zeros((3, 3)), X, lim), "two columns"),
            ((_theta(2), np.zeros(6), X, lim), "two columns"),
            ((_theta(2), locs, np.ones((4, 2)), lim), "nrow"),
            ((_theta(2), locs, X, [0.5]), "smooth_limits"),
    ):
        with pytest.raises(RError, match=msg):
            R.call("_cocons_cov_rns", *args)
    with pytest.raises(RError, match="differ in columns"):
        R.call("_cocons_cov_rns_pred", _theta(2), locs, np.zeros((2, 2)), X, np.ones((2, 3)), lim)
    with pytest.raises(RError, match="rowpointers"):
        R.call("_cocons_cov_rns_taper", _theta(2), locs, X, np.array([1, 2, 3]), np.array([1, 2, 3]), lim)
    with pytest.raises(RError, match="z must have"):
        R.call("_cocons_n2ll_dense", 0, _theta(2), locs, X, lim, np.zeros((4, 1)), None, np.zeros(2))
    with pytest.raises(RError, match="x_betas must have"):
        R.call("_cocons_n2ll_dense", 1, _theta(2), locs, X, lim, np.zeros((3, 1)), np.ones((2, 1)), np.zeros(2))


def test_context_entries_refuse_what_is_not_a_live_context(R):
    for ptr, msg in ((None, "external pointer"), (np.zeros(1), "external pointer"),
                     (ExtPtr(R.lib.rmock_extptr(None)), "released")):
        with pytest.raises(RError, match=msg):
            R.call("_cocons_ctx_sim", ptr, np.zeros((3, 1)))
        with pytest.raises(RError, match=msg):
            R.call("_cocons_ctx_n2ll", ptr, 0, _theta(1), 1, 1, [0.5, 2.5], np.zeros(1))
    assert R.call("_cocons_ctx_free", None) is None  # freeing NULL is harmless, as the on.exit() of the R side needs


@pytest.mark.skipif(HAVE_GPU, reason="a GPU is present")
def test_no_cpu_fallback_behind_the_r_boundary(R):
    th, locs, X = _theta(2), np.array([[0.0, 0], [1, 0], [0, 1]]), np.column_stack([np.ones(3), np.arange(3.0)])
    for name, args in (("_cocons_cov_rns", (th, locs, X, [0.5, 2.5])),
                       ("_cocons_cov_rns_classic", (th, locs, X)),
                       ("_cocons_cov_rns_pred", (th, locs, locs[:2], X, X[:2], [0.5, 2.5])),
                       ("_cocons_n2ll_dense", (0, th, locs, X, [0.5, 2.5], np.zeros((3, 1)), None, np.zeros(2))),
                       ("_cocons_ctx_new", (locs, X, np.zeros((3, 1)), 0))):
        with pytest.raises(RError, match="no CPU fallback|no CUDA device"):
            R.call(name, *args)


def test_the_miniature_runtime_catches_api_misuse(tmp_path):
    """the checks have teeth: a glue with a missing PROTECT, an unbalanced stack and a buffer lost on an error path"""
    bad = tmp_path / "bad_glue.c"
    bad.write_text(r'''
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
SEXP bad_unprotected(SEXP n) {
  SEXP a = Rf_allocVector(REALSXP, 4);          /* not protected ... */
  SEXP b = PROTECT(Rf_allocVector(REALSXP, 4)); /* ... across this allocation */
  REAL(a)[0] = REAL(b)[0] = 1.0;
  UNPROTECT(1);
  return a;
}
SEXP bad_imbalance(SEXP n) { return PROTECT(Rf_allocVector(REALSXP, 1)); }
SEXP bad_leak(SEXP n) {
  double* buf = (double*)malloc(64);
  if (Rf_asInteger(n) > 0) Rf_error("raised while owning a buffer");
  free(buf);
  return R_NilValue;
}
SEXP good(SEXP n) {
  SEXP a = PROTECT(Rf_allocVector(REALSXP, 2)), b = PROTECT(Rf_allocMatrix(REALSXP, 1, 2));
  REAL(a)[0] = 1, REAL(a)[1] = 2, REAL(b)[0] = REAL(a)[0], REAL(b)[1] = REAL(a)[1];
  UNPROTECT(2);
  return b;
}
static const R_CallMethodDef entries[] = {{"bad_unprotected", (DL_FUNC)&bad_unprotected, 1},
                                          {"bad_imbalance", (DL_FUNC)&bad_imbalance, 1},
                                          {"bad_leak", (DL_FUNC)&bad_leak, 1}, {"good", (DL_FUNC)&good, 1}, {NULL, NULL, 0}};
void R_init_cocons(DllInfo* dll) { R_registerRoutines(dll, NULL, entries, NULL, NULL); }
''')
    r = RMock(tmp_path, glue_source=str(bad))
    with pytest.raises(RCheckError, match="not protected across an allocation"):
        r.call("bad_unprotected", 1)
    with pytest.raises(RCheckError, match="stack imbalance"):
        r.call("bad_imbalance", 1)
    with pytest.raises(RCheckError, match="not freed"):
        r.call("bad_leak", 1)
    assert r.call("bad_leak", 0) is None
    assert np.array_equal(r.call("good", 1), np.array([[1.0, 2.0]]))
    r.release_all()


# ---- the same .Call sequences on the CPU: the glue linked against the HOST BUILD of the library (tests/host_emul) -----
@pytest.fixture(scope="module")
def R_host(tmp_path_factory, emu):
    r = RMock(tmp_path_factory.mktemp("rmock_host"), library=emu._name)
    yield r
    r.release_all()


def test_dot_call_sequences_end_to_end_on_the_host_build(R_host, cov_cases, taper_cases, n2ll_cases, datasets):
    """R object -> glue -> C ABI -> kernels -> R object, every layer executed (the kernels by the host emulation):
    the GPU tests below, at the sizes the emulation finishes in seconds"""
    test_cov_entry_points_through_dot_call_match_the_goldens(R_host, cov_cases)
    test_taper_entry_points_through_dot_call_match_the_goldens(R_host, taper_cases)
    test_not_positive_definite_is_a_status_not_an_r_error(R_host, n2ll_cases, datasets)
    test_objectives_through_dot_call_match_the_goldens(R_host, "holes777_ragged", n2ll_cases, datasets)


def test_context_lifecycle_on_the_host_build(R_host, product_on_host, n2ll_cases, datasets):
    test_context_lifecycle_through_dot_call(R_host, n2ll_cases, datasets)


# ---- GPU: what an R session would call --------------------------------------------------------------------------------
gpu = pytest.mark.gpu


def _r_cov(R, case):
    th = theta_dict(case["theta6"])
    th["mean"] = np.zeros(len(th["scale"]))  # getCovMatrix passes the whole list (R/getFunctions.R:48): extra names are fine
    if "locs_pred" in case:
        return R.call("_cocons_cov_rns_pred", th, case["locs"], case["locs_pred"], case["X"], case["X_pred"],
                      case["limits"])
    if "limits" in case:
        return R.call("_cocons_cov_rns", th, case["locs"], case["X"], case["limits"])
    return R.call("_cocons_cov_rns_classic", th, case["locs"], case["X"])


@gpu
def test_cov_entry_points_through_dot_call_match_the_goldens(R, cov_cases):
    report = {}
    for name, case in cov_cases.items():
        got = _r_cov(R, case)
        assert got.shape == case["out"].shape  # a matrix with a dim attribute, as Rcpp's NumericMatrix is
        report[name] = relerr(got, case["out"])
    bad = {k: v for k, v in report.items() if not v < 1e-12}
    assert not bad, bad
    # integer matrices are coerced (src/RcppExports.cpp:34-36)
    th = _theta(2)
    locs_i = np.array([[0, 0], [1, 0], [0, 1]], dtype=np.int32)
    X_i = np.column_stack([np.ones(3), np.arange(3)]).astype(np.int32)
    a = R.call("_cocons_cov_rns", th, locs_i, X_i, [0.5, 0.5])
    b = R.call("_cocons_cov_rns", th, locs_i.astype(float), X_i.astype(float), [0.5, 0.5])
    assert a.shape == (3, 3) and np.array_equal(a, b)


@gpu
def test_taper_entry_points_through_dot_call_match_the_goldens(R, taper_cases):
    for name in ("taper_general", "taper_nu15", "taper_duplicates", "taper_pred_general"):
        case = taper_cases[name]
        th = theta_dict(case["theta6"])
        col, row = case["colindices"].astype(np.int32), case["rowpointers"].astype(np.int32)
        if "locs_pred" in case:
            got = R.call("_cocons_cov_rns_taper_pred", th, case["locs"], case["locs_pred"], case["X"], case["X_pred"],
                         col, row, case["limits"])
        else:
            # the reference's wrappers take the spam slots as doubles too (src/RcppExports.cpp:99-100)
            got = R.call("_cocons_cov_rns_taper", th, case["locs"], case["X"], col.astype(float), row, case["limits"])
        assert got.shape == case["out"].shape and relerr(got, case["out"]) < 1e-12, name


def _r_objective(R, kind, c, locs, X, z, x_betas=None):
    """GetNeg2loglikelihood{,Profile,REML} of cocons_b200/rglue/R/cocons_b200.R, line for line"""
    p, n = c["p"], c["n"]
    pp = dict(c["par_pos"])
    theta = c["theta"]
    if kind != 0:
        pp["mean"] = np.zeros(p, dtype=bool)
        theta = theta[p:]
    tl = cb.getModelLists(theta, pp, "diff")
    theta_minus_mean = {k: tl[k] for k in ASPECTS}  # theta_list[-1]
    out = R.call("_cocons_n2ll_dense", kind, theta_minus_mean, locs, X, c["limits"], z, x_betas, tl["mean"])
    if out[0] > 0:
        return 1e6
    r = z.shape[1]
    if kind == 2:
        rank = out[3]
        return float(np.sum((n - rank) * np.log(2 * np.pi) + 2 * out[1] + 2 * out[2] + out[4:])
                     + cb.api._getPen((n - rank) * r, c["lambda"], tl, c["limits"]))
    return float(np.sum(n * np.log(2 * np.pi) + 2 * out[1] + out[4:]) + cb.api._getPen(n * r, c["lambda"], tl, c["limits"]))


@gpu
@pytest.mark.parametrize("name", ["holes1500_general_pen", "holes777_ragged", "holesbm1000_r10"])
def test_objectives_through_dot_call_match_the_goldens(R, name, n2ll_cases, datasets):
    c = n2ll_cases[name]
    locs, X, z = case_design(c, datasets)
    got = {"ml": _r_objective(R, 0, c, locs, X, z),
           "profile": _r_objective(R, 1, c, locs, X, z, x_betas=X),
           "reml": _r_objective(R, 2, c, locs, X, cb.reml_contrasts(X, z))}
    errs = {k: abs(got[k] - c["values"][k]) / abs(c["values"][k]) for k in got}
    assert all(v < 1e-9 for v in errs.values()), errs


@gpu
def test_not_positive_definite_is_a_status_not_an_r_error(R, n2ll_cases, datasets):
    c = n2ll_cases["holes300_notpd"]
    locs, X, z = case_design(c, datasets)
    assert _r_objective(R, 0, c, locs, X, z) == c["values"]["ml"] == 1e6  # `safe` logic of R/neg2loglikelihood.R:200-206


@gpu
def test_context_lifecycle_through_dot_call(R, n2ll_cases, datasets):
    """the factor-reuse sequence of rglue/R/cocons_b200.R: new -> factor -> predict / sim -> free"""
    c = n2ll_cases["holes777_ragged"]
    locs, X, z = case_design(c, datasets)
    n, p = X.shape
    tl = cb.getModelLists(c["theta"], c["par_pos"], "diff")
    th = {k: tl[k] for k in ASPECTS}
    ctx = R.call("_cocons_ctx_new", locs, X, z, 0)
    assert isinstance(ctx, ExtPtr)
    out = R.call("_cocons_ctx_n2ll", ctx, 0, th, p, 1, c["limits"], tl["mean"])
    v = n * np.log(2 * np.pi) + 2 * out[1] + out[4] + cb.api._getPen(n, c["lambda"], tl, c["limits"])
    assert abs(v - c["values"]["ml"]) < 1e-9 * abs(c["values"]["ml"])
    assert R.call("_cocons_ctx_factor", ctx, 0, th, p, c["limits"])[0] == 0
    resid = z[:, 0] - X @ tl["mean"]
    m = 40
    sto, expl = R.call("_cocons_ctx_predict", ctx, locs[:m], X[:m], resid)
    assert sto.shape == (m,) and expl.shape == (m,)
    assert np.max(np.abs(sto - resid[:m])) < 1e-8 * np.max(np.abs(resid))  # kriging interpolates at training sites
    eps = np.random.default_rng(5).standard_normal((n, 3))
    draws = R.call("_cocons_ctx_sim", ctx, eps)
    assert draws.shape == (n, 3) and np.all(np.isfinite(draws))
    with cb.DenseLikelihood(locs, X, z) as ref:  # the same draws through the Python host mirror
        ref.factor(tl, c["limits"])
        assert np.array_equal(ref.sim(eps), draws)
    for args, msg in (((ctx, np.zeros((n + 1, 1))), "n x k"), ((ctx, np.zeros(n)), "n x k")):
        with pytest.raises(RError, match=msg):
            R.call("_cocons_ctx_sim", *args)
    with pytest.raises(RError, match="n elements"):
        R.call("_cocons_ctx_predict", ctx, locs[:m], X[:m], resid[:-1])
    assert R.call("_cocons_ctx_free", ctx) is None
    with pytest.raises(RError, match="released"):
        R.call("_cocons_ctx_sim", ctx, eps)
