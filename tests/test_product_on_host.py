"""The product's own API on the CPU: cocons_b200.api bound (fixture `product_on_host`, tests/conftest.py) to the HOST
BUILD of libcocons_b200.so's sources - capi.cu, assembly.cu, chol.cu, solve.cu, taper.cu compiled by g++ against the
CUDA execution-model stand-in of tests/host_emul, every kernel executed thread for thread.  The checks are the GPU
parity tests themselves (tests/test_gpu_*.py), called here at sizes the emulation finishes in seconds: what `-m gpu`
proves on a B200 for the device build, this proves on any machine for the host orchestration (context layout, Morton
permutation and its inverse, right-hand-side blocks, Gram algebra, QR rank, prediction / simulation on the kept
factor, tapered sinks) and for the kernels' arithmetic and index math.  Test scaffolding: nothing here is reachable
from the product, which has no CPU path (test_host.py::test_no_cpu_fallback_without_a_device)."""
import numpy as np
import pytest

import cocons_b200 as cb
import test_gpu_cov
import test_gpu_n2ll
import test_gpu_predict_sim
import test_gpu_taper
from cocons_b200 import _lib
from oracle import rmirror

TL = test_gpu_predict_sim.TL
LIM = test_gpu_predict_sim.LIM


def test_the_binding_is_the_host_build_and_only_for_the_test(product_on_host):
    assert _lib.lib() is product_on_host
    assert _lib.lib().cocons_version() >= 100 and _lib.lib().cocons_device_count() == 1


def test_outside_the_fixture_the_product_still_has_no_cpu_path():
    if _lib.lib().cocons_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(cb.CoconsError, match="no CPU fallback|no CUDA device"):
        cb.cov_rns({k: np.zeros(1) for k in _lib.ASPECTS}, np.zeros((3, 2)), np.ones((3, 1)), [0.5, 0.5])


# ---- covariance builders (tests/test_gpu_cov.py) ----------------------------------------------------------------------
def test_covariance_entry_points(product_on_host, cov_cases):
    test_gpu_cov.test_golden_cases(cov_cases)
    test_gpu_cov.test_quirks_survive_on_the_device(cov_cases)
    test_gpu_cov.test_against_oracle_on_seeded_inputs(129, 4, 4)
    test_gpu_cov.test_argument_errors()


# ---- objectives (tests/test_gpu_n2ll.py) --------------------------------------------------------------------------------
def _small_case(n2ll_cases, n):
    """a golden case cut down to n sites, values from the literal restatement of the reference (oracle/rmirror.py)"""
    c = dict(n2ll_cases["holes1500_general_pen"])
    c["n"] = n
    return c


def test_the_three_objectives_one_shot_and_resident(product_on_host, n2ll_cases, datasets):
    from conftest import case_design
    c = _small_case(n2ll_cases, 300)
    locs, X, z = case_design(c, datasets)
    n, p, lam = c["n"], c["p"], c["lambda"]
    ppm = dict(c["par_pos"], mean=np.zeros(p, dtype=bool))
    th = c["theta"][p:]
    zc = rmirror.reml_contrast(X, z)
    want = {"ml": rmirror.neg2loglik(c["theta"], c["par_pos"], locs, X, c["limits"], z, n, lam),
            "profile": rmirror.neg2loglik_profile(th, ppm, locs, X, c["limits"], z, n, X, lam),
            "reml": rmirror.neg2loglik_reml(th, ppm, locs, X, X, c["limits"], zc, n, lam)}
    c["values"] = want
    got = test_gpu_n2ll._values(c, locs, X, z)
    assert all(abs(got[k] - want[k]) < 1e-9 * abs(want[k]) for k in want), (got, want)
    with cb.DenseLikelihood(locs, X, z) as ctx:
        again = test_gpu_n2ll._values(c, locs, X, z, ctx=ctx)
        assert again == test_gpu_n2ll._values(c, locs, X, z, ctx=ctx)  # bit for bit
    assert all(abs(again[k] - want[k]) < 1e-9 * abs(want[k]) for k in want)


def test_profile_betas_and_not_positive_definite(product_on_host, n2ll_cases, datasets):
    from conftest import case_design, relerr
    c = _small_case(n2ll_cases, 200)
    locs, X, z = case_design(c, datasets)
    p = c["p"]
    ppm = dict(c["par_pos"], mean=np.zeros(p, dtype=bool))
    tl = cb.getModelLists(c["theta"][p:], ppm, "diff")
    with cb.DenseLikelihood(locs, X, z) as ctx:
        ctx.set_xbetas(X)
        ctx.factor(tl, c["limits"])
        betas = ctx.profile_betas(_lib.PROFILE)
    assert relerr(betas, rmirror.profile_betas(tl, locs, X, c["limits"], X, z)) < 1e-8
    bad = n2ll_cases["holes300_notpd"]
    locs, X, z = case_design(bad, datasets)
    assert cb.GetNeg2loglikelihood(bad["theta"], bad["par_pos"], locs, X, bad["limits"], z, bad["n"], bad["lambda"]) == 1e6
    with pytest.raises(ArithmeticError, match="Cholesky error"):
        cb.GetNeg2loglikelihood(bad["theta"], bad["par_pos"], locs, X, bad["limits"], z, bad["n"], bad["lambda"],
                                safe=False)


def test_factor_reconstructs_sigma(product_on_host, datasets):
    locs, X, z, _, _ = test_gpu_predict_sim._setup(datasets, 300, 10)
    S = rmirror._cov.cov_rns(TL, locs, X, LIM)
    with cb.DenseLikelihood(locs, X, z) as ctx:
        ctx.factor(TL, LIM)
        L, perm = ctx.get_factor()
        rows, pos = ctx.factor_rows(np.array([0, 17, 299]))
    assert np.max(np.abs(L @ L.T - S[np.ix_(perm, perm)])) < 1e-13 * np.max(S)
    assert np.max(np.abs(rows @ rows.T - S[np.ix_([0, 17, 299], [0, 17, 299])])) < 1e-13 * np.max(S)


# ---- cocoPredict / cocoSim on the kept factor (tests/test_gpu_predict_sim.py) ----------------------------------------
def test_prediction_and_simulation_on_the_kept_factor(product_on_host, datasets):
    test_gpu_predict_sim.test_predict_matches_reference_route(datasets, 300, 70)
    test_gpu_predict_sim.test_marginal_simulation_entrywise_in_sorted_order(datasets)


def test_conditional_simulation(product_on_host, datasets):
    test_gpu_predict_sim.test_conditional_simulation(datasets)


def test_the_reference_smoke_sequence(product_on_host, datasets):
    """tests/coco_test.R:16-46 in miniature: coco -> cocoOptim -> getCovMatrix -> cocoPredict -> cocoSim"""
    test_gpu_predict_sim.test_smoke_sequence_of_the_reference_test_script(datasets)


# ---- tapered model (tests/test_gpu_taper.py) ----------------------------------------------------------------------------
def test_tapered_entries_and_objective(product_on_host, taper_cases):
    test_gpu_taper.test_golden_entries(taper_cases)
    test_gpu_taper.test_quirks_survive_on_the_device(taper_cases)


def test_tapered_objectives(product_on_host, taper_cases):
    test_gpu_taper.test_objectives_match_goldens(taper_cases)


def test_tapered_not_positive_definite_and_malformed_patterns(product_on_host, taper_cases):
    test_gpu_taper.test_not_positive_definite_and_malformed_patterns(taper_cases)


def test_sparse_predict(product_on_host, taper_cases, datasets):
    test_gpu_taper.test_sparse_predict_matches_golden(taper_cases, datasets)


# ---- callers (tests/test_gpu_predict_sim.py) ------------------------------------------------------------------------------
def test_hessian_points_evaluated_by_a_pool_of_contexts(product_on_host, datasets):
    """getHessian (R/getFunctions.R:925-1034): its finite-difference points go to a pool of contexts driven from host
    threads.  (The 'pml' and the penalised two-step fits of the same file need hundreds of evaluations: GPU only.)"""
    test_gpu_predict_sim.test_hessian_mirrors_the_reference_scheme(datasets)


def test_more_of_the_gpu_suite_at_small_sizes(product_on_host, datasets):
    for n in (100, 128, 517):
        test_gpu_n2ll.test_factor_reconstructs_sigma(n, datasets)
    test_gpu_predict_sim.test_predict_after_fixed_smoothness_factor(datasets)


def test_measurement_helper_runs(product_on_host):
    """cocons_bench_syrk (the stand-alone trailing-update timing of bench.py / tools): argument checks and one run"""
    import ctypes
    ms = ctypes.c_double(-1.0)
    L = _lib.lib()
    assert L.cocons_bench_syrk(0, 256, 32, 1, ctypes.byref(ms)) == 0 and ms.value >= 0.0
    assert L.cocons_bench_syrk(0, 200, 32, 1, ctypes.byref(ms)) < 0  # n must be a multiple of 128
    assert b"multiple of 128" in L.cocons_last_error()
