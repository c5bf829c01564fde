"""CPU-side checks of the product: the C-ABI library loads and exports every symbol the header
declares, the host mirror of the R interface agrees with the oracle's literal restatement, the
device math header agrees with mpmath when compiled for the host, and - without a GPU - every
computing entry point fails loudly instead of falling back."""
import ctypes
import os
import re
import subprocess

import mpmath as mp
import numpy as np
import pytest

import cocons_b200 as cb
from cocons_b200 import _lib
from oracle import rmirror

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "cocons_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cocons_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol():
    L = _lib.lib()
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), "libcocons_b200.so does not export %s" % s
        assert s in _lib.SIGNATURES, "no ctypes signature for %s" % s
    assert sorted(_lib.SIGNATURES) == syms
    assert L.cocons_version() >= 100


def test_library_is_sm100a_dmma_code():
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    if not sass:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in sass
    assert "DMMA.8x8x4" in sass  # FP64 tensor-core trailing update
    assert "UBLKCP" in sass  # bulk-copy (TMA) operand feed of the trailing update
    assert "SYNCS.ARRIVE.TRANS64" in sass  # mbarrier expect_tx / complete_tx pipeline


def test_sumsmoothlone_matches_oracle():
    from oracle import cov
    rng = np.random.default_rng(0)
    x = rng.standard_normal(12) * np.array([1, 1e-5, 1, 1e-6, 0, 1, 1, 1e-4, 1, 1e-3, 1e-7, 2])
    assert cb.sumsmoothlone(x, 0.3) == cov.sumsmoothlone(x, 0.3)
    assert cb.sumsmoothlone([], 0.3) == 0.0


@pytest.mark.skipif(_lib.lib().cocons_device_count() > 0, reason="a GPU is present")
def test_no_cpu_fallback_without_a_device():
    th = {k: np.zeros(1) for k in _lib.ASPECTS}
    with pytest.raises(cb.CoconsError, match="no CPU fallback|no CUDA device"):
        cb.cov_rns(th, np.zeros((3, 2)), np.ones((3, 1)), [0.5, 0.5])
    with pytest.raises(cb.CoconsError):
        cb.DenseLikelihood(np.zeros((3, 2)), np.ones((3, 1)), np.zeros(3))
    with pytest.raises(cb.CoconsError):
        cb.GetNeg2loglikelihood(np.zeros(2), {"mean": 0.0, "std.dev": np.array([True]), "scale": np.array([True]),
                                              "aniso": 0.0, "tilt": 0.0, "smooth": 0.5, "nugget": -np.inf},
                                np.zeros((3, 2)), np.ones((3, 1)), [0.5, 0.5], np.zeros(3), 3, (0, 0, 0))


def test_model_lists_scale_and_penalty_match_the_literal_restatement():
    rng = np.random.default_rng(5)
    par_pos = {"mean": np.array([True, True, False, True]), "std.dev": np.array([True, False, True, True]),
               "scale": np.array([True, True, True, False]), "aniso": np.array([True, False, False, False]),
               "tilt": 0.0, "smooth": np.array([True, True, True, True]), "nugget": -np.inf}
    k = sum(int(v.sum()) for v in par_pos.values() if isinstance(v, np.ndarray))
    theta = rng.standard_normal(k)
    for typ in ("diff", "classic"):
        a, b = cb.getModelLists(theta, par_pos, typ), rmirror.get_model_lists(theta, par_pos, typ)
        assert list(a) == list(b)
        for key in a:
            assert np.array_equal(a[key], b[key], equal_nan=True)
    X = np.column_stack([np.ones(40), rng.standard_normal((40, 3)) * [1, 5, 0.1] + [0, 3, -2]])
    a, b = cb.getScale(X), rmirror.get_scale(X)
    for key in a:
        assert np.array_equal(a[key], b[key])
    a2 = cb.getScale(X[:7], a["mean.vector"], a["sd.vector"])["std.covs"]
    assert np.array_equal(a2, rmirror.get_scale(X[:7], b["mean.vector"], b["sd.vector"])["std.covs"])
    tl = cb.getModelLists(theta, par_pos, "diff")
    tl["nugget"][0] = -2.0
    from cocons_b200.api import _getPen
    for lam in ((0, 0, 0), (0.1, 0.02, 0.5)):
        assert _getPen(40, lam, tl, [0.5, 2.5]) == rmirror.get_pen(40, lam, tl, [0.5, 2.5])


def test_design_matrix_mirror(datasets):
    H = datasets["holes_training"][:30]
    data = {"x": H[:, 0], "y": H[:, 1], "cov_x": H[:, 2], "cov_y": H[:, 3], "z": H[:, 4]}
    ml = {"mean": "~ 1 + cov_x", "std.dev": "~ 1 + cov_x + cov_y", "scale": "~ 1 + cov_y", "aniso": 0, "tilt": 0,
          "smooth": 1.5, "nugget": -np.inf}
    dm = cb.getDesignMatrix(ml, data)
    assert dm["colnames"] == ["(Intercept)", "cov_x", "cov_y"]
    assert np.array_equal(dm["model.matrix"][:, 1], H[:, 2])
    assert dm["par.pos"]["mean"].tolist() == [True, True, False]
    assert dm["par.pos"]["scale"].tolist() == [True, False, True]
    assert dm["par.pos"]["smooth"] == 1.5 and np.isneginf(dm["par.pos"]["nugget"])
    obj = cb.coco("dense", data, H[:, :2], H[:, 4], ml)
    assert list(obj.model_list) == list(cb.api.DICTIONARY)
    assert obj.info["smooth.limits"].tolist() == [1.5, 1.5]  # R/cocons.R:157-162
    assert obj.z.shape == (30, 1)
    with pytest.raises(ValueError):
        cb.coco("foo", data, H[:, :2], H[:, 4], ml)  # tests/coco_test.R:260-267


def test_device_bessel_header_on_host_against_mpmath(tmp_path):
    """bessel.cuh also compiles as plain C++: check the exact code the kernel inlines."""
    src = tmp_path / "bh.cpp"
    src.write_text('#include "bessel.cuh"\n'
                   'extern "C" double h_k(double nu, double x){ return cocons::bessel_k(nu, x); }\n'
                   'extern "C" double h_m(double nu, double x){ return cocons::matern_corr(nu, x); }\n')
    so = tmp_path / "libbh.so"
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC",
                           "-I" + os.path.join(ROOT, "cocons_b200", "csrc"), str(src), "-o", str(so)])
    lib = ctypes.CDLL(str(so))
    for f in (lib.h_k, lib.h_m):
        f.argtypes, f.restype = [ctypes.c_double] * 2, ctypes.c_double
    mp.mp.dps = 40
    rng = np.random.default_rng(1)
    nus = [0.05, 0.3, 0.5, 1.0, 1.5, 1.9, 2.5, 3.0, 3.7, 5.2] + list(rng.uniform(0.02, 3.5, 10))
    xs = [1e-12, 1e-6, 0.1, 1.0, 2.0, 2.0000001, 5, 16.5, 24.9, 25, 25.1, 50, 300, 705.9] + list(
        np.exp(rng.uniform(np.log(1e-6), np.log(705), 40)))
    worst_k = worst_m = 0.0
    for nu in nus:
        for x in xs:
            ref = mp.besselk(mp.mpf(nu), mp.mpf(x))
            worst_k = max(worst_k, float(abs((mp.mpf(lib.h_k(nu, x)) - ref) / ref)))
            refm = mp.mpf(2) ** (1 - mp.mpf(nu)) / mp.gamma(mp.mpf(nu)) * mp.mpf(x) ** mp.mpf(nu) * ref
            worst_m = max(worst_m, float(abs((mp.mpf(lib.h_m(nu, x)) - refm) / refm)))
    assert worst_k < 1e-14, worst_k
    assert worst_m < 3e-14, worst_m


def test_r_glue_compiles_and_registers_the_reference_entry_points():
    """rglue/cocons_glue.c replaces src/RcppExports.cpp: same .Call names and arities
    (src/RcppExports.cpp:105-113) plus the fused entries; syntax-checked against the stub R API."""
    glue = os.path.join(ROOT, "cocons_b200", "rglue")
    subprocess.check_call(["gcc", "-fsyntax-only", "-Wall", "-Wextra", "-Wno-unused-parameter",
                           "-I" + os.path.join(glue, "stub"), os.path.join(glue, "cocons_glue.c")])
    text = open(os.path.join(glue, "cocons_glue.c")).read()
    for name, nargs in (("_cocons_sumsmoothlone", 3), ("_cocons_cov_rns", 4), ("_cocons_cov_rns_pred", 6),
                        ("_cocons_cov_rns_classic", 3), ("_cocons_cov_rns_taper_pred", 8),
                        ("_cocons_cov_rns_taper", 6)):
        assert '{"%s", (DL_FUNC)&%s, %d}' % (name, name, nargs) in text
    assert "void R_init_cocons(DllInfo* dll)" in text
    # every C-ABI function the glue calls is declared in the public header
    called = set(re.findall(r"\b(cocons_[a-z0-9_]+)\s*\(", text))
    assert called <= set(_header_symbols()) | {"cocons_ctx"}, called - set(_header_symbols())
    rsrc = open(os.path.join(glue, "R", "cocons_b200.R")).read()
    for fn in ("cov_rns <- function(theta, locs, x_covariates, smooth_limits)",
               "cov_rns_pred <- function(theta, locs, locs_pred, x_covariates, x_covariates_pred, smooth_limits)",
               "cov_rns_classic <- function(theta, locs, x_covariates)",
               "GetNeg2loglikelihood <- function(theta, par.pos, locs, x_covariates, smooth.limits, z, n, lambda, "
               "safe = TRUE)"):
        assert fn in rsrc


def test_spam_stand_ins_against_brute_force():
    """nearest_dist (kd-tree) against an O(n m) distance threshold, on random sites incl. duplicates; the
    spam slot holder round-trips through a dense array."""
    rng = np.random.default_rng(12)
    for n, m, delta in ((1, 1, 0.5), (40, 0, 0.3), (75, 31, 0.45), (64, 64, 0.0)):
        x = rng.uniform(-1, 1, (n, 2))
        if n > 3:
            x[3] = x[1]
        y = None if m == 0 else rng.uniform(-1, 1, (m, 2))
        if m > 5:
            y[5] = x[0]
        sp = cb.nearest_dist(x, y, delta=delta)
        yy = x if y is None else y
        d = np.sqrt((x[:, None, 0] - yy[None, :, 0]) ** 2 + (x[:, None, 1] - yy[None, :, 1]) ** 2)
        keep = d <= delta
        assert sp.dimension == (n, yy.shape[0])
        assert sp.rowpointers[0] == 1 and sp.rowpointers[-1] - 1 == keep.sum() == sp.entries.shape[0]
        rows = np.repeat(np.arange(n), np.diff(sp.rowpointers))
        assert np.array_equal(np.argwhere(keep), np.column_stack([rows, sp.colindices - 1]))  # row-major, ascending
        assert np.allclose(sp.entries, d[keep], rtol=0, atol=1e-15)
        dense = cb.cov_wend1(sp, (max(delta, 1e-9), 1)).toarray()
        assert dense.shape == (n, yy.shape[0]) and np.all(dense[~keep] == 0)
        if y is None and delta > 0:
            assert np.all(np.diag(dense) == 1.0) and np.allclose(dense, dense.T)
    assert sp.colindices.dtype == np.int32 and sp.rowpointers.dtype == np.int32


def test_taper_entry_points_validate_the_pattern_before_touching_a_device():
    th = {k: np.zeros(2) for k in _lib.ASPECTS}
    locs, X = np.zeros((3, 2)), np.ones((3, 2))
    with pytest.raises(cb.CoconsError, match="malformed pattern"):
        cb.cov_rns_taper(th, locs, X, [1, 2, 3], [1, 2, 3, 5], [0.5, 2.5])  # rowpointers end beyond nnz + 1
    with pytest.raises(cb.CoconsError, match="malformed pattern"):
        cb.cov_rns_taper(th, locs, X, [1, 2, 3], [0, 1, 2, 3], [0.5, 2.5])  # 0-based rowpointers
    with pytest.raises(cb.CoconsError, match="outside"):
        cb.cov_rns_taper(th, locs, X, [1, 2, 4], [1, 2, 3, 4], [0.5, 2.5])  # column 4 of 3
    with pytest.raises(cb.CoconsError, match="decrease"):
        cb.cov_rns_taper(th, locs, X, [1, 2, 3], [1, 3, 2, 4], [0.5, 2.5])
    with pytest.raises(ValueError, match="only for sparse"):
        cb.getDensityFromDelta(cb.coco("dense", {"a": np.zeros(3)}, locs, np.zeros(3),
                                       {"std.dev": "~ 1", "scale": "~ 1"}), 0.1)
    with pytest.raises(ValueError, match="taper"):
        cb.coco("sparse", {"a": np.zeros(3)}, locs, np.zeros(3), {"std.dev": "~ 1", "scale": "~ 1"})
    with pytest.raises(ValueError, match="dense"):
        cb.coco("dense", {"a": np.zeros(3)}, locs, np.zeros(3), {"std.dev": "~ 1", "scale": "~ 1"},
                info={"taper": cb.cov_wend1, "delta": 0.1})


def test_two_step_model_pruning_mirrors_the_reference(datasets):
    """.cocons.update.coco.first.step (R/checkFunctions.R:515-603) on a fabricated first-step output: small
    coefficients leave their formula and the boundaries; an aspect left with <= 1 coefficient becomes "~1"; a
    lone small intercept stays.  Host logic only - no device call."""
    from cocons_b200.api import _update_coco_first_step
    H = datasets["holes_training"][:40]
    data = {"x": H[:, 0], "y": H[:, 1], "cov_x": H[:, 2], "cov_y": H[:, 3]}
    ml = {"mean": "~ 1 + cov_x", "std.dev": "~ 1 + cov_x + cov_y", "scale": "~ 1 + cov_x + cov_y", "aniso": 0,
          "tilt": 0, "smooth": 1.5, "nugget": "~ 1"}
    obj = cb.coco("dense", data, H[:, :2], H[:, 4], ml, info={"lambda.Sigma": 0.1, "sparse.point": 1e-4})
    #        mean (2)      std.dev (3)        scale (3)          nugget (1)
    par = np.array([0.3, 2e-5, 0.5, 1e-6, 0.2, 1e-7, 3e-5, 4e-5, 5e-5])
    bounds = {"theta_init": np.arange(9.0), "theta_lower": np.arange(9.0) - 10, "theta_upper": np.arange(9.0) + 10}
    pen = _update_coco_first_step(obj, {"par": par}, bounds)
    # getEstims works on the 'diff' image: std.dev = (a + b) / 2, scale = (a - b) / 2 for shared columns
    est = cb.getModelLists(par, cb.getDesignMatrix(ml, data)["par.pos"], "diff")
    assert abs(est["std.dev"][1]) <= 1e-4 and abs(est["scale"][1]) <= 1e-4  # cov_x is small in both
    assert pen.model_list["mean"] == "~1"            # 1 of 2 small -> n - 1 -> "~1"
    assert pen.model_list["std.dev"] == "~1 + cov_y"  # only cov_x small
    assert pen.model_list["nugget"].replace(" ", "") == "~1"  # a lone small intercept stays an intercept
    assert pen.model_list["aniso"] == 0 and pen.model_list["smooth"] == 1.5
    kept = pen.info["boundaries"]["theta_init"]
    assert 1.0 not in kept and 3.0 not in kept and 0.0 in kept and 8.0 in kept
    assert all(len(v) == len(kept) for v in pen.info["boundaries"].values())


def test_trapezoid_band_of_the_device_bessel_against_mpmath(tmp_path):
    """The middle band 2 < x < 18 (most pairs of a dense model) is evaluated by the trapezoidal rule on the
    integral representation (bessel.cuh, bessel_k_trap_scaled): dense grid against 40-digit mpmath, including
    both band edges, nu = 0 and the nu limit at which the code switches back to CF2."""
    src = tmp_path / "bt.cpp"
    src.write_text('#include "bessel.cuh"\n'
                   'extern "C" double h_t(double nu, double x){ return cocons::bessel_k_trap_scaled(nu, x); }\n'
                   'extern "C" double h_e(double y){ return cocons::exp_poly(y); }\n'
                   'extern "C" int h_band(double nu, double x){ return cocons::bessel_band(nu, x); }\n')
    so = tmp_path / "libbt.so"
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC",
                           "-I" + os.path.join(ROOT, "cocons_b200", "csrc"), str(src), "-o", str(so)])
    lib = ctypes.CDLL(str(so))
    lib.h_t.argtypes, lib.h_t.restype = [ctypes.c_double] * 2, ctypes.c_double
    lib.h_band.argtypes, lib.h_band.restype = [ctypes.c_double] * 2, ctypes.c_int
    lib.h_e.argtypes, lib.h_e.restype = [ctypes.c_double], ctypes.c_double
    # the table-driven exponential of the hot loops: within 1.5 ulp of the true value over its whole range
    ys = np.concatenate([-np.exp(np.random.default_rng(3).uniform(np.log(1e-9), np.log(708), 20000)),
                         np.random.default_rng(4).uniform(-50, 40, 5000), [0.0, -0.3465, 0.3466, -707.9, 700.0]])
    worst_e = max(abs(lib.h_e(float(y)) / np.exp(np.longdouble(float(y))) - 1) for y in ys)
    assert worst_e < 3.4e-16, worst_e  # 1.5 ulp
    assert lib.h_e(-709.0) == 0.0
    assert lib.h_band(1.2, 2.0) == 0 and lib.h_band(1.2, 2.0000001) == 3 and lib.h_band(3.0, 24.999) == 3
    assert lib.h_band(1.2, 25.0) == 2 and lib.h_band(6.5, 10.0) == 1 and lib.h_band(3.5, 30.0) == 1
    assert lib.h_band(3.5, 17.9) == 3 and lib.h_band(3.5, 18.0) == 1
    mp.mp.dps = 40
    worst = 0.0
    for x in list(np.linspace(2.0000001, 24.9999, 47)) + [2.5, 3.0, 7.77]:
        for nu in list(np.linspace(0.0, 6.0, 25)) + [0.5, 1.5, 2.5]:
            if lib.h_band(float(nu), float(x)) != 3:
                continue
            ref = mp.besselk(mp.mpf(float(nu)), mp.mpf(float(x))) * mp.e ** mp.mpf(float(x))
            worst = max(worst, float(abs((mp.mpf(lib.h_t(float(nu), float(x))) - ref) / ref)))
    assert worst < 2.5e-15, worst


def test_two_term_logarithm_and_merged_exponent_of_the_matern_factor(tmp_path):
    """(Q/2)^nu e^{-Q} comes from ONE table-driven exponential of nu (H + L) - Q, with ln(Q/2) = H + L from
    log_hl (bessel.cuh).  Checked on the host build of the header against 50-digit mpmath: the logarithm to
    1e-18 absolute over the whole range the assembly feeds it, the exponential with a low argument word to
    1.5 ulp, and the Matern factor per band - tighter than the plain-double route it replaced."""
    src = tmp_path / "bl.cpp"
    src.write_text('#include "bessel.cuh"\n'
                   'extern "C" void h_log(double x, double* H, double* L){ cocons::log_hl(x, *H, *L); }\n'
                   'extern "C" double h_e2(double y, double ylo){ return cocons::exp_poly2(y, ylo); }\n'
                   'extern "C" double h_m(double nu, double x){ return cocons::matern_corr(nu, x); }\n')
    so = tmp_path / "libbl.so"
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC",
                           "-I" + os.path.join(ROOT, "cocons_b200", "csrc"), str(src), "-o", str(so)])
    lib = ctypes.CDLL(str(so))
    dptr = ctypes.POINTER(ctypes.c_double)
    lib.h_log.argtypes = [ctypes.c_double, dptr, dptr]
    lib.h_e2.argtypes, lib.h_e2.restype = [ctypes.c_double] * 2, ctypes.c_double
    lib.h_m.argtypes, lib.h_m.restype = [ctypes.c_double] * 2, ctypes.c_double
    mp.mp.dps = 50
    rng = np.random.default_rng(11)
    xs = np.concatenate([np.exp(rng.uniform(np.log(1e-17), np.log(400), 4000)), 1 + rng.uniform(-0.02, 0.02, 1000),
                         [1.0, 0.5, 2.0, 0.9999999999999999, 1.0000000000000002, 353.0, 1.1e-16]])
    H, L = ctypes.c_double(), ctypes.c_double()
    worst = 0.0
    for x in xs:
        lib.h_log(float(x), ctypes.byref(H), ctypes.byref(L))
        worst = max(worst, float(abs(mp.mpf(H.value) + mp.mpf(L.value) - mp.log(mp.mpf(float(x))))))
        assert abs(L.value) <= 1e-2 * max(abs(H.value), 1e-300) or abs(H.value) < 1e-2
    assert worst < 1e-18, worst
    worst = 0.0
    for y, ylo in zip(rng.uniform(-700, 30, 3000), rng.uniform(-1e-13, 1e-13, 3000)):
        ref = mp.e ** (mp.mpf(float(y)) + mp.mpf(float(ylo)))
        worst = max(worst, float(abs((mp.mpf(lib.h_e2(float(y), float(ylo))) - ref) / ref)))
    assert worst < 3.4e-16, worst
    assert lib.h_e2(-709.0, 0.0) == 0.0
    nus = list(rng.uniform(0.25, 2.6, 12)) + [0.5, 1.5, 2.5]
    for lo, hi, bar in ((2.0001, 25.0, 1.5e-15), (25.0, 100.0, 1.5e-15), (100.0, 705.0, 1.5e-15)):
        worst = 0.0
        for x in np.exp(rng.uniform(np.log(lo), np.log(hi), 25)):
            for nu in nus:
                nu_, x_ = mp.mpf(float(nu)), mp.mpf(float(x))
                ref = mp.mpf(2) ** (1 - nu_) / mp.gamma(nu_) * x_ ** nu_ * mp.besselk(nu_, x_)
                worst = max(worst, float(abs((mp.mpf(lib.h_m(float(nu), float(x))) - ref) / ref)))
        assert worst < bar, (lo, hi, worst)


@pytest.mark.parametrize("n_pad", [128, 256, 2176, 5632, 12032, 50048])
def test_forward_substitution_work_units_cover_the_factor_and_are_issued_in_dependency_order(n_pad):
    """The dataflow forward substitution (csrc/solve.cu, K6b) hands its work units to CTAs through a ticket
    counter; it cannot deadlock, for any number of resident CTAs, iff every unit depends only on units with
    smaller tickets.  Host-side check of the unit table: the chunks of a row tile [0, I) exactly once, the last
    chunk of a row is the one that finishes it, and for every unit the row-finishing units of all rows below
    its end column J1 (the y it consumes) come earlier in the issue order."""
    L = _lib.lib()
    count = L.cocons_debug_solve_units(n_pad, None, 0)
    T = n_pad // 128
    u = np.zeros((count, 4), dtype=np.int32)
    assert L.cocons_debug_solve_units(n_pad, u.ctypes.data_as(_lib._i32p), count) == count
    finisher = {}  # row -> ticket of its row-finishing unit
    covered = {I: [] for I in range(T)}
    for ticket, (I, J0, J1, ch) in enumerate(u):
        assert 0 <= J0 <= J1 <= I < T and J1 - J0 <= 16
        covered[int(I)].append((int(J0), int(J1), int(ch)))
        if J1 == I:
            assert int(I) not in finisher
            finisher[int(I)] = ticket
    assert sorted(finisher) == list(range(T))
    for I, chunks in covered.items():
        chunks.sort()
        assert [c[2] for c in chunks] == list(range(len(chunks)))          # chunk indices 0..k in column order
        edges = [c[0] for c in chunks] + [chunks[-1][1]]
        assert edges[0] == 0 and edges[-1] == I and all(a[1] == b[0] for a, b in zip(chunks, chunks[1:]))
    first_needed = np.array([finisher[r] for r in range(T)])
    for ticket, (I, J0, J1, ch) in enumerate(u):
        if J1 > 0:  # consumes y_0 .. y_{J1 - 1}: all of them finished by earlier tickets
            assert first_needed[:J1].max() < ticket, (ticket, I, J0, J1)
        if J1 == I and ch > 0:  # the finisher also waits for the row's other chunks
            others = [t for t, (I2, _, J12, _) in enumerate(u) if I2 == I and J12 != I] if T <= 48 else None
            if others is not None:
                assert max(others) < ticket


def test_trailing_update_tile_numbering_is_a_bijection(emu):
    """csrc/chol.cu numbers the tiles of a (lower-trapezoid) update band by band and a CTA decodes its tile in
    O(1) (tile_decode): exhaustive host check over ~19 000 shapes (every ni <= 260 tile rows - eight bands - with
    full and lower-trapezoid column ranges, plus the n = 100 000 / 200 000 shapes) that the decode is a bijection onto
    the expected tile set.  The functions are __host__ __device__; they are taken from the host build of chol.cu
    (tests/host_emul; tools/micro/tile_decode_check.cu is the same loop up to 420 tile rows as an nvcc program)."""
    shapes = ctypes.c_long()
    wrong = emu.emu_tile_decode_check(260, ctypes.byref(shapes))
    assert shapes.value > 15000 and wrong == 0, (shapes.value, wrong)
