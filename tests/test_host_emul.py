"""CPU checks of the SHIPPED assembly kernels, executed thread for thread on the host.

cocons_b200/csrc/assembly.cu and taper.cu are compiled with g++ against the small CUDA execution-model shim in
tests/host_emul/ (launches rewritten mechanically, everything else as it ships) and run against the same goldens
and the same 1e-12 bar as the GPU parity tests (tests/test_gpu_cov.py, tests/test_gpu_taper.py): the reference's
operation order in the pair arithmetic (src/cocons_full.cpp:257-313), the staging of the column sites through
shared memory, the tile / slice / slab index math of the resident and the block-cyclic layouts, the Morton
ordering with the carried caller index of the coincident-pair quirk (:284-286), the CSR sinks of the tapered model.
This is test scaffolding: the product has no CPU path (tests/test_host.py::test_no_cpu_fallback_without_a_device),
and the device build of the same source is what `-m gpu` measures.  What it cannot see is CUDA's own libm
(exp / sin / cos differ from glibc's by an ulp) and anything about memory ordering on the device."""
import ctypes

import numpy as np
import pytest

from conftest import relerr, theta_dict
from host_emul import build as emul_build
from oracle import cov

TOL = 1e-12
DIFF, CLASSIC = 0, 1  # COCONS_PAR_DIFF / COCONS_PAR_CLASSIC (include/cocons_b200.h)


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    lib, barriers, launches = emul_build.build(tmp_path_factory.mktemp("host_emul"))
    lib._barriers, lib._rewritten = barriers, launches
    return lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _theta6(th, p):
    return np.ascontiguousarray(np.stack([np.asarray(th[k], dtype=np.float64).reshape(p) for k in cov.ASPECTS]))


def emu_cov(emu, th, locs, X, limits=None, par=DIFF):
    locs, X = _f(locs), _f(X)
    n, p = X.shape
    t6 = _theta6(th, p)
    out = np.full((n, n), np.nan, order="F")
    lim = None if limits is None else np.asarray(limits, dtype=np.float64)
    emu.emu_cov_square(par, n, p, _p(locs), _p(X), _p(t6), None if lim is None else _p(lim), _p(out))
    return out


def emu_pred(emu, th, locs, locs_pred, X, X_pred, limits):
    locs, X, lp, Xp = _f(locs), _f(X), _f(locs_pred), _f(X_pred)
    (n, p), m = X.shape, Xp.shape[0]
    t6, lim = _theta6(th, p), np.asarray(limits, dtype=np.float64)
    out = np.full((m, n), np.nan, order="F")
    emu.emu_cov_pred(n, m, p, _p(locs), _p(lp), _p(X), _p(Xp), _p(t6), _p(lim), _p(out))
    return out


def _case(emu, case):
    th = theta_dict(case["theta6"])
    if "locs_pred" in case:
        return emu_pred(emu, th, case["locs"], case["locs_pred"], case["X"], case["X_pred"], case["limits"])
    if "limits" in case:
        return emu_cov(emu, th, case["locs"], case["X"], case["limits"])
    return emu_cov(emu, th, case["locs"], case["X"], par=CLASSIC)


def test_the_rewrite_is_mechanical_and_complete(emu):
    """every kernel of the two files is launched through the shim; the three with __syncthreads() get real threads"""
    assert emu._barriers == {"site_stage_kernel": False, "assemble_lower_kernel": True, "assemble_cross_kernel": True,
                             "symmetrize_kernel": True, "taper_site_stage_kernel": False,
                             "taper_entries_kernel": False, "taper_pad_diag_kernel": False}
    assert emu._rewritten == 7


def test_shipped_kernels_reproduce_the_goldens_on_the_host(emu, cov_cases):
    report = {}
    for name, case in cov_cases.items():
        got = _case(emu, case)
        assert got.shape == case["out"].shape and not np.any(np.isnan(got)), name
        report[name] = relerr(got, case["out"])
    bad = {k: v for k, v in report.items() if not v < TOL}
    assert not bad, "relative error above 1e-12: %s (all: %s)" % (bad, report)
    assert emu.emu_barrier_launches() > 0


def test_reference_quirks_in_the_emulated_kernels(emu, cov_cases):
    S = _case(emu, cov_cases["degenerate_nu1_fixed"])
    assert S[3, 50] == S[3, 3] and S[50, 3] == S[3, 3]
    S = _case(emu, cov_cases["general_duplicates"])
    assert S[5, 90] == S[5, 5] and S[90, 5] == S[5, 5] and S[17, 100] == S[17, 17]  # App. B-2
    assert np.array_equal(S, S.T)


@pytest.mark.parametrize("n,p,seed", [(1, 1, 0), (2, 2, 1), (127, 3, 2), (128, 3, 3), (129, 4, 4), (300, 5, 5)])
def test_emulated_kernels_against_oracle_on_seeded_inputs(emu, n, p, seed):
    rng = np.random.default_rng(seed)
    locs = rng.uniform(-1, 1, (n, 2))
    X = np.column_stack([np.ones(n), rng.standard_normal((n, p - 1))])
    th = {k: 0.25 * rng.standard_normal(p) for k in cov.ASPECTS}
    th["scale"][0] = -1.4
    th["nugget"][0] = -3.0
    for lim in ([0.5, 2.5], [0.3, 0.9]):
        got = emu_cov(emu, th, locs, X, lim)
        assert relerr(got, cov.cov_rns(th, locs, X, lim)) < TOL, lim
        assert np.array_equal(got, got.T)
    th2 = dict(th, smooth=np.zeros(p))
    for lim in ([0.5, 0.5], [1.5, 1.5], [2.5, 2.5], [1.0, 1.0]):
        assert relerr(emu_cov(emu, th2, locs, X, lim), cov.cov_rns(th2, locs, X, lim)) < TOL, lim
    assert relerr(emu_cov(emu, th, locs, X, par=CLASSIC), cov.cov_rns_classic(th, locs, X)) < TOL
    m = max(1, n // 3)
    lp = rng.uniform(-1, 1, (m, 2))
    lp[0] = locs[n // 2]
    Xp = np.column_stack([np.ones(m), rng.standard_normal((m, p - 1))])
    got = emu_pred(emu, th, locs, lp, X, Xp, [0.5, 2.5])
    assert relerr(got, cov.cov_rns_pred(th, locs, lp, X, Xp, [0.5, 2.5])) < TOL


def test_column_slices_of_small_problems_give_the_same_bits(emu, cov_cases, monkeypatch):
    """gridDim.z CTAs share a tile at small n (launch_assemble_lower): whatever the slice count, the same entries"""
    case = cov_cases["stripes_general_p4"]
    ref = None
    for slices in ("1", "2", "8", "32"):
        monkeypatch.setenv("COCONS_ASM_SLICES", slices)
        got = _case(emu, case)
        ref = got if ref is None else ref
        assert np.array_equal(got, ref), slices


# ---- the context's layout: Morton order, padding to 128, lower triangle only ---------------------------------------
def _sorted_padded(emu, locs, X):
    n, p = X.shape
    n_pad = (n + 127) // 128 * 128
    perm = np.empty(n, dtype=np.int64)
    emu.emu_morton(n, _p(_f(locs)), _p(perm))
    assert sorted(perm.tolist()) == list(range(n))
    Ls, Xs = np.zeros((n_pad, 2), order="F"), np.zeros((n_pad, p), order="F")
    Ls[:n], Xs[:n] = locs[perm], X[perm]
    orig = np.arange(n_pad, dtype=np.int32)
    orig[:n] = perm
    return n_pad, perm, Ls, Xs, orig


def _ctx_lower(emu, th, locs, X, lim):
    n, p = X.shape
    n_pad, perm, Ls, Xs, orig = _sorted_padded(emu, locs, X)
    A = np.full((n_pad, n_pad), np.nan, order="F")
    emu.emu_ctx_lower(DIFF, n, n_pad, p, _p(Ls), _p(Xs), _p(_theta6(th, p)), _p(np.asarray(lim, dtype=np.float64)),
                      _p(orig), _p(A))
    return n_pad, perm, A


def _duplicated_problem(n, p, seed):
    rng = np.random.default_rng(seed)
    locs = rng.uniform(-1, 1, (n, 2))
    X = np.column_stack([np.ones(n), rng.standard_normal((n, p - 1))])
    for a, b in ((5, n - 7), (n // 2, 11), (n - 1, 0)):  # coincident sites with different covariates
        locs[a] = locs[b]
    th = {k: 0.25 * rng.standard_normal(p) for k in cov.ASPECTS}
    th["scale"][0], th["nugget"][0] = -1.4, -3.0
    return locs, X, th


def test_context_layout_morton_order_padding_and_the_carried_caller_index(emu):
    n, p = 333, 3
    locs, X, th = _duplicated_problem(n, p, 21)
    lim = [0.5, 2.5]
    ref = cov.cov_rns(th, locs, X, lim)
    n_pad, perm, A = _ctx_lower(emu, th, locs, X, lim)
    assert n_pad == 384
    ii, jj = np.tril_indices(n)
    # entry (s, t), s >= t, of the context's matrix is the reference's entry (perm[s], perm[t]) - including the
    # coincident pairs, whose value belongs to the lower CALLER index (src/cocons_full.cpp:284-286)
    assert relerr(A[ii, jj], ref[perm[ii], perm[jj]]) < TOL
    for a, b in ((5, n - 7), (n // 2, 11), (n - 1, 0)):
        s, t = np.flatnonzero(perm == a)[0], np.flatnonzero(perm == b)[0]
        assert A[max(s, t), min(s, t)] == ref[min(a, b), min(a, b)]
    # padding: identity; whole diagonal tiles are symmetric; tiles above the diagonal are never touched
    pad = A[n:, :]
    assert np.array_equal(pad[:, n:], np.eye(n_pad - n)) and not np.any(pad[:, :n])
    for t in range(n_pad // 128):
        D = A[128 * t:128 * (t + 1), 128 * t:128 * (t + 1)]
        assert np.array_equal(D, D.T)
        assert np.all(np.isnan(A[128 * t:128 * (t + 1), 128 * (t + 1):]))


@pytest.mark.parametrize("world", [2, 3])
def test_block_cyclic_slabs_hold_exactly_the_resident_matrix(emu, world):
    """csrc/dist.cu: 512-wide column panels dealt in a snake over the ranks, every rank assembling its own panels in
    place - in one launch (cyclic slab mode) or one launch per panel - must give, bit for bit, the columns of the
    resident matrix.  n_pad = 1152: three panels, the last one 128 wide."""
    n, p = 1100, 3
    locs, X, th = _duplicated_problem(n, p, 22)
    lim = [0.5, 2.5]
    n_pad, perm, A = _ctx_lower(emu, th, locs, X, lim)
    _, _, Ls, Xs, orig = _sorted_padded(emu, locs, X)
    npanels = (n_pad + 511) // 512
    assert (n_pad, npanels) == (1152, 3)
    t6, limv = _theta6(th, p), np.asarray(lim, dtype=np.float64)
    seen = set()
    for rank in range(world):
        def owner(K):  # snake_owner of csrc/dist.cu
            rnd, pos = divmod(K, world)
            return world - 1 - pos if rnd & 1 else pos
        mine = [K for K in range(npanels) if owner(K) == rank]
        nlocal = max([K // world + 1 for K in mine], default=0)
        for cyclic in (1, 0):
            slab = np.full((n_pad, 512 * max(nlocal, 1)), np.nan, order="F")
            emu.emu_dist_slabs(DIFF, n, n_pad, p, _p(Ls), _p(Xs), _p(t6), _p(limv), _p(orig), world, rank, nlocal,
                               cyclic, _p(slab))
            for K in mine:
                lp, w = K // world, min(512, n_pad - 512 * K)
                got, want = slab[:, 512 * lp:512 * lp + w], A[:, 512 * K:512 * K + w]
                rows = slice(512 * K, n_pad)  # from the panel's diagonal tile down
                assert np.array_equal(got[rows], want[rows], equal_nan=True), (world, rank, K, cyclic)  # NaN: untouched
                seen.add(K)
    assert seen == set(range(npanels))


# ---- tapered model -------------------------------------------------------------------------------------------------
def _emu_taper(emu, case):
    th = theta_dict(case["theta6"])
    locs, X = _f(case["locs"]), _f(case["X"])
    n, p = X.shape
    col = np.ascontiguousarray(case["colindices"], dtype=np.int32)
    row = np.ascontiguousarray(case["rowpointers"], dtype=np.int32)
    out = np.full(len(col), np.nan)
    lim = np.asarray(case["limits"], dtype=np.float64)
    if "locs_pred" in case:
        lp, Xp = _f(case["locs_pred"]), _f(case["X_pred"])
        emu.emu_taper_entries(n, lp.shape[0], p, _p(locs), _p(lp), _p(X), _p(Xp), _p(_theta6(th, p)), _p(lim), _p(col),
                              _p(row), len(col), _p(out))
    else:
        emu.emu_taper_entries(n, 0, p, _p(locs), None, _p(X), None, _p(_theta6(th, p)), _p(lim), _p(col), _p(row),
                              len(col), _p(out))
    return out


def test_shipped_taper_kernels_reproduce_the_goldens_on_the_host(emu, taper_cases):
    report = {}
    for name, case in taper_cases.items():
        if name == "obj":
            continue
        got = _emu_taper(emu, case)
        assert not np.any(np.isnan(got)), name
        report[name] = relerr(got, case["out"])
    bad = {k: v for k, v in report.items() if not v < TOL}
    assert not bad, "relative error above 1e-12: %s (all: %s)" % (bad, report)


def test_dense_sink_of_the_tapered_objective(emu, taper_cases):
    """TS_LOWER: taper[e] * cov[e] scattered onto the pattern's lower triangle in the context's (Morton) order, unit
    diagonal in the padding - what the blocked Cholesky factors for GetNeg2loglikelihoodTaper."""
    case = taper_cases["taper_duplicates"]
    th = theta_dict(case["theta6"])
    locs, X = np.asarray(case["locs"]), np.asarray(case["X"])
    n, p = X.shape
    col = np.ascontiguousarray(case["colindices"], dtype=np.int32)
    row = np.ascontiguousarray(case["rowpointers"], dtype=np.int32)
    n_pad, perm, Ls, Xs, _ = _sorted_padded(emu, locs, X)
    inv = np.empty(n, dtype=np.int32)
    inv[perm] = np.arange(n, dtype=np.int32)
    taper = np.random.default_rng(3).uniform(0.1, 1.0, len(col))
    A = np.zeros((n_pad, n_pad), order="F")
    emu.emu_taper_lower(n, n_pad, p, _p(Ls), _p(Xs), _p(_theta6(th, p)), _p(np.asarray(case["limits"], dtype=float)),
                        _p(col), _p(row), len(col), _p(taper), _p(inv), _p(A))
    want = np.zeros((n_pad, n_pad))
    entries = _emu_taper(emu, case)
    i = np.repeat(np.arange(n), np.diff(row.astype(np.int64)))
    j = col.astype(np.int64) - 1
    keep = i >= j
    s, t = inv[i[keep]].astype(np.int64), inv[j[keep]].astype(np.int64)
    want[np.maximum(s, t), np.minimum(s, t)] = taper[keep] * entries[keep]
    want[np.arange(n, n_pad), np.arange(n, n_pad)] = 1.0
    assert np.array_equal(A, want)
    assert relerr(entries, case["out"]) < TOL
