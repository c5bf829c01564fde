"""CPU checks of the SHIPPED assembly and solve kernels, executed thread for thread on the host.

cocons_b200/csrc/assembly.cu, taper.cu and solve.cu are compiled with g++ against the small CUDA execution-model shim in
tests/host_emul/ (launches rewritten mechanically, everything else as it ships) and run against the same goldens
and the same 1e-12 bar as the GPU parity tests (tests/test_gpu_cov.py, tests/test_gpu_taper.py): the reference's
operation order in the pair arithmetic (src/cocons_full.cpp:257-313), the staging of the column sites through
shared memory, the tile / slice / slab index math of the resident and the block-cyclic layouts, the Morton
ordering with the carried caller index of the coincident-pair quirk (:284-286), the CSR sinks of the tapered model;
the forward substitution (dataflow kernel with its work-unit table, cooperative kernel, two-kernel path), the
log-determinant and the Gram reductions against LAPACK.
This is test scaffolding: the product has no CPU path (tests/test_host.py::test_no_cpu_fallback_without_a_device),
and the device build of the same source is what `-m gpu` measures.  What it cannot see is CUDA's own libm
(exp / sin / cos differ from glibc's by an ulp) and anything about memory ordering on the device."""
import ctypes

import numpy as np
import pytest

from conftest import relerr, theta_dict
from oracle import cov

TOL = 1e-12
DIFF, CLASSIC = 0, 1  # COCONS_PAR_DIFF / COCONS_PAR_CLASSIC (include/cocons_b200.h)


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
    """every kernel of libcocons_b200.so's sources is launched through the shim; those with barriers run on fibers"""
    with_barriers = {k for k, v in emu._barriers.items() if v}
    assert with_barriers == {"assemble_lower_kernel", "assemble_cross_kernel", "symmetrize_kernel", "logdet_kernel",
                             "fwd_solve_coop_kernel", "fwd_tile_solve_kernel", "fwd_tile_update_kernel",
                             "fwd_solve_flow_kernel", "gram_partial_kernel", "gemm_nt_tma_kernel", "potrf_tile_kernel",
                             "potrf_tile_blocked_kernel", "local_logdet_kernel", "acc_update_kernel",
                             "trmm_lower_kernel", "checksum_partial_kernel"}
    assert len(emu._barriers) == 29 and emu._rewritten == 32  # every __global__ / <<<...>>> of csrc/*.cu


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


# ---- forward substitution, log-determinant, Gram reductions (csrc/solve.cu) ----------------------------------------
def _factor_problem(T, seed):
    """a well-conditioned lower factor of T x T tiles and its inverted diagonal tiles (zeros above the diagonal)"""
    import scipy.linalg as sla
    n_pad = 128 * T
    rng = np.random.default_rng(seed)
    M = rng.standard_normal((n_pad, 40))
    L = np.linalg.cholesky(M @ M.T / 40 + np.eye(n_pad) * 2.0)
    W = np.zeros((T, 128, 128))
    for J in range(T):
        W[J] = sla.solve_triangular(L[128 * J:128 * (J + 1), 128 * J:128 * (J + 1)], np.eye(128), lower=True)
    winv = np.ascontiguousarray(np.stack([np.asfortranarray(W[J]).ravel(order="F") for J in range(T)]))
    return n_pad, np.asfortranarray(L), winv


def _emu_solve(emu, n_pad, L, winv, B, mode):
    nrhs = B.shape[1]
    buf = np.zeros((n_pad, 2 * nrhs), order="F")
    buf[:, :nrhs] = B
    assert emu.emu_forward_solve(n_pad, _p(L), _p(winv), _p(buf), nrhs, mode) == 0
    return buf[:, :nrhs].copy()


@pytest.mark.parametrize("T,nrhs", [(1, 1), (3, 2), (5, 3), (18, 1), (18, 5), (35, 11)])
def test_forward_substitution_kernels_on_the_host(emu, T, nrhs):
    """L Y = B by the dataflow kernel (K6b, through the library's own launcher and work-unit table), by the cooperative
    kernel (K6) and by the two-kernels-per-step path, against LAPACK.  T = 18 and 35 give rows whose tile columns are
    cut into two / three chunks (kSolveChunk = 16), nrhs = 11 takes two passes of the launcher (8 + 3)."""
    import scipy.linalg as sla
    n_pad, L, winv = _factor_problem(T, 100 + T)
    B = np.random.default_rng(T).standard_normal((n_pad, nrhs))
    want = sla.solve_triangular(L, B, lower=True)
    scale = np.max(np.abs(want))
    got = {mode: _emu_solve(emu, n_pad, L, winv, B, mode) for mode in ((0, 1, 2) if T <= 18 else (0,))}
    for mode, Y in got.items():
        assert np.max(np.abs(Y - want)) < 1e-12 * scale, (mode, np.max(np.abs(Y - want)) / scale)
    # the same arithmetic in the cooperative kernel and in its two-kernel restatement: the same bits
    if 1 in got:
        assert np.array_equal(got[1], got[2])
    # the dataflow kernel gives the same bits whatever the schedule was: here, twice
    assert np.array_equal(got[0], _emu_solve(emu, n_pad, L, winv, B, 0))


@pytest.mark.parametrize("T,nrhs,blocks", [(18, 1, 4), (35, 3, 8), (35, 11, 3)])
def test_dataflow_substitution_with_truly_concurrent_blocks(emu, T, nrhs, blocks):
    """K6b's blocks talk to each other through global memory only: a ticket counter, the front of finished tile rows,
    per-row counters of partial sums, and y itself read through an 'unset' bit pattern.  Here several blocks run AT
    THE SAME TIME on OS threads (pre-emptive scheduling, every interleaving the kernel is exposed to), and the result
    must be, bit for bit, the one a single block computes alone - "results do not depend on the schedule"
    (csrc/solve.cu) - without a stalled wait (the kernel reports one through its error word instead of hanging)."""
    n_pad, L, winv = _factor_problem(T, 300 + T)
    B = np.random.default_rng(T + nrhs).standard_normal((n_pad, nrhs))
    alone = _emu_solve(emu, n_pad, L, winv, B, 0)
    old = emu.emu_set_concurrent_blocks(blocks)
    try:
        for _ in range(4):
            assert np.array_equal(_emu_solve(emu, n_pad, L, winv, B, 0), alone)
    finally:
        emu.emu_set_concurrent_blocks(old)


def test_logdet_and_gram_kernels_on_the_host(emu):
    n_pad, L, _ = _factor_problem(5, 9)
    n = n_pad - 37
    assert abs(emu.emu_logdet(_p(L), n, n_pad) - np.sum(np.log(np.diag(L)[:n]))) < 1e-12 * n
    Y = np.asfortranarray(np.random.default_rng(4).standard_normal((n_pad, 3)))
    G = np.zeros((3, 3))
    emu.emu_gram(_p(Y), n, n_pad, 3, _p(G))
    want = Y[:n].T @ Y[:n]
    assert np.max(np.abs(G - want)) < 1e-12 * np.max(np.abs(want)) and np.array_equal(G, G.T)


# ---- blocked Cholesky (csrc/chol.cu): DMMA tile kernel, bulk-copy / mbarrier GEMM, look-ahead driver ----------------
def _spd(n_pad, seed, shift=2.0):
    M = np.random.default_rng(seed).standard_normal((n_pad, 40))
    return M @ M.T / 40 + np.eye(n_pad) * shift


def _lower_with_diagonal_tiles(S, fill=np.nan):
    """what the assembly leaves behind: the lower triangle plus whole diagonal tiles; the rest is never read"""
    n_pad = S.shape[0]
    A = np.full((n_pad, n_pad), fill, order="F")
    A[np.tril_indices(n_pad)] = S[np.tril_indices(n_pad)]
    for J in range(n_pad // 128):
        A[128 * J:128 * (J + 1), 128 * J:128 * (J + 1)] = S[128 * J:128 * (J + 1), 128 * J:128 * (J + 1)]
    return A


def test_ptx_helpers_are_the_ones_with_stand_ins(emu):
    """rule (1) of rewrite_ptx: every inline-PTX helper of the kernels is mapped to a documented stand-in"""
    assert emu.ptx_helpers == ["bulk_g2s", "dmma884", "ld_acquire_u32", "ld_relaxed_u64", "mbar_arrive",
                               "mbar_arrive_after", "mbar_expect_tx", "mbar_init", "mbar_wait", "st_release_u32"]


def test_diagonal_tile_kernel_on_the_host(emu):
    """K3b: factor + explicit inverse of a 128 x 128 tile inside a larger matrix; dpotrf's info on a bad pivot"""
    S = _spd(128, 1, shift=1.0)
    big = np.full((300, 200), np.nan, order="F")
    big[50:178, 30:158] = S
    W = np.full((128, 128), np.nan, order="F")
    tile = big[50:, 30:]  # element (0, 0) of the tile, leading dimension 300
    assert emu.emu_potrf_tile(_p(tile), 300, _p(W), 1000) == 0
    L = np.linalg.cholesky(S)
    got = big[50:178, 30:158]
    assert np.max(np.abs(got - L)) < 1e-14 and np.array_equal(np.triu(got, 1), np.zeros((128, 128)))
    assert np.max(np.abs(W - np.linalg.inv(L))) < 1e-13 and np.array_equal(np.triu(W, 1), np.zeros((128, 128)))
    assert np.all(np.isnan(big[:50])) and np.all(np.isnan(big[178:])) and np.all(np.isnan(big[:, :30]))
    bad = S.copy()
    bad[70, 70] = -1.0  # leading minor of order 71 is not positive definite
    A = np.asfortranarray(bad)
    assert emu.emu_potrf_tile(_p(A), 128, _p(W), 1000) == 1000 + 71
    # K3, the register-resident column sweep kept behind COCONS_POTRF=1 (half-warp shuffles, mbarriers in static
    # shared memory, one column ahead): same outputs
    A = np.asfortranarray(S.copy())
    W = np.full((128, 128), np.nan, order="F")
    assert emu.emu_potrf_tile_sweep(_p(A), 128, _p(W), 0) == 0
    assert np.max(np.abs(A - L)) < 1e-14 and np.array_equal(np.triu(A, 1), np.zeros((128, 128)))
    assert np.max(np.abs(W - np.linalg.inv(L))) < 1e-13 and np.array_equal(np.triu(W, 1), np.zeros((128, 128)))
    A = np.asfortranarray(bad)
    assert emu.emu_potrf_tile_sweep(_p(A), 128, _p(W), 1000) == 1000 + 71


@pytest.mark.parametrize("mode,M,N,K,lower", [(0, 384, 256, 48, 0), (0, 384, 384, 144, 1), (1, 256, 128, 128, 0)])
def test_dmma_gemm_with_the_bulk_copy_pipeline_on_the_host(emu, mode, M, N, K, lower):
    """K4 / K5: C (-)= A B^T.  Four-stage ring of bulk copies on mbarriers with a rotating producer warp, fragment
    layout of mma.m8n8k4.f64, tile rasterisation (lower_only: tiles above the diagonal are not touched)."""
    rng = np.random.default_rng(M + N + K)
    A = np.asfortranarray(rng.standard_normal((M, K)))
    B = np.asfortranarray(rng.standard_normal((N, K)))
    C0 = np.asfortranarray(rng.standard_normal((M, N)))
    C = C0.copy(order="F")
    emu.emu_gemm_nt(mode, M, N, K, _p(A), M, _p(B), N, _p(C), M, lower)
    want = A @ B.T if mode == 1 else C0 - A @ B.T
    if lower:  # 128 x 64 tiles whose column block lies above the diagonal row block are skipped
        bi, bj = np.arange(M)[:, None] // 128, np.arange(N)[None, :] // 64
        touched = bi >= bj // 2
        assert np.array_equal(C[~touched], C0[~touched])
        assert np.max(np.abs(C[touched] - want[touched])) < 1e-12
    else:
        assert np.max(np.abs(C - want)) < 1e-12


@pytest.mark.parametrize("T", [1, 2, 5])
def test_cholesky_driver_on_the_host(emu, T):
    """chol_factor(): panel steps (tile kernel, in-place panel solve X = A W^T, in-panel update), trailing updates
    split for the look-ahead, against LAPACK; the strict upper triangle outside the diagonal tiles is never read"""
    n_pad = 128 * T
    S = _spd(n_pad, 40 + T)
    A = _lower_with_diagonal_tiles(S)
    W = np.full((T, 128, 128), np.nan)
    assert emu.emu_chol_factor(n_pad, _p(A), _p(W)) == 0
    L = np.linalg.cholesky(S)
    assert np.max(np.abs(A[np.tril_indices(n_pad)] - L[np.tril_indices(n_pad)])) < 1e-13
    for J in range(T):
        WJ = W[J].T  # stored column-major
        assert np.max(np.abs(WJ - np.linalg.inv(L[128 * J:128 * (J + 1), 128 * J:128 * (J + 1)]))) < 1e-12
    bad = S.copy()
    k = n_pad - 30
    bad[k, k] = -5.0
    A = _lower_with_diagonal_tiles(bad)
    assert emu.emu_chol_factor(n_pad, _p(A), _p(W)) == k + 1  # dpotrf's info: order of the failing leading minor


def test_one_objective_evaluation_by_the_shipped_kernels_on_the_host(emu):
    """The whole hot path of one -2 loglik (R/neg2loglikelihood.R:183-222) - site stage, pairwise assembly in the
    context's layout, blocked Cholesky, log-determinant, forward substitution, Gram reduction - each step by the
    kernel source that ships, executed on the CPU, against the literal restatement of the reference (the same problem
    __graft_entry__.smoke() evaluates on the GPU)."""
    import cocons_b200 as cb
    from oracle import rmirror
    rng = np.random.default_rng(20261018)
    n, p = 400, 3
    locs = rng.uniform(-1, 1, (n, 2))
    X = cb.getScale(np.column_stack([np.ones(n), (locs[:, 0] + 1) / 2, (locs[:, 1] + 1) / 2]))["std.covs"]
    z = rng.standard_normal(n)
    par_pos = {k: np.ones(p, dtype=bool) for k in ("mean", "std.dev", "scale", "aniso", "tilt", "smooth", "nugget")}
    theta = np.array([0.1, 0.3, -0.2, -1.4, 0.35, -0.05, 1.8, -0.05, 0.25, 0.1, 0.2, -0.1, 0.3, -0.2, 0.1, 0.2, 0.3,
                      -0.2, -4, 0.1, 0.1])
    lim = [0.5, 2.5]
    tl = cb.getModelLists(theta, par_pos, "diff")
    n_pad, perm, A = _ctx_lower(emu, tl, locs, X, lim)
    T = n_pad // 128
    W = np.zeros((T, 128, 128))
    assert emu.emu_chol_factor(n_pad, _p(A), _p(W)) == 0
    logdet = emu.emu_logdet(_p(A), n, n_pad)
    rhs = np.zeros((n_pad, 2), order="F")
    rhs[:n, 0] = (z - X @ tl["mean"])[perm]
    assert emu.emu_forward_solve(n_pad, _p(A), _p(W), _p(rhs), 1, 0) == 0
    G = np.zeros((1, 1))
    emu.emu_gram(_p(rhs), n, n_pad, 1, _p(G))
    got = n * np.log(2 * np.pi) + 2 * logdet + G[0, 0]
    want = rmirror.neg2loglik(theta, par_pos, locs, X, lim, z, n, (0.0, 0.0, 0.0))
    assert abs(got - want) < 1e-10 * abs(want), (got, want)


# ---- race check: every CUDA thread a ThreadSanitizer fiber ----------------------------------------------------------
def test_race_check_of_the_shipped_kernels(tmp_path):
    """compute-sanitizer cannot run on the GPU pool, so the kernels' intra-block synchronisation had never been
    race-checked (round-1 review).  Here the host build runs under ThreadSanitizer with every CUDA thread a TSan fiber
    switched WITHOUT implied synchronisation: the only happens-before edges are the kernels' own barriers, mbarrier
    hand-overs and atomics (tests/host_emul/cuda_runtime.h); GEMM, factorisation and assembly run four blocks at a time
    on OS threads, so conflicts BETWEEN the blocks of a launch are seen as well.  One pass over the DMMA GEMM (bulk-copy ring with slot
    re-use, lower-only update, in-place panel product), the blocked Cholesky, the three forward substitutions,
    log-determinant, Gram, the pairwise assembly, then the C ABI on a resident context (REML / ML objectives, prediction,
    marginal and conditional draws, tapered objective) must be silent (tools/emul_racecheck.sh adds the block-cyclic
    path of csrc/dist.cu: silent too, profiles/r02_emul_memcheck_racecheck.md); a deliberately racy kernel and a mutation that
    drops the consumers' hand-back of a ring slot must both be reported."""
    import os
    import subprocess
    from host_emul import build as emul_build
    tsan = subprocess.run(["gcc", "-print-file-name=libtsan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(tsan) or not os.path.exists(tsan):
        pytest.skip("this toolchain has no ThreadSanitizer runtime")
    exe = emul_build.build_racecheck(tmp_path)
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 exitcode=0")
    env.pop("COCONS_EMUL_DROP_HANDBACK", None)
    clean = subprocess.run([exe, "--quick"], env=env, capture_output=True, text=True, timeout=600)
    assert clean.returncode == 0 and "racecheck done" in clean.stdout, clean.stderr[-2000:]
    assert "ThreadSanitizer" not in clean.stderr, clean.stderr[:3000]
    racy = subprocess.run([exe, "--racy"], env=env, capture_output=True, text=True, timeout=600)
    assert "WARNING: ThreadSanitizer: data race" in racy.stderr and "racy_kernel" in racy.stderr
    assert "racy_blocks_kernel" in racy.stderr  # two concurrent blocks storing to one global location
    mutated = subprocess.run([exe, "--gemm"], env=dict(env, COCONS_EMUL_DROP_HANDBACK="1"), capture_output=True,
                             text=True, timeout=600)
    assert "WARNING: ThreadSanitizer: data race" in mutated.stderr
    assert "gemm_nt_tma_kernel" in mutated.stderr and "ptx_bulk_g2s" in mutated.stderr
