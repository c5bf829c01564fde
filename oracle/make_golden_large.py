"""TEST INFRASTRUCTURE ONLY (oracle/): -2 loglik goldens at the sizes the metric lives at.

    python -m oracle.make_golden_large --sites 50000 --points base,w0,t0     # ~10 min per point on 8 cores
    python -m oracle.make_golden_large --sites 20000 --points base

For bench.py's synthetic model (bench.synthetic / bench.THETA / bench.theta_at) the covariance
matrix is produced by the REFERENCE's own compiled source (oracle/_ref/libcocons_ref.so =
/root/reference/src/cocons_full.cpp, cov_rns :40-321) and the objective by LAPACK exactly as
R/neg2loglikelihood.R:183-222 does it (chol -> dpotrf, sum(log(diag)), forwardsolve -> dtrtrs, crossprod).

The reference's pair loop has no threads and no row-range entry, so the n x n matrix is put together
from calls of the UNMODIFIED cov_rns on subsets of the sites: for site blocks a <= b, cov_rns on the
sites of a and b together (in their original relative order, so the roles the pair loop gives the
lower / higher index - kahan operand order, the coincident-site rule :284-286 - are those of the full
call) yields block (b, a) of the full matrix; an entry of cov_rns depends on its two sites and theta
only.  `--check` verifies that claim bit for bit against one direct full call at a smaller n.
The results are appended to tests/golden/n2ll_large.json with their provenance.
"""
import argparse
import json
import mmap
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cov  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "n2ll_large.json")
LIMITS = [0.5, 2.5]

_shared = {}


def point_theta(name):
    """Named evaluation points: 'base' = bench.THETA, 'wK' / 'tK' = bench.theta_at(K, 0) (warm-up / timed
    steps of the default bench count from 0), 'rR' = bench.theta_at(0, R)."""
    import bench
    if name == "base":
        return {k: v.copy() for k, v in bench.THETA.items()}
    if name[0] in "wt":
        return bench.theta_at(int(name[1:]), 0)
    if name[0] == "r":
        return bench.theta_at(0, int(name[1:]))
    raise ValueError(name)


def _block_pair(args):
    a, b = args
    S, n, blocks, theta, locs, X = (_shared[k] for k in ("S", "n", "blocks", "theta", "locs", "X"))
    ia, ib = blocks[a], blocks[b]
    idx = ia if a == b else np.concatenate([ia, ib])  # a < b: already in increasing site order
    sub = cov.cov_rns(theta, locs[idx], X[idx], LIMITS, kind="reference")
    M = np.frombuffer(S, dtype=np.float64).reshape((n, n), order="F")
    if a == b:
        M[np.ix_(ia, ia)] = sub
    else:
        na = len(ia)
        M[np.ix_(ib, ia)] = sub[na:, :na]  # lower block (rows b, columns a)
    return len(idx)


def assemble(theta, locs, X, block, workers):
    """Full covariance (lower triangle valid) in a shared buffer, by the reference's cov_rns on block pairs."""
    n = locs.shape[0]
    S = mmap.mmap(-1, n * n * 8)  # anonymous shared mapping, inherited by the forked workers
    blocks = [np.arange(s, min(n, s + block)) for s in range(0, n, block)]
    _shared.update(S=S, n=n, blocks=blocks, theta=theta, locs=locs, X=X)
    pairs = [(a, b) for a in range(len(blocks)) for b in range(a, len(blocks))]
    pairs.sort(key=lambda ab: ab[0] != ab[1], reverse=True)  # the big jobs first
    with mp.get_context("fork").Pool(workers) as pool:
        for _ in pool.imap_unordered(_block_pair, pairs, chunksize=1):
            pass
    return np.frombuffer(S, dtype=np.float64).reshape((n, n), order="F")


def objective(M, z, mean_vec, X, block=0):
    """R/neg2loglikelihood.R:200-222 on the assembled matrix (in place): n log 2pi + 2 sum log diag + |L^-1 r|^2.

    block == 0: one LAPACK dpotrf + dtrtrs on the whole matrix.  The LP64 OpenBLAS behind scipy faults once the
    matrix holds more than 2^31 elements (n > 46 340), so for larger n the same right-looking factorisation is
    driven block by block (block x block dpotrf, panel dtrsm, trailing update by dgemm on lower blocks, blocked
    forward substitution) - every library call then sees at most block x n elements.  `--check-blocked` compares
    the two routes at n = 20 000."""
    from scipy.linalg import lapack, solve_triangular
    n = M.shape[0]
    resid = z - X @ mean_vec
    if block <= 0:
        c, info = lapack.dpotrf(M, lower=1, overwrite_a=1, clean=0)
        assert info == 0, "dpotrf info=%d" % info
        logdet = float(np.sum(np.log(np.diagonal(c))))
        y, info = lapack.dtrtrs(c, resid, lower=1, trans=0)
        assert info == 0
    else:
        starts = list(range(0, n, block))
        for k0 in starts:
            k1 = min(n, k0 + block)
            D, info = lapack.dpotrf(np.asfortranarray(M[k0:k1, k0:k1]), lower=1, overwrite_a=1, clean=1)
            assert info == 0, "dpotrf info=%d in block at %d" % (info, k0)
            M[k0:k1, k0:k1] = D
            if k1 == n:
                break
            # panel: P = A[k1:, k0:k1] L_kk^-T
            P = solve_triangular(D, np.ascontiguousarray(M[k1:, k0:k1].T), lower=True, check_finite=False).T
            M[k1:, k0:k1] = P
            for i0 in starts:
                if i0 < k1:
                    continue
                i1 = min(n, i0 + block)
                Pi = P[i0 - k1:i1 - k1]
                for j0 in starts:
                    if j0 < k1 or j0 > i0:
                        continue
                    j1 = min(n, j0 + block)
                    M[i0:i1, j0:j1] -= Pi @ P[j0 - k1:j1 - k1].T
        logdet = float(np.sum(np.log(np.diagonal(M))))
        y = np.empty(n)
        for k0 in starts:
            k1 = min(n, k0 + block)
            rhs = resid[k0:k1] - (M[k0:k1, :k0] @ y[:k0] if k0 else 0.0)
            y[k0:k1] = solve_triangular(np.asfortranarray(M[k0:k1, k0:k1]), rhs, lower=True, check_finite=False)
    quad = float(y @ y)
    return {"logdet_half": logdet, "quad": quad, "neg2loglik": n * np.log(2 * np.pi) + 2 * logdet + quad}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sites", type=int, default=50000)
    ap.add_argument("--points", default="base")
    ap.add_argument("--block", type=int, default=2500)
    ap.add_argument("--workers", type=int, default=os.cpu_count())
    ap.add_argument("--check", action="store_true", help="block-pair assembly == one direct cov_rns call (n = 3000)")
    ap.add_argument("--check-blocked", action="store_true", help="blocked LAPACK route == one dpotrf (n = 20 000)")
    ap.add_argument("--sigma-samples", default="", help="comma list of n: covariance of bench.sampled_sites(n) of "
                    "bench.north_star_problem(n) by the reference's cov_rns -> tests/golden/sigma_samples.npz")
    args = ap.parse_args()
    import bench
    if args.check:
        locs, X, z = bench.synthetic(3000)
        th = point_theta("base")
        A = assemble(th, locs, X, 700, args.workers)
        B = cov.cov_rns(th, locs, X, LIMITS, kind="reference")
        il = np.tril_indices(3000)
        same = np.array_equal(A[il], B[il])
        print("block-pair assembly bit-equal to the direct reference call (lower triangle, n=3000):", same)
        sys.exit(0 if same else 1)
    if args.sigma_samples:
        path = os.path.join(ROOT, "tests", "golden", "sigma_samples.npz")
        out = dict(np.load(path)) if os.path.exists(path) else {}
        for n in (int(x) for x in args.sigma_samples.split(",")):
            locs, X, _, _, _, th = bench.north_star_problem(n)
            sites = bench.sampled_sites(n)
            out["sites_n%d" % n] = sites
            out["sigma_n%d" % n] = cov.cov_rns(th, np.asfortranarray(locs[sites]), np.asfortranarray(X[sites]), LIMITS,
                                               kind="reference")
            print("n=%d: %d sites, Sigma diag %.4f..%.4f" % (n, len(sites), np.diag(out["sigma_n%d" % n]).min(),
                                                          np.diag(out["sigma_n%d" % n]).max()))
        np.savez(path, **out)
        sys.exit(0)
    if args.check_blocked:
        locs, X, z = bench.synthetic(20000)
        th = point_theta("base")
        A = assemble(th, locs, X, args.block, args.workers)
        a = objective(A.copy(order="F"), z, th["mean"], X, block=0)
        b = objective(A, z, th["mean"], X, block=4000)
        rel = {k: abs(a[k] - b[k]) / abs(a[k]) for k in a}
        print("blocked vs direct at n=20000:", rel)
        sys.exit(0 if max(rel.values()) < 1e-11 else 1)
    n = args.sites
    locs, X, z = bench.synthetic(n)
    out = {}
    if os.path.exists(GOLD):
        with open(GOLD) as f:
            out = json.load(f)
    out.setdefault("generator", "oracle/make_golden_large.py")
    out.setdefault("covariance", "oracle/_ref/libcocons_ref.so: /root/reference/src/cocons_full.cpp cov_rns (:40-321) compiled "
                   "unmodified, called on site-block pairs (bit-equal to one full call, --check)")
    out.setdefault("algebra", "LAPACK dpotrf / dtrtrs (OpenBLAS via scipy), R/neg2loglikelihood.R:200-222; n > 40 000: the same "
                   "factorisation driven in 4000-wide blocks (dpotrf / dtrsm / dgemm per block), because the LP64 OpenBLAS "
                   "faults beyond 2^31 matrix elements")
    out.setdefault("data", "bench.synthetic(n): numpy PCG64 seed %d; theta = bench.THETA / bench.theta_at" % bench.SEED)
    cases = out.setdefault("cases", {})
    for name in args.points.split(","):
        th = point_theta(name)
        t0 = time.time()
        M = assemble(th, locs, X, args.block, args.workers)
        t1 = time.time()
        res = objective(M, z, th["mean"], X, block=4000 if n > 40000 else 0)
        t2 = time.time()
        del M
        res.update(n=n, point=name, theta={k: [float(x) for x in v] for k, v in th.items()}, assembly_s=round(t1 - t0, 1),
                   lapack_s=round(t2 - t1, 1), cores=args.workers)
        cases["n%d_%s" % (n, name)] = res
        print(name, json.dumps({k: res[k] for k in ("neg2loglik", "logdet_half", "quad", "assembly_s", "lapack_s")}), flush=True)
        with open(GOLD, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
