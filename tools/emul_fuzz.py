"""Differential fuzz of the product's API on the HOST BUILD of its own sources (tests/host_emul), against the literal
restatement of the reference (oracle/rmirror.py): random shapes around the tile / chunk / right-hand-side boundaries
(n, p, r, q, m, k), every objective, prediction, marginal and conditional draws.  Run it under the sanitizers with

    COCONS_EMUL_SANITIZE=1 LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" \
        ASAN_OPTIONS=detect_leaks=0 python tools/emul_fuzz.py [cases] [seed] [--dist]

to turn an out-of-bounds access at an odd shape into a report with the kernel's source line.  Needs no GPU; not part of
the product."""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cocons_b200 as cb  # noqa: E402
from cocons_b200 import _lib  # noqa: E402
from host_emul import build  # noqa: E402
from oracle import cov, rmirror  # noqa: E402


def bind():
    lib = build.build(tempfile.mkdtemp())[0]
    for name, (res, args) in _lib.SIGNATURES.items():
        f = getattr(lib, name)
        f.restype, f.argtypes = res, args
    _lib._lib = lib


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def one_case(rng, k):
    n = int(rng.choice([1, 2, 3, 17, 127, 128, 129, 200, 255, 256, 257, 300, 385]))
    p = int(rng.integers(1, 5))
    r = int(rng.choice([1, 1, 2, 3, 7, 9]))
    m = int(rng.choice([1, 2, 50, 127, 128, 129, 200]))
    kd = int(rng.choice([1, 2, 5, 9]))
    locs = rng.uniform(-1, 1, (n, 2))
    lp = rng.uniform(-1, 1, (m, 2))
    if n > 3 and m > 1:
        lp[1] = locs[2]  # a prediction site on a training site
    X = np.column_stack([np.ones(n), rng.standard_normal((n, p - 1))])
    Xp = np.column_stack([np.ones(m), rng.standard_normal((m, p - 1))])
    z = rng.standard_normal((n, r))
    tl = {a: 0.2 * rng.standard_normal(p) for a in ("mean",) + cov.ASPECTS}
    tl["scale"][0], tl["nugget"][0] = -1.4, -2.0
    lim = [0.5, 2.5]
    pp = {a: np.ones(p, dtype=bool) for a in tl}
    theta = np.concatenate([tl["mean"], tl["std.dev"] + tl["scale"], tl["std.dev"] - tl["scale"], tl["aniso"],
                            tl["tilt"], tl["smooth"], tl["nugget"]])
    tl = cb.getModelLists(theta, pp, "diff")
    lam = (0.0, 0.0, 0.0)
    errs = {}
    errs["ml"] = abs(cb.GetNeg2loglikelihood(theta, pp, locs, X, lim, z, n, lam)
                     / rmirror.neg2loglik(theta, pp, locs, X, lim, z, n, lam) - 1)
    if n > p + 1:
        ppm = dict(pp, mean=np.zeros(p, dtype=bool))
        th = theta[p:]
        errs["profile"] = abs(cb.GetNeg2loglikelihoodProfile(th, ppm, locs, X, lim, z, n, X, lam)
                              / rmirror.neg2loglik_profile(th, ppm, locs, X, lim, z, n, X, lam) - 1)
        zc = rmirror.reml_contrast(X, z)
        errs["reml"] = abs(cb.GetNeg2loglikelihoodREML(th, ppm, locs, X, X, lim, zc, n, lam)
                           / rmirror.neg2loglik_reml(th, ppm, locs, X, X, lim, zc, n, lam) - 1)
    with cb.DenseLikelihood(locs, X, z) as ctx:
        ctx.factor(tl, lim)
        L, perm = ctx.get_factor()
        S = cov.cov_rns(tl, locs, X, lim)
        errs["factor"] = rel(L @ L.T, S[np.ix_(perm, perm)])
        resid = z[:, 0] - X @ tl["mean"]
        sto, expl = ctx.predict(lp, Xp, resid)
        C = cov.cov_rns_pred(tl, locs, lp, X, Xp, lim)
        Si_r = np.linalg.solve(S, resid)
        errs["predict"] = rel(sto, C @ Si_r)
        errs["explained"] = rel(expl, np.einsum("ij,ij->i", C, np.linalg.solve(S, C.T).T))
        eps = rng.standard_normal((n, kd))
        want = np.empty((n, kd))
        want[perm] = L @ eps[perm]
        errs["sim"] = rel(ctx.sim(eps), want)
        if m <= 200:
            epm = rng.standard_normal((m, kd))
            Su = cov.cov_rns(tl, lp, Xp, lim)
            schur = Su - C @ np.linalg.solve(S, C.T)
            w = np.linalg.eigvalsh((schur + schur.T) / 2)
            if w.min() > 1e-8 * w.max():  # coincident sites make the Schur complement singular: the draw is undefined
                errs["sim_cond"] = rel(ctx.sim_cond(lp, Xp, epm), np.linalg.cholesky(schur) @ epm)
    tol = {"ml": 1e-9, "profile": 1e-9, "reml": 1e-9, "factor": 1e-12, "predict": 1e-7, "explained": 1e-7, "sim": 1e-10,
           "sim_cond": 1e-6}
    bad = {a: v for a, v in errs.items() if not v < tol[a]}
    print("case %2d  n=%3d p=%d r=%d m=%3d k=%d  %s%s" % (k, n, p, r, m, kd, " ".join("%s=%.0e" % kv for kv in errs.items()),
                                                      "   <-- " + str(bad) if bad else ""), flush=True)
    return not bad


def dist_cases(rng):
    """the block-cyclic driver (cocons_b200.distributed + csrc/dist.cu, one rank) against the resident context at
    sizes around the 512-wide panel boundary: same logdet / quadratic forms / rank for ML, Profile and REML"""
    from cocons_b200.distributed import DistributedDenseLikelihood
    from host_emul.panel_ops import EmulatedPanelOps
    ok = True
    for n in (1, 2, 100, 129, 511, 512, 513, 640, 700):
        p, r = int(rng.integers(1, 4)), int(rng.choice([1, 2, 5]))
        locs = rng.uniform(-1, 1, (n, 2))
        X = np.column_stack([np.ones(n), rng.standard_normal((n, p - 1))])
        z = rng.standard_normal((n, r))
        tl = {a: 0.2 * rng.standard_normal(p) for a in ("mean",) + cov.ASPECTS}
        tl["scale"][0], tl["nugget"][0] = -1.4, -2.0
        worst = 0.0
        with cb.DenseLikelihood(locs, X, z) as ctx, \
                DistributedDenseLikelihood(locs, X, z, ops=EmulatedPanelOps(_lib.lib(), locs, X, z, 0, 1)) as d:
            kinds = [_lib.ML] + ([_lib.PROFILE, _lib.REML] if n > p + 1 else [])
            if n > p + 1:
                ctx.set_xbetas(X)
                d.set_xbetas(X)
            for kind in kinds:
                a, b = ctx.terms(kind, tl, [0.5, 2.5], tl["mean"]), d.terms(kind, tl, [0.5, 2.5], tl["mean"])
                worst = max(worst, abs(a["logdet"] - b["logdet"]) / max(abs(a["logdet"]), 1e-300),
                            float(np.max(np.abs(a["quad"] - b["quad"]) / np.abs(a["quad"]))),
                            abs(a["logdet_w"] - b["logdet_w"]), float(a["rank"] != b["rank"]))
        print("dist  n=%3d p=%d r=%d  resident vs block-cyclic %.0e" % (n, p, r, worst), flush=True)
        ok = ok and worst < 1e-10
    return ok


def main():
    argv = [a for a in sys.argv[1:] if a != "--dist"]
    cases = int(argv[0]) if len(argv) > 0 else 20
    seed = int(argv[1]) if len(argv) > 1 else 1
    bind()
    rng = np.random.default_rng(seed)
    t0, ok = time.time(), True
    for k in range(cases):
        ok = one_case(rng, k) and ok
    if "--dist" in sys.argv:
        ok = dist_cases(rng) and ok
    _lib.lib().cocons_release_workspace()
    print("%d cases in %.0f s: %s" % (cases, time.time() - t0, "all within tolerance" if ok else "FAILURES"))
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
