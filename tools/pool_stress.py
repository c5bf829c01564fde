"""Stress of concurrent evaluations on one GPU (DenseLikelihoodPool, 8 in flight) on the stripes data set:
every value must equal the single-context value of the same point.  Usage: python tools/pool_stress.py [rounds]"""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
import cocons_b200 as cb
from cocons_b200 import _lib

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 3
which = sys.argv[2] if len(sys.argv) > 2 else "stripes"
size = int(sys.argv[3]) if len(sys.argv) > 3 else 8
D = np.load("tests/golden/datasets.npz")
gold = {c["name"]: c for c in json.load(open("tests/golden/n2ll_cases.json"))["cases"]}
c = gold[which + "_full_general"]
M = D[which + "_training"]
n = c["n"]
X = cb.getScale(np.column_stack([np.ones(n)] + [M[:n, k] for k in ((2, 3, 4) if which == "stripes" else (2, 3))]))["std.covs"]
pp = {k: (np.array(v, dtype=bool) if isinstance(v, list) else v) for k, v in c["par_pos"].items()}
tl = cb.getModelLists(np.array(c["theta"]), pp, "diff")
pts = []
for k in range(32):
    t = {a: v.copy() for a, v in tl.items()}
    t["scale"][0] += 1e-4 * k
    pts.append(t)


def f(ctx, t):
    try:
        v = ctx.terms(_lib.ML, t, c["limits"], t["mean"])["logdet"]
    except cb.NotPositiveDefinite as e:
        v = float("nan") + 0 * e.k
    cs = np.zeros(2)
    _lib.lib().cocons_ctx_debug_checksums(ctx._h, _lib.ptr(cs))
    CHECK[id(t)] = (cs[0], cs[1])
    return v


CHECK = {}


with cb.DenseLikelihood(M[:n, :2], X, M[:n, -1]) as ctx:
    ref = [f(ctx, t) for t in pts]
ref_check = {i: CHECK[id(t)] for i, t in enumerate(pts)}
print("single-context values finite:", int(np.sum(np.isfinite(ref))), "of", len(ref), flush=True)
bad = npd = 0
worst = 0.0
with cb.DenseLikelihoodPool(M[:n, :2], X, M[:n, -1], size=size) as pool:
    for r in range(rounds):
        vals = pool.map(f, pts)
        rel = [abs(a - b) / abs(b) if np.isfinite(a) else float("inf") for a, b in zip(vals, ref)]
        nb = sum(1 for x in rel if x > 1e-13)
        npd += sum(1 for a in vals if not np.isfinite(a))
        worst = max(worst, max(x for x in rel if np.isfinite(x)))
        bad += nb
        for i, (t, x) in enumerate(zip(pts, rel)):
            if x > 1e-13:
                a, b = CHECK[id(t)], ref_check[i]
                print("   point %2d: rel %.2e  assembly checksum %s  factor checksum %s" % (
                    i, x, "same" if a[0] == b[0] else "DIFFERS %.3e" % (abs(a[0] - b[0]) / abs(b[0])),
                    "same" if a[1] == b[1] else "differs %.3e" % (abs(a[1] - b[1]) / abs(b[1]))), flush=True)
        print("round %d: %d of %d beyond 1e-13 (not-PD reports: %d), max finite rel diff %.2e" % (
            r, nb, len(pts), sum(1 for a in vals if not np.isfinite(a)), max(x for x in rel if np.isfinite(x))), flush=True)
print("POOL_STRESS", which, "in_flight", size, "beyond_1e-13:", bad, "not_pd:", npd, "worst_rel:", worst)
