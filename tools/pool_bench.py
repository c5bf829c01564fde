"""Evaluations/s on the reference's own data sets (holes n = 5570, stripes n = 11 977) with 1, 2, 4, 8
evaluations in flight on one GPU (the optimiser's independent finite-difference points)."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import cocons_b200 as cb
from cocons_b200 import _lib
D = np.load("tests/golden/datasets.npz")
gold = {c["name"]: c for c in json.load(open("tests/golden/n2ll_cases.json"))["cases"]}
for name, key, cols in (("holes_full_general", "holes_training", [2, 3]), ("stripes_full_general", "stripes_training", [2, 3, 4])):
    c = gold[name]
    M = D[key]; n = c["n"]
    X = cb.getScale(np.column_stack([np.ones(n)] + [M[:n, k] for k in cols]))["std.covs"]
    pp = {k: (np.array(v, dtype=bool) if isinstance(v, list) else v) for k, v in c["par_pos"].items()}
    tl = cb.getModelLists(np.array(c["theta"]), pp, "diff")
    z = M[:n, -1]
    pts = []
    for k in range(32):
        t = {a: v.copy() for a, v in tl.items()}
        t["scale"][0] += 1e-4 * k
        pts.append(t)
    for size in (1, 2, 4, 8):
        with cb.DenseLikelihoodPool(M[:n, :2], X, z, size=size) as pool:
            f = lambda ctx, t: ctx.terms(_lib.ML, t, c["limits"], t["mean"])["logdet"]
            pool.map(f, pts[:size])
            t0 = time.perf_counter()
            vals = pool.map(f, pts)
            dt = time.perf_counter() - t0
        print("%s n=%d in_flight=%d: %.1f evals/s (%.2f ms each)  logdet[0]=%.10f" % (name, n, size, len(pts) / dt, 1e3 * dt / len(pts), vals[0]), flush=True)
