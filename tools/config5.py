"""BASELINE.json configs[4] and the north_star single-GPU target: one full evaluation at n = 100 000 on
one B200, then cocoPredict (20 000 prediction sites, type "pred") and cocoSim (one marginal draw) on
the kept factor.  Prints timings and size-independent sanity properties."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench
import cocons_b200 as cb
from cocons_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
rng = np.random.default_rng(bench.SEED)
locs_all, X_all, z_all = bench.synthetic(n + m)  # one stream: first n train, next m prediction sites
locs, X, z = locs_all[:n], np.asfortranarray(X_all[:n]), z_all[:n]
lp, Xp = locs_all[n:], np.asfortranarray(X_all[n:])
tl = dict(bench.THETA)
out = {"n": n, "m": m}
with cb.DenseLikelihood(locs, X, z) as ctx:
    for rep in range(2):
        t0 = time.perf_counter()
        t = ctx.terms(_lib.ML, tl, bench.LIMITS, tl["mean"])
        wall = time.perf_counter() - t0
    tm = ctx.timings()
    out["eval"] = dict(tm, wall_s=wall, chol_tflops=bench.flops_chol(n) / tm["factor_ms"] / 1e9,
                       value=n * np.log(2 * np.pi) + 2 * t["logdet"] + float(t["quad"][0]))
    print("EVAL", json.dumps(out["eval"]), flush=True)
    # the objective call keeps the factor: predict and simulate on it
    resid = z - X @ tl["mean"]
    t0 = time.perf_counter()
    sto, expl = ctx.predict(lp, Xp, resid)
    out["predict_s"] = time.perf_counter() - t0
    prior = np.exp(Xp @ tl["std.dev"]) + np.exp(Xp @ tl["nugget"])
    out["predict_checks"] = {"finite": bool(np.all(np.isfinite(sto)) and np.all(np.isfinite(expl))),
                             "explained_between_0_and_prior": bool(np.all(expl >= 0) and np.all(expl <= prior * (1 + 1e-9))),
                             "mean_sd_pred": float(np.mean(np.sqrt(np.abs(prior - expl))))}
    # chunk independence: a 300-site subset alone must reproduce the same numbers
    sub = rng.choice(m, 300, replace=False)
    s2, e2 = ctx.predict(lp[sub], Xp[sub], resid)
    out["predict_checks"]["subset_rel"] = float(max(np.max(np.abs(s2 - sto[sub]) / (np.abs(sto[sub]) + 1e-300)),
                                                    np.max(np.abs(e2 - expl[sub]) / np.abs(expl[sub]))))
    print("PREDICT", out["predict_s"], json.dumps(out["predict_checks"]), flush=True)
    eps = rng.standard_normal((n, 1))
    t0 = time.perf_counter()
    draw = ctx.sim(eps)
    out["sim_s"] = time.perf_counter() - t0
    # |L^-1 draw|^2 = |eps|^2: check through the variance scale instead (cheap): draw has the marginal variances
    out["sim_checks"] = {"finite": bool(np.all(np.isfinite(draw))),
                         "var_ratio": float(np.mean(draw[:, 0] ** 2 / (np.exp(X @ tl["std.dev"]) + np.exp(X @ tl["nugget"]))))}
    print("SIM", out["sim_s"], json.dumps(out["sim_checks"]), flush=True)
json.dump(out, open("gpurun_out/config5.json", "w"), indent=1)
