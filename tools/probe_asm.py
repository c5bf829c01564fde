import sys
import numpy as np
sys.path.insert(0, ".")
import bench
import cocons_b200 as cb
from cocons_b200 import _lib
n = 20000
locs, X, z = bench.synthetic(n)
with cb.DenseLikelihood(locs, X, z) as ctx:
    for rep in range(3):
        t = ctx.terms(_lib.ML, bench.THETA, bench.LIMITS, bench.THETA["mean"])
    tm = ctx.timings()
    print("assembly_ms %.3f  Gpairs/s %.2f  factor_ms %.2f value %.10f" % (tm["assembly_ms"], n * (n - 1) / 2 / tm["assembly_ms"] / 1e6, tm["factor_ms"], 2 * t["logdet"] + t["quad"][0]))
    # fixed-smoothness fast path
    th = dict(bench.THETA, smooth=np.zeros(5))
    for rep in range(2):
        ctx.terms(_lib.ML, th, [1.5, 1.5], th["mean"])
    print("nu=1.5 closed form: assembly_ms %.3f" % ctx.timings()["assembly_ms"])
