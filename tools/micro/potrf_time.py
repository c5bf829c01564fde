import sys, time
import numpy as np
sys.path.insert(0, ".")
import bench
import cocons_b200 as cb
from cocons_b200 import _lib
for n in (1024, 5570):
    locs, X, z = bench.synthetic(n)
    with cb.DenseLikelihood(locs, X, z) as ctx:
        for rep in range(3):
            t = ctx.terms(_lib.ML, bench.THETA, bench.LIMITS, bench.THETA["mean"])
        print(n, ctx.timings())
