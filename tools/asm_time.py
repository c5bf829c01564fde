"""python tools/asm_time.py [lib.so] [n]: assembly time (CUDA events inside the library) of the bench model at n sites."""
import sys

import numpy as np

sys.path.insert(0, ".")
from cocons_b200 import _lib  # noqa: E402

if len(sys.argv) > 1 and sys.argv[1] != "-":
    _lib.LIB_PATH = sys.argv[1]
import bench  # noqa: E402
import cocons_b200 as cb  # noqa: E402

n = int(sys.argv[2]) if len(sys.argv) > 2 else 30000
locs, X, z = bench.synthetic(n)
with cb.DenseLikelihood(locs, X, z) as ctx:
    ms, vals = [], []
    for k in range(4):
        th = bench.theta_at(k, 0)
        t = ctx.terms(_lib.ML, th, bench.LIMITS, th["mean"])
        ms.append(ctx.timings()["assembly_ms"])
        vals.append(2 * t["logdet"] + float(t["quad"][0]))
    best = min(ms[1:])
    print("ASM_TIME %s n=%d: assembly %.3f ms = %.2f G pairs/s; value[0] %.12e" % (
        sys.argv[1] if len(sys.argv) > 1 else "default", n, best, n * (n - 1) / 2 / best / 1e6, vals[0]))
