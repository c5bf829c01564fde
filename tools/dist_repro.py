"""Run-to-run reproducibility of the distributed driver on ONE rank (all panels local, the per-panel
updates spread over COCONS_DIST_UPD_STREAMS streams): the same evaluation repeated must give bit-identical
terms, and the same terms whatever the number of update streams."""
import sys

import numpy as np

sys.path.insert(0, ".")
import bench
from cocons_b200 import _lib
from cocons_b200.distributed import DistributedDenseLikelihood

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
locs, X, z = bench.synthetic(n)
vals = []
with DistributedDenseLikelihood(locs, X, z) as d:
    for r in range(reps):
        t = d.terms(_lib.ML, bench.THETA, bench.LIMITS, bench.THETA["mean"])
        vals.append((t["logdet"], float(t["quad"][0])))
same = all(v == vals[0] for v in vals)
print("DIST_REPRO n=%d reps=%d identical=%s logdet=%.17g quad=%.17g logdet_spread=%.3e quad_spread=%.3e" % (
    n, reps, same, vals[0][0], vals[0][1], max(abs(v[0] - vals[0][0]) for v in vals) / abs(vals[0][0]),
    max(abs(v[1] - vals[0][1]) for v in vals) / abs(vals[0][1])))
