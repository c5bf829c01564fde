import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench, cocons_b200 as cb
from cocons_b200 import _lib
from cocons_b200.distributed import DistributedDenseLikelihood
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
locs, X, z = bench.synthetic(n)
th = bench.theta_at(0, 0)
th6 = _lib.pack_theta(th, X.shape[1]); lim = np.array(bench.LIMITS); mean = np.ascontiguousarray(th["mean"])
with DistributedDenseLikelihood(locs, X, z) as d:
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        d.ops.assemble(th6, lim, mean)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        d.factor(th6, lim, mean)   # assembles again inside, then factors
        torch.cuda.synchronize(); t2 = time.perf_counter()
        print("rep %d: assemble alone %.1f ms, assemble+factor %.1f ms -> factor %.1f ms = %.2f TFLOP/s" % (
            rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t2 - t1 - (t1 - t0)), bench.flops_chol(n) / (t2 - t1 - (t1 - t0)) / 1e12), flush=True)
with cb.DenseLikelihood(locs, X, z) as ctx:
    for rep in range(2):
        ctx.terms(_lib.ML, th, bench.LIMITS, th["mean"]); print("resident:", ctx.timings())
