"""One-off GPU probe: cuBLAS DGEMM ceiling (torch.matmul float64), our DMMA SYRK kernel at a few
shapes, and per-phase times of one objective evaluation at growing n."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import cocons_b200 as cb
from cocons_b200 import _lib

out = {}
dev = torch.device("cuda:0")
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    for _ in range(2):
        torch.matmul(a, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 5
    for _ in range(reps):
        torch.matmul(a, b)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out["cublas_dgemm_%d" % n] = {"ms": ms, "tflops": 2 * n ** 3 / ms / 1e9}
    print("cuBLAS dgemm", n, out["cublas_dgemm_%d" % n], flush=True)
del a, b
ms = _lib.ctypes.c_double()
for n, k in ((8192, 128), (8192, 512), (16384, 512), (32768, 512)):
    _lib.check(_lib.lib().cocons_bench_syrk(0, n, k, 3, _lib.ctypes.byref(ms)))
    flops = (n / 128) * (n / 128 + 1) / 2 * 2 * 128 * 128 * k
    out["syrk_%d_%d" % (n, k)] = {"ms": ms.value, "tflops": flops / ms.value / 1e9}
    print("our syrk", n, k, out["syrk_%d_%d" % (n, k)], flush=True)

rng = np.random.default_rng(20261018)
for n in (5000, 10000, 20000):
    locs = rng.uniform(-1, 1, (n, 2))
    c1, c2 = (locs[:, 0] + 1) / 2, (locs[:, 1] + 1) / 2
    X = cb.getScale(np.column_stack([np.ones(n), c1, c2, c1 * c2,
                                     0.5 + 0.5 * np.sin(np.pi * locs[:, 0]) * np.cos(np.pi * locs[:, 1])]))["std.covs"]
    z = rng.standard_normal(n)
    tl = {"mean": np.zeros(5), "std.dev": np.array([0.2, 0.15, 0.10, -0.05, 0.05]),
          "scale": np.array([-1.6, 0.2, -0.15, 0.1, -0.1]), "aniso": np.array([0.1, 0.2, -0.1, 0.05, 0]),
          "tilt": np.array([0.3, -0.2, 0.1, 0.1, -0.1]), "smooth": np.array([0.2, 0.3, -0.2, 0.1, 0.1]),
          "nugget": np.array([-4, 0.1, 0.1, 0, 0])}
    with cb.DenseLikelihood(locs, X, z) as ctx:
        for rep in range(2):
            t0 = time.time()
            t = ctx.terms(_lib.ML, tl, [0.5, 2.5], tl["mean"])
            wall = time.time() - t0
        tm = ctx.timings()
    out["eval_%d" % n] = dict(tm, wall_s=wall, chol_tflops=n ** 3 / 3 / tm["factor_ms"] / 1e9,
                              pairs_per_s=n * (n - 1) / 2 / tm["assembly_ms"] * 1e3,
                              value=2 * t["logdet"] + float(t["quad"][0]))
    print("eval", n, out["eval_%d" % n], flush=True)
json.dump(out, open("gpurun_out/probe_fp64.json", "w"), indent=1)
