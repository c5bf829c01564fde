"""Short GPU program for `ncu --set full`: one SYRK trailing update (n=16384, K=512) and one
objective evaluation at n=8192 (assembly, diagonal-tile, panel kernels)."""
import sys

import numpy as np

sys.path.insert(0, ".")
import cocons_b200 as cb
from cocons_b200 import _lib

ms = _lib.ctypes.c_double()
_lib.check(_lib.lib().cocons_bench_syrk(0, 16384, 512, 1, _lib.ctypes.byref(ms)))
print("syrk 16384x512: %.3f ms" % ms.value)
rng = np.random.default_rng(20261018)
n = 8192
locs = rng.uniform(-1, 1, (n, 2))
c1, c2 = (locs[:, 0] + 1) / 2, (locs[:, 1] + 1) / 2
X = cb.getScale(np.column_stack([np.ones(n), c1, c2, c1 * c2,
                                 0.5 + 0.5 * np.sin(np.pi * locs[:, 0]) * np.cos(np.pi * locs[:, 1])]))["std.covs"]
tl = {"mean": np.zeros(5), "std.dev": np.array([0.2, 0.15, 0.10, -0.05, 0.05]),
      "scale": np.array([-1.6, 0.2, -0.15, 0.1, -0.1]), "aniso": np.array([0.1, 0.2, -0.1, 0.05, 0]),
      "tilt": np.array([0.3, -0.2, 0.1, 0.1, -0.1]), "smooth": np.array([0.2, 0.3, -0.2, 0.1, 0.1]),
      "nugget": np.array([-4, 0.1, 0.1, 0, 0])}
with cb.DenseLikelihood(locs, X, rng.standard_normal(n)) as ctx:
    t = ctx.terms(_lib.ML, tl, [0.5, 2.5], tl["mean"])
    print(ctx.timings(), t["logdet"])
