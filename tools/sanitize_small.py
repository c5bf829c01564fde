"""Small end-to-end pass over every kernel, meant to be run under compute-sanitizer."""
import sys
import numpy as np
sys.path.insert(0, ".")
import cocons_b200 as cb
from cocons_b200 import _lib
from cocons_b200.distributed import DistributedDenseLikelihood

rng = np.random.default_rng(3)
n, m, p = 300, 70, 3
locs = rng.uniform(-1, 1, (n, 2))
X = cb.getScale(np.column_stack([np.ones(n), rng.standard_normal((n, p - 1))]))["std.covs"]
lp = rng.uniform(-1, 1, (m, 2))
Xp = np.column_stack([np.ones(m), rng.standard_normal((m, p - 1))])
z = rng.standard_normal((n, 2))
tl = {"mean": np.array([0.1, 0.3, -0.2]), "std.dev": np.array([0.2, 0.15, 0.1]), "scale": np.array([-1.6, 0.2, -0.15]),
      "aniso": np.array([0.1, 0.2, -0.1]), "tilt": np.array([0.3, -0.2, 0.1]), "smooth": np.array([0.2, 0.3, -0.2]),
      "nugget": np.array([-4, 0.1, 0.1])}
lim = [0.5, 2.5]
S = cb.cov_rns(tl, locs, X, lim)
C = cb.cov_rns_pred(tl, locs, lp, X, Xp, lim)
Sc = cb.cov_rns_classic(tl, locs, X)
S15 = cb.cov_rns(dict(tl, smooth=np.zeros(3)), locs, X, [1.5, 1.5])
with cb.DenseLikelihood(locs, X, z) as ctx:
    ctx.set_xbetas(X[:, :2])
    for kind in (_lib.ML, _lib.PROFILE, _lib.REML):
        t = ctx.terms(kind, tl, lim, tl["mean"])
    ctx.factor(tl, lim)
    sto, expl = ctx.predict(lp, Xp, z[:, 0])
    d = ctx.sim(rng.standard_normal((n, 2)))
    dc = ctx.sim_cond(lp, Xp, rng.standard_normal((m, 2)))
    betas = ctx.profile_betas(_lib.PROFILE)
n2 = 700
locs2 = rng.uniform(-1, 1, (n2, 2))
X2 = cb.getScale(np.column_stack([np.ones(n2), rng.standard_normal((n2, p - 1))]))["std.covs"]
with DistributedDenseLikelihood(locs2, X2, rng.standard_normal(n2)) as dd:
    t2 = dd.terms(_lib.ML, tl, lim, tl["mean"])
print("SANITIZE_OK", S.shape, C.shape, Sc.shape, float(t["logdet"]), float(t2["logdet"]), float(sto[0]), float(dc[0, 0]))
