import sys
import numpy as np
sys.path.insert(0, ".")
import cocons_b200 as cb
from oracle import cov
rng = np.random.default_rng(424242)
for trial in range(40):
    n = int(rng.integers(20, 160)); p = int(rng.integers(1, 6))
    locs = rng.uniform(-1, 1, (n, 2)) * rng.choice([0.05, 1.0, 20.0])
    X = np.column_stack([np.ones(n), rng.standard_normal((n, p - 1))])
    th = {k: rng.uniform(-0.6, 0.6, p) for k in cov.ASPECTS}
    th["scale"][0] = rng.uniform(-5.0, 1.5)
    th["tilt"] = rng.uniform(-2.5, 2.5, p)
    th["nugget"][0] = rng.choice([-np.inf, -6.0, -2.0, 0.5])
    if np.isneginf(th["nugget"][0]): th["nugget"][1:] = 0.0
    lo = rng.uniform(0.1, 1.5)
    lim = [lo, lo + rng.choice([0.0, 0.3, 1.0, 3.5])]
    if rng.random() < 0.3:
        th["smooth"] = np.zeros(p); lim = [rng.choice([0.5, 1.5, 2.5, 0.8]), 0.0]; lim[1] = lim[0]
    if rng.random() < 0.3 and n > 4:
        locs[n - 1] = locs[1]; locs[n // 2] = locs[0]
    m = int(rng.integers(1, 40))
    lp = rng.uniform(-1, 1, (m, 2)) * np.abs(locs).max(); lp[0] = locs[0]
    Xp = np.column_stack([np.ones(m), rng.standard_normal((m, p - 1))])
    thc = dict(th, smooth=rng.uniform(-0.5, 0.8, p))
    ref = cov.cov_rns(th, locs, X, lim); got = cb.cov_rns(th, locs, X, lim)
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.where(ref != 0, np.abs(got - ref) / np.abs(ref), 0)
    if rel.max() > 3e-12:
        i, j = np.unravel_index(np.argmax(rel), rel.shape)
        t = np.pi / (1 + np.exp(-(X @ th["tilt"])))
        r = np.exp(2 * (X[:, 1:] @ th["scale"][1:])) if p > 1 else np.ones(n)
        a = np.exp(X @ th["aniso"])
        print("trial", trial, "n", n, "p", p, "lim", lim, "scale0", th["scale"][0], "dom", np.abs(locs).max())
        print(" worst", i, j, "ref %.6e rel %.2e -logC %.1f" % (ref[i, j], rel[i, j], -np.log(abs(ref[i, j]))))
        print(" t_i %.6f t_j %.6f sin %.3e %.3e  r %.3e %.3e a %.3e %.3e" % (t[i], t[j], np.sin(t[i]), np.sin(t[j]), r[i], r[j], a[i], a[j]))
        s11 = (r[i] + r[j]) / 2; s22 = (r[i] * a[i] ** 2 + r[j] * a[j] ** 2) / 2; s12 = (r[i] * a[i] * np.cos(t[i]) + r[j] * a[j] * np.cos(t[j])) / 2
        print(" s11 %.4e s22 %.4e s12 %.4e det %.4e det/(s11 s22) %.3e" % (s11, s22, s12, s11 * s22 - s12 ** 2, (s11 * s22 - s12 ** 2) / (s11 * s22)))
        dx, dy = locs[i] - locs[j]
        quad = s22 * dx * dx + s11 * dy * dy - 2 * s12 * dx * dy
        print(" dx %.3e dy %.3e quad %.4e terms %.4e %.4e %.4e" % (dx, dy, quad, s22 * dx * dx, s11 * dy * dy, 2 * s12 * dx * dy))
        break
