"""Worker of tests/test_multiproc.py: world-size-2 gloo run of the host-side multi-GPU logic, with
  * a numpy stand-in for the per-rank kernels (default), or
  * COCONS_MP_EMULATED=<scratch dir>: the shipped csrc/dist.cu itself - its cocons_dist_* C ABI and every kernel behind
    it - compiled for the host by tests/host_emul and executed on the CPU, the panels travelling over gloo
(on a GPU the same driver runs over NCCL: tests/test_gpu_dist.py, bench.py --gpus N)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cocons_b200 import _lib  # noqa: E402
from cocons_b200.distributed import PANEL, DistributedDenseLikelihood, _NoSync, fan_out  # noqa: E402
from oracle import cov, rmirror  # noqa: E402


class NumpyPanelOps(_NoSync):
    """Same interface as CudaPanelOps, arithmetic in numpy/LAPACK on the local column panels."""

    def __init__(self, locs, X, z, rank, world):
        self.locs, self.X = np.asarray(locs, float), np.asarray(X, float)
        self.n, self.p = self.X.shape
        self.z = np.asarray(z, float).reshape(self.n, -1)
        self.r = self.z.shape[1]
        self.rank, self.world = rank, world
        self.n_pad = (self.n + 127) // 128 * 128
        self.npanels = (self.n_pad + PANEL - 1) // PANEL
        self.cols = {}
        self.info = 0
        self.xb = None

    def close(self):
        pass

    def owner(self, K):
        pos = K % self.world
        return self.world - 1 - pos if (K // self.world) & 1 else pos

    def width(self, K):
        return min(PANEL, self.n_pad - K * PANEL)

    def buffer(self, count):
        return torch.zeros(max(int(count), 1), dtype=torch.float64)

    def panel_elems(self, K):
        return max(self.n_pad - (K + 1) * PANEL, 0) * self.width(K)

    def set_xbetas(self, xb):
        self.xb = np.asarray(xb, float).reshape(self.n, -1)

    def assemble(self, theta6, limits, mean):
        th = {k: np.array(theta6[i]) for i, k in enumerate(cov.ASPECTS)}
        S = np.eye(self.n_pad)
        S[: self.n, : self.n] = cov.cov_rns(th, self.locs, self.X, limits)
        self.mean = None if mean is None else np.asarray(mean, float)
        self.cols = {K: S[:, K * PANEL:K * PANEL + self.width(K)].copy()
                     for K in range(self.npanels) if self.owner(K) == self.rank}

    def factor_panel(self, K):
        assert self.owner(K) == self.rank
        c, r0, w = self.cols[K], K * PANEL, self.width(K)
        try:
            L = np.linalg.cholesky(c[r0:r0 + w, :])
        except np.linalg.LinAlgError:
            self.info = max(self.info, r0 + 1)
            L = np.eye(w)
        c[r0:r0 + w, :] = L
        c[r0 + w:, :] = np.linalg.solve(L, c[r0 + w:, :].T).T

    def pack_panel(self, K, buf):
        rows = self.n_pad - (K + 1) * PANEL
        if rows > 0:
            buf.numpy()[: rows * self.width(K)] = self.cols[K][(K + 1) * PANEL:, :].ravel(order="F")

    def update(self, K, buf, lo, hi):
        rows = self.n_pad - (K + 1) * PANEL
        if rows <= 0:
            return
        P = buf.numpy()[: rows * self.width(K)].reshape(rows, self.width(K), order="F")
        for J in range(max(lo, K + 1), min(hi, self.npanels)):
            if self.owner(J) != self.rank:
                continue
            off = J * PANEL - (K + 1) * PANEL
            self.cols[J][J * PANEL:, :] -= P[off:, :] @ P[off:off + self.width(J), :].T

    def fill_rhs(self, kind, rhs):
        zc = self.z - (self.X @ self.mean)[:, None] if (kind == _lib.ML and self.mean is not None) else self.z
        blocks = [zc]
        if kind == _lib.PROFILE:
            blocks = [self.xb, self.z]
        elif kind == _lib.REML:
            blocks = [self.X, self.z]
        B = np.zeros((self.npanels * PANEL, sum(b.shape[1] for b in blocks)))
        B[: self.n] = np.column_stack(blocks)
        nr = B.shape[1]
        rhs.numpy()[:] = B.reshape(self.npanels, PANEL, nr).transpose(0, 2, 1).ravel()
        return nr

    def solve_block(self, K, bK, tK, acc, Y, nr):
        w, r0 = self.width(K), K * PANEL
        b = (bK.numpy() - tK.numpy()).reshape(nr, PANEL).T[:w]
        y = np.linalg.solve(np.tril(self.cols[K][r0:r0 + w, :]), b)
        Yv = Y.numpy().reshape(nr, self.n_pad).T
        Yv[r0:r0 + w] = y
        below = self.cols[K][r0 + w:, :] @ y
        A = acc.numpy().reshape(self.npanels, nr, PANEL)
        for t in range(below.shape[0]):
            g = r0 + w + t
            A[g // PANEL, :, g % PANEL] += below[t]

    def reduce_local(self, Y, nr, out2, gram):
        s = 0.0
        for K, c in self.cols.items():
            d = np.diag(c[K * PANEL:K * PANEL + self.width(K), :])
            g = K * PANEL + np.arange(self.width(K))
            s += np.sum(np.log(d[g < self.n]))
        Yv = Y.numpy().reshape(nr, self.n_pad).T[: self.n]
        out2.numpy()[:] = [s, self.info]
        gram.numpy()[:] = (Yv.T @ Yv).ravel()


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    rng = np.random.default_rng(11)
    n, p = 1100, 3  # 1152 padded rows = 3 panels (512, 512, 128): ragged last panel, uneven ownership
    locs = rng.uniform(-1, 1, (n, 2))
    X = rmirror.get_scale(np.column_stack([np.ones(n), rng.standard_normal((n, p - 1))]))["std.covs"]
    z = rng.standard_normal((n, 2))
    tl = {"mean": np.array([0.1, 0.3, -0.2]), "std.dev": np.array([0.2, 0.15, 0.1]),
          "scale": np.array([-1.6, 0.2, -0.15]), "aniso": np.array([0.1, 0.2, -0.1]),
          "tilt": np.array([0.3, -0.2, 0.1]), "smooth": np.array([0.2, 0.3, -0.2]), "nugget": np.array([-4, 0.1, 0.1])}
    lim = [0.5, 2.5]
    res = {}
    if os.environ.get("COCONS_MP_EMULATED"):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from host_emul import build as emul_build
        from host_emul.panel_ops import EmulatedPanelOps
        work = os.path.join(os.environ["COCONS_MP_EMULATED"], "rank%d" % rank)
        os.makedirs(work, exist_ok=True)
        lib = emul_build.build(work)[0]
        with DistributedDenseLikelihood(locs, X, z, ops=EmulatedPanelOps(lib, locs, X, z, rank, world)) as d:
            t = d.terms(_lib.ML, tl, lim, tl["mean"])
            res["ml"] = [t["logdet"]] + list(t["quad"])
            t = d.terms(_lib.REML, tl, lim)  # p design columns + r columns of z as right-hand sides, Gram algebra, rank
            res["reml"] = [t["logdet"], t["logdet_w"], float(t["rank"])] + list(t["quad"])
            res["panels"] = [d.npanels, [d.owner(K) for K in range(d.npanels)]]
        if rank == 0:
            S = cov.cov_rns(tl, locs, X, lim)
            R = rmirror.r_chol(S)
            y = rmirror._fwd(R, z - (X @ tl["mean"])[:, None])
            res["ml_ref"] = [float(np.sum(np.log(np.diag(R))))] + list((y * y).sum(axis=0))
            Yx, yz = rmirror._fwd(R, X), rmirror._fwd(R, z)
            W, b = Yx.T @ Yx, Yx.T @ yz
            quad = (yz * yz).sum(axis=0) - np.einsum("ij,ij->j", b, np.linalg.solve(W, b))
            res["reml_ref"] = [float(np.sum(np.log(np.diag(R)))), float(np.sum(np.log(np.diag(np.linalg.cholesky(W))))),
                               float(rmirror.r_qr_rank(X))] + list(quad)
            print("RESULT " + json.dumps(res))
        dist.destroy_process_group()
        return
    with DistributedDenseLikelihood(locs, X, z, ops=NumpyPanelOps(locs, X, z, rank, world)) as d:
        t = d.terms(_lib.ML, tl, lim, tl["mean"])
        res["ml"] = [t["logdet"]] + list(t["quad"])
        d.set_xbetas(X[:, :2])
        t = d.terms(_lib.PROFILE, tl, lim)
        res["profile"] = [t["logdet"], t["logdet_w"]] + list(t["quad"])
        try:
            # fixed non-half-integer smoothness makes Sigma singular
            d.terms(_lib.ML, dict(tl, smooth=np.zeros(3)), [1.0, 1.0])
            res["notpd"] = False
        except Exception as e:  # noqa: BLE001
            res["notpd"] = type(e).__name__
    # regime 1: fan-out of independent evaluations
    pts = [0.1 * k for k in range(7)]
    res["fan"] = fan_out(pts, lambda x: x * x + rank * 0.0)
    # the optimiser's batched finite-difference gradient over the same fan-out
    from cocons_b200.api import fd_value_and_grad
    quad = lambda th: float(np.sum((th - np.array([0.5, -1.0, 2.0])) ** 2) + 3.0)  # noqa: E731
    f0, g = fd_value_and_grad(quad, np.array([0.0, 0.0, 3.0]), np.array([-5.0, -5.0, -5.0]), np.array([5.0, 5.0, 3.0]),
                              1e-4)
    res["fd"] = [f0] + list(g)
    if rank == 0:
        S = cov.cov_rns(tl, locs, X, lim)
        R = rmirror.r_chol(S)
        y = rmirror._fwd(R, z - (X @ tl["mean"])[:, None])
        res["ml_ref"] = [float(np.sum(np.log(np.diag(R))))] + list((y * y).sum(axis=0))
        Yx, yz = rmirror._fwd(R, X[:, :2]), rmirror._fwd(R, z)
        W, b = Yx.T @ Yx, Yx.T @ yz
        quad = (yz * yz).sum(axis=0) - np.einsum("ij,ij->j", b, np.linalg.solve(W, b))
        res["profile_ref"] = [float(np.sum(np.log(np.diag(R)))), float(np.sum(np.log(np.diag(np.linalg.cholesky(W)))))] \
            + list(quad)
        print("RESULT " + json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
