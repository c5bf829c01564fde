"""Host driver of the multi-GPU dense likelihood (one process per GPU, torch.distributed).

Two regimes:

* fan_out()  - the matrix fits one GPU: the independent objective evaluations an optimiser step
  needs (the 2p+1 finite-difference points of R/optim.R:237-259, the Hessian points of
  R/getFunctions.R:979-1016) are dealt to the ranks; no data-path collective, only the scalar
  results are gathered.

* DistributedDenseLikelihood - the matrix does NOT fit one GPU (n = 200 000 is 320 GB): column
  panels of 512 are dealt round-robin to the ranks, every rank assembles and updates only its
  own panels (cocons_dist_* in libcocons_b200.so), and the exchange step is ONE broadcast of
  the packed, factored panel per outer step over NCCL, issued here, with one-panel look-ahead:
  the owner of panel K+1 applies update K to that panel first, factors it and starts its
  broadcast while every rank is still applying update K to the rest of its panels.

torch is plumbing here (process group, device buffers for the exchanged panels, stream order);
every flop is in the library's kernels.  `ops` abstracts the per-rank kernels so that the
ownership / look-ahead / reduction logic can be exercised on CPU with a numpy stand-in
(tests/test_multiproc.py, gloo, world size 2) - the product path always uses CudaPanelOps.
"""
import numpy as np

from . import _lib
from ._lib import NotPositiveDefinite

PANEL = 512


# ------------------------------------------------------------------------------------------
# regime 1: independent evaluations
# ------------------------------------------------------------------------------------------
def fan_out(points, fn, group=None):
    """Evaluate fn(point) for every point, rank r taking points[r::world]; every rank returns the
    full list of values in input order."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [fn(pt) for pt in points]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = [float(fn(pt)) for pt in points[rank::world]]
    per = (len(points) + world - 1) // world
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    buf = torch.full((per,), float("nan"), dtype=torch.float64, device=dev)
    buf[: len(mine)] = torch.tensor(mine, dtype=torch.float64)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    vals = [None] * len(points)
    for r in range(world):
        got = out[r].cpu().numpy()
        for k, idx in enumerate(range(r, len(points), world)):
            vals[idx] = float(got[k])
    return vals


# ------------------------------------------------------------------------------------------
# regime 2: one matrix over all GPUs
# ------------------------------------------------------------------------------------------
class CudaPanelOps:
    """Per-rank kernels: thin ctypes calls into cocons_dist_* (cocons_b200/csrc/dist.cu)."""

    def __init__(self, locs, X, z, rank, world, device, stream):
        import torch
        self.torch = torch
        self.locs, self.X = _lib.fmat(locs), _lib.fmat(X)
        self.n, self.p = self.X.shape
        self.z = _lib.fmat(z, rows=self.n)
        self.r = self.z.shape[1]
        self.h = _lib._vp()
        self.device = torch.device("cuda", device)
        _lib.check(_lib.lib().cocons_dist_create(int(device), int(rank), int(world), self.n, self.p, self.r,
                                                 _lib.ptr(self.locs), _lib.ptr(self.X), _lib.ptr(self.z), stream,
                                                 self.h))
        self.npanels = int(_lib.lib().cocons_dist_npanels(self.h))
        self.n_pad = int(_lib.lib().cocons_dist_npad(self.h))
        self.main = torch.cuda.current_stream()
        self.side = torch.cuda.ExternalStream(int(_lib.lib().cocons_dist_side_stream(self.h)), device=self.device)

    def close(self):
        if self.h is not None and self.h.value:
            _lib.lib().cocons_dist_destroy(self.h)
            self.h = _lib._vp()

    def buffer(self, count):
        return self.torch.zeros(max(int(count), 1), dtype=self.torch.float64, device=self.device)

    def panel_elems(self, K):
        return int(_lib.lib().cocons_dist_panel_elems(self.h, K))

    def set_xbetas(self, xb):
        xb = _lib.fmat(xb, rows=self.n)
        _lib.check(_lib.lib().cocons_dist_set_xbetas(self.h, xb.shape[1], _lib.ptr(xb)))

    def assemble(self, theta6, limits, mean):
        _lib.check(_lib.lib().cocons_dist_assemble(self.h, _lib.ptr(theta6), _lib.ptr(limits), _lib.ptr(mean)))

    def factor_panel(self, K, side=False):
        _lib.check(_lib.lib().cocons_dist_factor_panel(self.h, K, int(side)))

    def pack_panel(self, K, buf, side=False):
        _lib.check(_lib.lib().cocons_dist_pack_panel(self.h, K, buf.data_ptr(), int(side)))

    # stream plumbing for the look-ahead: the side stream waits for what the main stream has queued so far
    def side_after_main(self):
        ev = self.torch.cuda.Event()
        ev.record(self.main)
        self.side.wait_event(ev)

    def on_side(self):
        return self.torch.cuda.stream(self.side)

    def update(self, K, buf, lo, hi):
        _lib.check(_lib.lib().cocons_dist_update(self.h, K, buf.data_ptr(), lo, hi))

    def fill_rhs(self, kind, rhs):
        nr = _lib.ctypes.c_int()
        _lib.check(_lib.lib().cocons_dist_fill_rhs(self.h, int(kind), rhs.data_ptr(), _lib.ctypes.byref(nr)))
        return nr.value

    def solve_block(self, K, bK, tK, acc, Y, nr):
        _lib.check(_lib.lib().cocons_dist_solve_block(self.h, K, bK.data_ptr(), tK.data_ptr(), acc.data_ptr(),
                                                      Y.data_ptr(), nr))

    def reduce_local(self, Y, nr, out2, gram):
        _lib.check(_lib.lib().cocons_dist_reduce_local(self.h, Y.data_ptr(), nr, out2.data_ptr(), gram.data_ptr()))


class DistributedDenseLikelihood:
    """-2 loglik terms of ONE matrix spread over all ranks of `group` (same outputs as
    DenseLikelihood.terms).  Every rank passes the same locs / X / z."""

    def __init__(self, locs, x_covariates, z, group=None, device=None, ops=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if ops is None:
            device = torch.cuda.current_device() if device is None else device
            ops = CudaPanelOps(locs, x_covariates, z, self.rank, self.world, device,
                               torch.cuda.current_stream().cuda_stream)
        self.ops = ops
        self.n, self.p, self.r = ops.n, ops.p, ops.r
        self.npanels = ops.npanels
        self.q = 0
        biggest = max([ops.panel_elems(K) for K in range(self.npanels)] + [1])
        self.bufs = [ops.buffer(biggest), ops.buffer(biggest)]
        self.last_phase_s = {}
        import os
        self.profile = bool(os.environ.get("COCONS_DIST_PROFILE"))  # per-step broadcast-wait times (debugging)
        self.last_wait_ms = None

    def close(self):
        self.ops.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def owner(self, K):
        """Snake deal (0..N-1, N-1..0, ...), matching csrc/dist.cu: evens out the per-step update work."""
        pos = K % self.world
        return self.world - 1 - pos if (K // self.world) & 1 else pos

    def set_xbetas(self, xb):
        self.ops.set_xbetas(xb)
        self.q = np.asarray(xb).reshape(self.n, -1).shape[1]

    def _bcast(self, K, async_op):
        count = self.ops.panel_elems(K)
        if count == 0 or self.world == 1:
            return None
        return self.dist.broadcast(self.bufs[K % 2][:count], src=self._global(self.owner(K)), group=self.group,
                                   async_op=async_op)

    class _SideDone:
        """world == 1: stands in for the broadcast handle - the main stream waits for the side stream"""

        def __init__(self, ops):
            self.ops = ops
            self.ev = ops.torch.cuda.Event()
            self.ev.record(ops.side)

        def wait(self):
            self.ops.main.wait_event(self.ev)

    def _global(self, rank_in_group):
        if self.group is None:
            return rank_in_group
        return self.dist.get_global_rank(self.group, rank_in_group)

    # -- factorisation with one-panel look-ahead ------------------------------------------
    def factor(self, theta6, limits, mean=None):
        """Right-looking factorisation with one-panel look-ahead.  Step K on every rank:
             main stream : wait for panel K | (owner of K+1: apply update K to panel K+1) | apply update K
                           to the rest of the rank's panels
             side stream : (owner of K+1: factor panel K+1, pack it) | broadcast of panel K+1
           The side stream has high priority, so the small panel kernels and the NCCL broadcast overlap
           the main stream's DMMA updates; buffers alternate (K mod 2)."""
        ops, me = self.ops, self.rank
        two_streams = hasattr(ops, "on_side")
        ops.assemble(theta6, limits, mean)
        if self.owner(0) == me:
            ops.factor_panel(0)
            ops.pack_panel(0, self.bufs[0])
        work = self._bcast(0, async_op=True)
        prof = [] if (two_streams and self.profile) else None
        for K in range(self.npanels):
            if prof is not None:
                e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
                e0.record(ops.main)
            if work is not None:
                work.wait()  # the main stream now waits for panel K
            if prof is not None:
                e1.record(ops.main)
                prof.append((e0, e1))
            nxt = K + 1
            work = None
            if nxt >= self.npanels:
                break
            if not two_streams:  # CPU stand-in ops of the tests: same order, one queue
                if self.owner(nxt) == me:
                    ops.update(K, self.bufs[K % 2], nxt, nxt + 1)
                    ops.factor_panel(nxt)
                    ops.pack_panel(nxt, self.bufs[nxt % 2])
                work = self._bcast(nxt, async_op=True)
                ops.update(K, self.bufs[K % 2], nxt + 1, self.npanels)
                continue
            if self.owner(nxt) == me:
                ops.update(K, self.bufs[K % 2], nxt, nxt + 1)
            # everything queued on main so far (all of update K-1, the look-ahead piece of update K) precedes
            # the side-stream work: panel K+1 is current, and bufs[(K+1) % 2] is no longer being read
            ops.side_after_main()
            if self.owner(nxt) == me:
                ops.factor_panel(nxt, side=True)
                ops.pack_panel(nxt, self.bufs[nxt % 2], side=True)
            with ops.on_side():
                work = self._bcast(nxt, async_op=True)
                if work is None:
                    work = self._SideDone(ops)
            ops.update(K, self.bufs[K % 2], nxt + 1, self.npanels)
        if prof is not None:
            self.torch.cuda.synchronize()
            waits = [a.elapsed_time(b) for a, b in prof]
            self.last_wait_ms = {"total": float(sum(waits)), "first_100": float(sum(waits[:100])),
                                 "last_100": float(sum(waits[-100:])), "max": float(max(waits))}

    # -- solves + reductions --------------------------------------------------------------
    def terms(self, kind, theta_list, smooth_limits, mean=None):
        import time
        torch, dist = self.torch, self.dist
        th = _lib.pack_theta(theta_list, self.p)
        lim = np.ascontiguousarray(np.asarray(smooth_limits, dtype=np.float64))
        mean_v = None if mean is None else np.ascontiguousarray(np.asarray(mean, dtype=np.float64))
        t0 = time.perf_counter()
        self.factor(th, lim, mean_v if kind == _lib.ML else None)
        self._sync()
        t1 = time.perf_counter()
        ops = self.ops
        qx = self.q if kind == _lib.PROFILE else (self.p if kind == _lib.REML else 0)
        nr = qx + self.r
        rhs = ops.buffer(self.npanels * nr * PANEL)
        acc = ops.buffer(self.npanels * nr * PANEL)
        Y = ops.buffer(ops.n_pad * nr)
        got = ops.fill_rhs(kind, rhs)
        assert got == nr
        blk = nr * PANEL
        for K in range(self.npanels):
            tK = acc[K * blk:(K + 1) * blk]
            if self.world > 1:
                dist.reduce(tK, dst=self._global(self.owner(K)), op=dist.ReduceOp.SUM, group=self.group)
            if self.owner(K) == self.rank:
                ops.solve_block(K, rhs[K * blk:(K + 1) * blk], tK, acc, Y, nr)
        out2, gram = ops.buffer(2), ops.buffer(nr * nr)
        ops.reduce_local(Y, nr, out2, gram)
        if self.world > 1:
            flag = out2[1:2].clone()
            dist.all_reduce(out2[0:1], op=dist.ReduceOp.SUM, group=self.group)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
            dist.all_reduce(gram, op=dist.ReduceOp.SUM, group=self.group)
            out2[1:2] = flag
        self._sync()
        t2 = time.perf_counter()
        self.last_phase_s = {"assemble_factor_s": t1 - t0, "solve_s": t2 - t1}
        o2 = out2.cpu().numpy()
        G = gram.cpu().numpy().reshape(nr, nr)
        if o2[1] > 0:
            raise NotPositiveDefinite(int(o2[1]))
        return _terms_from_gram(kind, float(o2[0]), G, qx, self.r, self.ops.X)

    def _sync(self):
        if self.torch.cuda.is_available() and not isinstance(self.ops, _NoSync):
            self.torch.cuda.synchronize()


class _NoSync:
    """marker base for CPU stand-in ops used by the tests"""


def _terms_from_gram(kind, logdet, G, qx, r, X):
    """Same small algebra as cocons_n2ll's host epilogue (capi.cu): quadratic forms from the Gram
    matrix of the solved right-hand sides."""
    out = {"logdet": logdet, "logdet_w": 0.0, "rank": 0}
    if qx == 0:
        out["quad"] = np.diag(G)[:r].copy()
        return out
    W = G[:qx, :qx]
    Lw = np.linalg.cholesky(W)
    out["logdet_w"] = float(np.sum(np.log(np.diag(Lw))))
    quad = np.empty(r)
    for j in range(r):
        t = np.linalg.solve(Lw, G[:qx, qx + j])
        quad[j] = G[qx + j, qx + j] - t @ t
    out["quad"] = quad
    if kind == _lib.REML:  # qr(X)$rank, R/neg2loglikelihood.R:270 - the same routine the single-GPU path uses
        Xf = _lib.fmat(X)
        out["rank"] = int(_lib.lib().cocons_qr_rank(_lib.ptr(Xf), Xf.shape[0], Xf.shape[1], 1e-7))
    return out
