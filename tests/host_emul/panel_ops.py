"""TEST SCAFFOLDING: the per-rank operations of cocons_b200.distributed.DistributedDenseLikelihood backed by the
HOST-EMULATED build of cocons_b200/csrc/dist.cu (tests/host_emul/build.py): the same cocons_dist_* C ABI, the same
kernels (cyclic-slab assembly, blocked Cholesky of the owned panels, row packing, DMMA trailing updates, blocked
solve, local reductions), executed on the CPU; the exchanged buffers are CPU torch tensors, so the driver's
broadcasts / reduces run over gloo.  Mirrors CudaPanelOps call for call (one queue: no side stream)."""
import ctypes

import torch

from cocons_b200 import _lib
from cocons_b200.distributed import _NoSync


class EmulatedPanelOps(_NoSync):
    def __init__(self, lib, locs, X, z, rank, world):
        self.lib = lib
        for name, (res, args) in _lib.SIGNATURES.items():
            if name.startswith("cocons_dist_"):
                fn = getattr(lib, name)
                fn.restype, fn.argtypes = res, args
        self.locs, self.X = _lib.fmat(locs), _lib.fmat(X)
        self.n, self.p = self.X.shape
        self.z = _lib.fmat(z, rows=self.n)
        self.r = self.z.shape[1]
        self.h = _lib._vp()
        self._check(lib.cocons_dist_create(0, int(rank), int(world), self.n, self.p, self.r, _lib.ptr(self.locs),
                                           _lib.ptr(self.X), _lib.ptr(self.z), None, self.h))
        self.npanels = int(lib.cocons_dist_npanels(self.h))
        self.n_pad = int(lib.cocons_dist_npad(self.h))

    def _check(self, rc):
        if rc < 0:
            raise RuntimeError("emulated cocons_dist call failed (%d): %s" % (rc, self.lib.emu_last_error().decode()))

    def close(self):
        if self.h is not None and self.h.value:
            self.lib.cocons_dist_destroy(self.h)
            self.h = _lib._vp()

    def buffer(self, count):
        return torch.zeros(max(int(count), 1), dtype=torch.float64)

    def panel_elems(self, K):
        return int(self.lib.cocons_dist_panel_elems(self.h, K))

    def set_xbetas(self, xb):
        xb = _lib.fmat(xb, rows=self.n)
        self._check(self.lib.cocons_dist_set_xbetas(self.h, xb.shape[1], _lib.ptr(xb)))

    def assemble(self, theta6, limits, mean):
        self._check(self.lib.cocons_dist_assemble(self.h, _lib.ptr(theta6), _lib.ptr(limits), _lib.ptr(mean)))

    def factor_panel(self, K):
        self._check(self.lib.cocons_dist_factor_panel(self.h, K, 0))

    def pack_panel(self, K, buf):
        self._check(self.lib.cocons_dist_pack_panel(self.h, K, buf.data_ptr(), 0))

    def update(self, K, buf, lo, hi):
        self._check(self.lib.cocons_dist_update(self.h, K, buf.data_ptr(), lo, hi))

    def fill_rhs(self, kind, rhs):
        nr = ctypes.c_int()
        self._check(self.lib.cocons_dist_fill_rhs(self.h, int(kind), rhs.data_ptr(), ctypes.byref(nr)))
        return nr.value

    def solve_block(self, K, bK, tK, acc, Y, nr):
        self._check(self.lib.cocons_dist_solve_block(self.h, K, bK.data_ptr(), tK.data_ptr(), acc.data_ptr(),
                                                     Y.data_ptr(), nr))

    def reduce_local(self, Y, nr, out2, gram):
        self._check(self.lib.cocons_dist_reduce_local(self.h, Y.data_ptr(), nr, out2.data_ptr(), gram.data_ptr()))
