"""GPU parity of the block-cyclic multi-GPU path's per-rank kernels (cocons_dist_*).  With one
process the driver owns every panel: the panel-by-panel factorisation, packed-panel updates and
the blocked solve must reproduce the single-GPU objective terms.  (World size > 1 is exercised by
tools/dist_check.py under torchrun and, for the host logic, by tests/test_multiproc.py on gloo.)"""
import numpy as np
import pytest

import cocons_b200 as cb
from cocons_b200 import _lib
from cocons_b200.distributed import DistributedDenseLikelihood
from conftest import case_design

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["holes777_ragged", "holes1500_general", "stripes2000_p4"])
def test_single_rank_distributed_path_matches_resident_context(name, n2ll_cases, datasets):
    c = n2ll_cases[name]
    locs, X, z = case_design(c, datasets)
    p = c["p"]
    tl = cb.getModelLists(c["theta"], c["par_pos"], "diff")
    with cb.DenseLikelihood(locs, X, z) as ctx:
        ctx.set_xbetas(X)
        ref = {k: ctx.terms(k, tl, c["limits"], tl["mean"]) for k in (_lib.ML, _lib.PROFILE, _lib.REML)}
    with DistributedDenseLikelihood(locs, X, z) as d:
        d.set_xbetas(X)
        for kind in (_lib.ML, _lib.PROFILE, _lib.REML):
            got = d.terms(kind, tl, c["limits"], tl["mean"])
            assert abs(got["logdet"] - ref[kind]["logdet"]) < 1e-11 * abs(ref[kind]["logdet"])
            # two orchestrations of the same kernels: different blocking of the solve, so rounding-level agreement
            assert np.allclose(got["quad"], ref[kind]["quad"], rtol=1e-10, atol=0)
            assert abs(got["logdet_w"] - ref[kind]["logdet_w"]) <= 1e-9 * max(1.0, abs(ref[kind]["logdet_w"]))
    # and the value against the committed golden
    n = c["n"]
    v = n * np.log(2 * np.pi) + 2 * ref[_lib.ML]["logdet"] + ref[_lib.ML]["quad"][0]
    assert abs(v - c["values"]["ml"]) < 1e-9 * abs(v)


def test_not_positive_definite_is_reported(n2ll_cases, datasets):
    c = n2ll_cases["holes300_notpd"]
    locs, X, z = case_design(c, datasets)
    tl = cb.getModelLists(c["theta"], c["par_pos"], "diff")
    with DistributedDenseLikelihood(locs, X, z) as d:
        with pytest.raises(cb.NotPositiveDefinite):
            d.terms(_lib.ML, tl, c["limits"], tl["mean"])


def test_two_factorisation_drivers_agree_at_scale():
    """Size-independent cross-check at a size no CPU oracle reaches in seconds (n = 20 000): the
    look-ahead driver of the resident context and the panel-by-panel driver of the distributed
    path are different orchestrations of the same kernels and must give the same terms."""
    import bench
    n = 20000
    locs, X, z = bench.synthetic(n)
    with cb.DenseLikelihood(locs, X, z) as ctx:
        a = ctx.terms(_lib.ML, bench.THETA, bench.LIMITS, bench.THETA["mean"])
    with DistributedDenseLikelihood(locs, X, z) as d:
        b = d.terms(_lib.ML, bench.THETA, bench.LIMITS, bench.THETA["mean"])
    assert abs(a["logdet"] - b["logdet"]) < 1e-11 * abs(a["logdet"])
    assert abs(a["quad"][0] - b["quad"][0]) < 1e-10 * abs(a["quad"][0])
    with DistributedDenseLikelihood(locs, X, z) as d:  # and the distributed driver is bit-reproducible
        for _ in range(3):
            c = d.terms(_lib.ML, bench.THETA, bench.LIMITS, bench.THETA["mean"])
            assert c["logdet"] == b["logdet"] and c["quad"][0] == b["quad"][0]
