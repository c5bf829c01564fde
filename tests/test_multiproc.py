"""World-size-2 gloo run (CPU) of the multi-GPU host logic: panel ownership, look-ahead order,
the per-panel reduce of the distributed solve, the final all-reduces and the evaluation fan-out."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def _run_world_of_two(**extra_env):
    port = _free_port()
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE="2", OMP_NUM_THREADS="2",
               **extra_env)
    procs = []
    for rank in range(2):
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_mp_worker.py")],
                                      env=dict(env, RANK=str(rank), LOCAL_RANK=str(rank)), stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=900) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-3000:]
    line = [l for l in outs[0][0].splitlines() if l.startswith("RESULT ")][0]
    return json.loads(line[len("RESULT "):])


def test_world_size_two_with_the_shipped_dist_kernels_on_the_host(tmp_path):
    """The block-cyclic Cholesky of ONE matrix over two ranks with the product's own csrc/dist.cu - the cocons_dist_*
    C ABI and every kernel behind it (cyclic-slab assembly, DMMA panel factorisation, row packing, trailing updates,
    blocked solve, local reductions), compiled for the host and run on the CPU (tests/host_emul) - driven by the
    product's DistributedDenseLikelihood, the packed panels broadcast and the solve blocks reduced over gloo.
    n = 1100: three panels (512, 512, 128), dealt 0, 1, 1 by the snake."""
    res = _run_world_of_two(COCONS_MP_EMULATED=str(tmp_path))
    assert res["panels"] == [3, [0, 1, 1]]
    assert np.allclose(res["ml"], res["ml_ref"], rtol=1e-11, atol=0), (res["ml"], res["ml_ref"])
    assert np.allclose(res["reml"], res["reml_ref"], rtol=1e-10, atol=0), (res["reml"], res["reml_ref"])


def test_world_size_two_matches_single_process_reference():
    res = _run_world_of_two()
    assert np.allclose(res["ml"], res["ml_ref"], rtol=1e-10, atol=0)
    assert np.allclose(res["profile"], res["profile_ref"], rtol=1e-9, atol=0)
    assert res["notpd"] == "NotPositiveDefinite"
    assert np.allclose(res["fan"], [(0.1 * k) ** 2 for k in range(7)])
    # f = |th - c|^2 + 3 at th = (0, 0, 3), c = (.5, -1, 2); third coordinate sits on its upper bound
    assert abs(res["fd"][0] - (0.25 + 1 + 1 + 3)) < 1e-12
    assert np.allclose(res["fd"][1:3], [-1.0, 2.0], atol=1e-6)
    assert abs(res["fd"][3] - 2.0) < 2e-4  # one-sided at the bound
