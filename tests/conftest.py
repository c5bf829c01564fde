import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def datasets():
    return dict(np.load(os.path.join(GOLD, "datasets.npz"), allow_pickle=False))


@pytest.fixture(scope="session")
def cov_cases():
    raw = np.load(os.path.join(GOLD, "cov_cases.npz"), allow_pickle=False)
    cases = {}
    for key in raw.files:
        name, field = key.split("__")
        cases.setdefault(name, {})[field] = raw[key]
    return cases


@pytest.fixture(scope="session")
def taper_cases():
    raw = np.load(os.path.join(GOLD, "taper_cases.npz"), allow_pickle=False)
    cases = {}
    for key in raw.files:
        name, field = key.split("__")
        cases.setdefault(name, {})[field] = raw[key]
    return cases


@pytest.fixture(scope="session")
def n2ll_cases():
    with open(os.path.join(GOLD, "n2ll_cases.json")) as f:
        doc = json.load(f)
    out = {}
    for c in doc["cases"]:
        c = dict(c)
        c["par_pos"] = {k: (np.array(v, dtype=bool) if isinstance(v, list) else v) for k, v in c["par_pos"].items()}
        c["theta"] = np.array(c["theta"])
        out[c["name"]] = c
    return out


@pytest.fixture(scope="session")
def emu(tmp_path_factory):
    """libcocons_b200.so's sources compiled for the HOST against the CUDA execution-model stand-in of tests/host_emul
    (test scaffolding; one build per session)."""
    from host_emul import build as emul_build
    lib, barriers, launches = emul_build.build(tmp_path_factory.mktemp("host_emul"))
    lib._barriers, lib._rewritten = barriers, launches
    return lib


@pytest.fixture
def product_on_host(emu, monkeypatch):
    """The product's Python layer (cocons_b200.api / _lib) bound, for the duration of ONE test, to the host build of
    the library's own sources instead of libcocons_b200.so - so that the host orchestration of capi.cu and the kernels
    behind it can be checked without a GPU.  Never reachable from the product: _lib.lib() itself only ever loads
    libcocons_b200.so, which has no CPU path."""
    from cocons_b200 import _lib
    for name, (res, args) in _lib.SIGNATURES.items():
        fn = getattr(emu, name)
        fn.restype, fn.argtypes = res, args
    monkeypatch.setattr(_lib, "_lib", emu)
    assert _lib.lib() is emu and emu.cocons_device_count() == 1
    yield emu
    emu.cocons_release_workspace()


def theta_dict(theta6):
    from oracle.cov import ASPECTS
    return {k: np.array(theta6[i]) for i, k in enumerate(ASPECTS)}


def relerr(a, b):
    """max |a-b| / |b| over entries, with exact agreement required where b == 0."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    zero = b == 0
    if np.any(zero) and not np.array_equal(a[zero], b[zero]):
        return np.inf
    if np.all(zero):
        return 0.0
    return float(np.max(np.abs(a[~zero] - b[~zero]) / np.abs(b[~zero])))


def case_design(case, datasets):
    """(locs, X_std, z, x_betas=X) of an n2ll golden case, rebuilt from the dataset fixture with the
    PRODUCT's getScale (so the host mirror is exercised too)."""
    from cocons_b200 import getScale
    n = case["n"]
    if case["dataset"] == "holes":
        M = datasets["holes_training"]
        cols, z = [2, 3], M[:n, 4]
    elif case["dataset"] == "holes_bm":
        M = datasets["holes_bm_training"]
        cols, z = [2, 3], datasets["holes_bm_training_z"][:n]
    else:
        M = datasets["stripes_training"]
        cols, z = [2, 3, 4], M[:n, 5]
    X = getScale(np.column_stack([np.ones(n)] + [M[:n, c] for c in cols]))["std.covs"]
    return M[:n, :2], X, np.asarray(z, dtype=np.float64).reshape(n, -1)
