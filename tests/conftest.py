import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def lib_built():
    """libdeplex_b200.so must exist (built by __graft_entry__.build()); build it here if a compiler is around."""
    from deplex_b200 import _capi
    if not os.path.exists(_capi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _capi.LIB_PATH


def load_frame(name):
    """(depth uint16 (480,640), K dict, ini path) of a shipped reference frame (tests/golden fixtures)."""
    cfgname = {"tum": "TUM_fr3_long_val", "icl": "ICL_living_room"}[name]
    depth = np.load(os.path.join(GOLDEN, f"{name}_depth.npz"))["depth"]
    K = np.loadtxt(os.path.join(GOLDEN, cfgname + ".K"), dtype=np.float32)
    k = dict(fx=float(K[0, 0]), fy=float(K[1, 1]), cx=float(K[0, 2]), cy=float(K[1, 2]))
    return depth, k, os.path.join(GOLDEN, cfgname + ".ini")


def frame_cloud(name, layout="rowmajor"):
    from deplex_b200 import synth
    depth, k, ini = load_frame(name)
    return synth.depth_to_cloud(depth, k, layout), ini


def to_oracle_cfg(oracle, cfg):
    """deplex_b200.Config -> OracleConfig (same field names)."""
    return oracle.OracleConfig(**cfg.as_dict())
