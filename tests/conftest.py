import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _philox_rounds_from_library():
    """The oracle's restatement of the kernels' generator uses the round count the library was BUILT with."""
    import ctypes
    from oracle import philox
    path = os.path.join(ROOT, "dlpm_b200", "libdlpm_b200.so")
    if os.path.exists(path):
        try:
            lib = ctypes.CDLL(path)
            lib.dlpm_b200_philox_rounds.restype = ctypes.c_int
            philox.ROUNDS = int(lib.dlpm_b200_philox_rounds())
        except (OSError, AttributeError):
            pass
    yield


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session")
def golden():
    return load_golden


def sub(g, prefix):
    """dict of the arrays under 'prefix/' in an npz."""
    return {k[len(prefix) + 1:]: g[k] for k in g.files if k.startswith(prefix + "/")}
