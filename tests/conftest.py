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


def pytest_sessionstart(session):
    """The native library is git-ignored (built in tree): build it when missing or stale so both the
    CPU suite (ABI export checks) and the GPU suite run against the current sources."""
    import __graft_entry__ as ge
    try:
        if ge._stale():
            ge.build()
    except Exception as e:  # no nvcc on this box: the tests that need the library will say so
        print(f"[conftest] could not build librobchar_b200.so: {e}")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def golden():
    return load_golden
