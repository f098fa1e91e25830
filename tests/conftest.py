import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib_built():
    """Build libmpxb.so (and the host test harness) once if it is missing."""
    import __graft_entry__ as g
    from mpx_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        g.build()
    return _lib.LIB_PATH
