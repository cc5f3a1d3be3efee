import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Make sure the C-ABI library and the oracle exist (compiles with nvcc/gcc if stale)."""
    import __graft_entry__ as g
    g.build()
    return g.LIB


@pytest.fixture(scope="session")
def gpu(built):
    import llpf_b200 as L
    import ctypes as C
    lib = L.load_library()
    n = C.c_int()
    rc = lib.llpf_device_count(C.byref(n))
    if rc != 0 or n.value < 1:
        pytest.fail("GPU test selected but no CUDA device is visible (the product has no CPU fallback)")
    return L
