"""The engine headers must stay NVRTC-clean (no host C library headers, no unannotated host-only code on the device path):
the planned user-function hook compiles them at run time.  Needs no GPU (offline target sm_100a)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def test_engine_compiles_under_nvrtc():
    try:
        import nvrtc_probe
    except Exception as e:  # noqa: BLE001
        pytest.skip(f"cuda-python nvrtc bindings not available: {e}")
    rc, dt, log, nb = nvrtc_probe.compile_engine()
    assert rc == 0, log[:2000]
    assert nb > 100_000


def test_user_model_compiles_under_nvrtc():
    """DYN == 2: a nonlinear model given as device source (dynamics + measurement log-likelihood) inlined into the sweep."""
    try:
        import nvrtc_probe
    except Exception as e:  # noqa: BLE001
        pytest.skip(f"cuda-python nvrtc bindings not available: {e}")
    rc, dt, log, nb = nvrtc_probe.compile_engine(nvrtc_probe.USER_SRC)
    assert rc == 0, log[:2000]
    assert nb > 100_000

