import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _gpu_available():
    """True when the CUDA library is built and sees an sm_100 device (no engine is created)."""
    try:
        from pynngp_b200 import _lib

        return _lib.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # a plain `pytest tests` on a box without a B200 runs the CPU tier and reports the GPU tier as skipped
    # (`-m gpu` on the GPU box is unaffected: there the device exists)
    if _gpu_available():
        return
    skip = pytest.mark.skip(reason="no B200 visible (or libnngp_b200.so not built): GPU parity tests need the device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
