import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "schwarzchild-einsteinpy-raytracing_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


def _cuda_device_count():
    """Devices the CUDA driver sees, without importing torch (0 when there is no driver)."""
    import ctypes
    try:
        cu = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int(0)
        if cu.cuInit(0) != 0 or cu.cuDeviceGetCount(ctypes.byref(n)) != 0:
            return 0
        return n.value
    except OSError:
        return 0


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a box without a GPU skips the gpu-marked tests instead of failing in them.  Selecting them
    explicitly (`-m gpu`, as the driver does on the GPU box) or SCHW_REQUIRE_GPU=1 keeps the hard failure: there is no
    CPU fallback to fall back to, and a GPU box without a usable device must not look green."""
    if config.getoption("-m") or os.environ.get("SCHW_REQUIRE_GPU") == "1" or _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device (run with -m gpu on the GPU box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return load


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure)."""
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def handle():
    """GPU handle through the C ABI.  No skip: on the GPU box a missing library or device is a failure."""
    import schw_b200
    return schw_b200.Handle()


def relerr(a, b, floor=1e-300):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)
