import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_device_count():
    """number of CUDA devices the driver exposes (0 without a driver), asked of libcudart directly: no torch import at collection"""
    import ctypes
    for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
        try:
            rt = ctypes.CDLL(name)
        except OSError:
            continue
        n = ctypes.c_int(0)
        return n.value if rt.cudaGetDeviceCount(ctypes.byref(n)) == 0 else 0
    return 0


def pytest_collection_modifyitems(config, items):
    """a plain `pytest` on a machine without a GPU skips the -m gpu tests instead of failing inside the CUDA runtime"""
    if not any("gpu" in it.keywords for it in items) or _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device (these are the -m gpu parity tests; the product itself fails loudly without one)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def evp_lib():
    """the CUDA library, built in-tree; GPU tests call through its C ABI only."""
    from cice_b200 import build
    build.build()
    from cice_b200 import dyn_evp
    return dyn_evp
