import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def nbgpu_lib():
    """The product library; GPU tests must run through it (no fallback)."""
    from nbots_b200 import capi
    L = capi.lib()
    assert L.nbgpu_device_count() > 0, "a GPU test was started without a CUDA device"
    capi.check(L.nbgpu_init(-1))
    return L
