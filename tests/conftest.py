import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def capi():
    """The product C ABI; GPU tests fail loudly when the library or the device is missing."""
    from plinopt_b200 import capi as c
    c.lib()
    if c.device_count() < 1:
        pytest.fail("no CUDA device visible: the engine has no CPU fallback")
    return c
