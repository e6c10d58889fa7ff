import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def hs():
    """The product package; the CUDA library must be present (no fallback)."""
    import hyperelasticsolver_b200 as H
    H.lib()
    return H


@pytest.fixture(scope="session")
def gpu(hs):
    if hs.lib().hs_device_count() <= 0:
        pytest.fail("-m gpu tests need a CUDA device; none visible (the library has no CPU fallback)")
    return hs
