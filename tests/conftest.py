import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def ctx():
    """One engine context per test session; fails loudly without a GPU or the built .so."""
    import raven_b200.backend as B
    c = B.create_context()
    yield c
    c.sync()


@pytest.fixture(scope="session")
def oracle():
    from tests import harness
    return harness.get_oracle()
