import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "reference: needs the reference tree mounted at /root/reference")


@pytest.fixture(scope="session")
def built():
    """libbya.so + the C mask oracle, built once per session (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g

    g.build()
    return True
