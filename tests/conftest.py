import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the C oracle and the CUDA library exist (both compile without a GPU)."""
    import __graft_entry__ as g
    from blockpuzzle_gym_b200 import _lib
    from oracle import coracle
    if not os.path.exists(_lib.LIB_PATH) or not os.path.exists(coracle._SO):
        g.build()
    yield
