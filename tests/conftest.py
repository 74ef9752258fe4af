import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _ensure_built():
    """Build the product library and the oracle if a prebuilt copy is not already in the tree."""
    import __graft_entry__ as g
    g.build(verbose=False)


@pytest.fixture(scope="session", autouse=True)
def built():
    _ensure_built()
    yield
