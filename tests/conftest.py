import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running CPU test, not in the default selection")


@pytest.fixture(scope="session")
def oracle():
    """ctypes handle on the CPU oracle (built on demand; test infrastructure)."""
    import oracle_lib
    return oracle_lib.load()


@pytest.fixture(scope="session")
def sq():
    """The product: ctypes binding over libsquava_b200.so."""
    import squava_b200
    return squava_b200
