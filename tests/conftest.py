import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure).  Builds liboracle.so on first use."""
    from oracle import oracle as O
    O.build()
    O.lib()
    return O


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def gpu_lib():
    """The CUDA library on a box that has a B200; GPU tests fail (not skip) if it is unusable."""
    from fluidaudio_b200 import _lib
    L = _lib.load()
    assert _lib.device_count() >= 1, "no sm_100a device visible: GPU tests must run on the B200 box"
    return L
