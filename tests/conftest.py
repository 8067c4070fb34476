import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ours():
    """the product library (CUDA); fails loudly if it was not built"""
    from zpic_b200 import load
    return load("em2d")


@pytest.fixture(scope="session")
def ref():
    """the unmodified reference, strict build (oracle/_ref, tests only)"""
    from tests import helpers
    lib = helpers.load_ref("em2d")
    if lib is None:
        pytest.skip("oracle/_ref not built (run `make -C oracle ref` where /root/reference exists)")
    return lib
