import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_cuda_device():
    import glob
    return bool(glob.glob("/dev/nvidia[0-9]*"))       # (no torch import at collection time)


def pytest_collection_modifyitems(config, items):
    """a plain `pytest` on a machine without a CUDA device skips the gpu-marked tests instead of failing them
    (the product has no CPU path to fall back to; on a GPU box nothing is skipped)"""
    if _have_cuda_device():
        return
    skip = pytest.mark.skip(reason="no CUDA device on this machine (the CUDA path is the only path)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


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
