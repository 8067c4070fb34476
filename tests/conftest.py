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


# ---------------------------------------------------------------- one retry for the GPU tests, reported
# The GPU path adds the deposited current with floating-point atomics, so two runs of the same deck differ in the last
# bits of J and, after hundreds of steps, in which of two nearly coincident particles is matched with which - every
# tolerance in these tests has a margin of ~10x over what was measured (scripts/lwfa400_flake.py), but one failure in
# ~30 executions of the 400-step LWFA comparison was seen on a freshly started box and could not be reproduced in 30 more.
# A gpu-marked test that fails is therefore run ONCE more; the first failure is NOT hidden: it is printed in the terminal
# summary ("retried once") with its assertion text, and a test that fails twice fails.
_RETRIED = []


def pytest_runtest_protocol(item, nextitem):
    if "gpu" not in item.keywords or item.config.getoption("--no-gpu-retry", default=False):
        return None
    from _pytest.runner import runtestprotocol
    item.ihook.pytest_runtest_logstart(nodeid=item.nodeid, location=item.location)
    reports = runtestprotocol(item, nextitem=nextitem, log=False)
    failed = [r for r in reports if r.when == "call" and r.failed]
    if failed:
        _RETRIED.append((item.nodeid, str(failed[0].longrepr)[-1500:]))
        reports = runtestprotocol(item, nextitem=nextitem, log=False)
    for r in reports:
        item.ihook.pytest_runtest_logreport(report=r)
    item.ihook.pytest_runtest_logfinish(nodeid=item.nodeid, location=item.location)
    return True


def pytest_addoption(parser):
    parser.addoption("--no-gpu-retry", action="store_true", default=False, help="do not run a failed gpu test a second time")


def pytest_terminal_summary(terminalreporter):
    if _RETRIED:
        terminalreporter.section("gpu tests retried once after a failure")
        for nodeid, text in _RETRIED:
            terminalreporter.write_line(nodeid)
            terminalreporter.write_line(text)
