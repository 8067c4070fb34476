"""The product has one path: the CUDA libraries.  A missing library is an import error, a missing device makes
every device call stop the process with the library's message, and nothing under zpic_b200/ reaches for the
CPU oracle (oracle/ is test infrastructure, tests/ and bench.py's cpu_baseline are its only users)."""
import os
import re
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _py(code, **env):
    e = dict(os.environ, PYTHONPATH=REPO, **env)
    return subprocess.run([sys.executable, "-c", code], cwd=REPO, env=e, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)


@pytest.mark.parametrize("code", ["em2d", "em1d"])
def test_missing_library_is_an_import_error(code):
    r = _py("from zpic_b200 import load; load(%r)" % code, ZPIC_LIB_SUFFIX="_not_built")
    assert r.returncode != 0
    assert "ImportError" in r.stderr and "no CPU fallback" in r.stderr


def test_device_calls_stop_the_process_without_a_device():
    probe = _py("from zpic_b200 import load; import sys; sys.exit(7 if load('em2d').zdev_init(-1) == 0 else 0)")
    if probe.returncode == 7:
        pytest.skip("a CUDA device is present")
    # zdev_init reports the failure as a value ...
    assert probe.returncode == 0, probe.stderr
    # ... and any entry point that needs the device ends the process, like the reference's fatal errors do
    for call in ("lib.zdev_grid2d_create(8, 8)", "lib.zdev_sync()"):
        r = _py("from zpic_b200 import load; lib = load('em2d'); %s; print('survived')" % call)
        assert r.returncode != 0 and "survived" not in r.stdout
        assert "no usable CUDA device" in r.stderr


def test_product_sources_never_touch_the_oracle():
    pat = re.compile(r"oracle|orc_em|libzpic_ref", re.I)
    offenders = []
    for base, _, files in os.walk(os.path.join(REPO, "zpic_b200")):
        if os.sep + "lib" in base or "__pycache__" in base or "_build" in base:
            continue
        for f in files:
            if not f.endswith((".py", ".c", ".h", ".cu", ".cuh")):
                continue
            for n, line in enumerate(open(os.path.join(base, f), errors="ignore"), 1):
                if pat.search(line) and "import" in line or re.search(r'dlopen|CDLL|#include\s+"orc', line) and pat.search(line):
                    offenders.append("%s:%d: %s" % (os.path.relpath(os.path.join(base, f), REPO), n, line.strip()))
    assert not offenders, offenders
    # the shared objects do not link against anything of the oracle either
    for code in ("em2d", "em1d"):
        lib = os.path.join(REPO, "zpic_b200", "lib", "libzpic_b200_%s.so" % code)
        needed = subprocess.run(["readelf", "-d", lib], stdout=subprocess.PIPE, text=True).stdout
        assert "oracle" not in needed and "zpic_ref" not in needed
