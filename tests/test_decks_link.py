"""Drop-in check of INTEGRATION.md section 1 on the CPU: the reference's UNMODIFIED main.c and every shipped
input deck, compiled where they lie against the reference's own headers (the decks include "../simulation.h"),
link against libzpic_b200_<code>.so instead of the reference objects (link-time symbol replacement,
SURVEY.md 8b; the struct layouts of include/<code> are the reference's, tests/test_host_init.py).  Without a CUDA device the executable must stop with the
library's error message - there is no CPU fallback to fall into.  Needs the reference tree (decks and main.c
are read where they lie, nothing is copied into the repository); skipped on machines without it."""
import os
import re
import subprocess

import pytest

from zpic_b200 import build as zbuild

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("ZPIC_REFERENCE", "/root/reference")


def _decks():
    out = []
    for code in ("em2d", "em1d"):
        d = os.path.join(REF, code, "input")
        if os.path.isdir(d):
            out += [(code, f) for f in sorted(os.listdir(d)) if f.endswith(".c")]
    return out or [("em2d", None)]


def _have_cuda():
    from zpic_b200 import load
    return load("em2d").zdev_init(-1) == 0


@pytest.mark.parametrize("code,deck", _decks())
def test_reference_main_and_deck_link_against_the_library(code, deck, tmp_path):
    if deck is None:
        pytest.skip("reference tree not present")
    lib = zbuild.lib_path(code)
    assert os.path.exists(lib), "run python __graft_entry__.py first"
    # main.c selects its deck with an #include line (em2d/main.c:32-36): point that line at the deck under test
    main_src = open(os.path.join(REF, code, "main.c")).read()
    main_src, n = re.subn(r'^#include "input/[\w\-]+\.c"', '#include "%s"' % os.path.join(REF, code, "input", deck),
                          main_src, count=1, flags=re.M)
    assert n == 1
    main_c = tmp_path / "main.c"
    main_c.write_text(main_src)
    exe = tmp_path / "zpic"
    cmd = ["gcc", "-std=c99", "-O1", "-I" + os.path.join(REF, code), str(main_c), "-o", str(exe), "-L" + os.path.dirname(lib),
           "-l" + os.path.basename(lib)[3:-3], "-Wl,-rpath," + os.path.dirname(lib), "-lm"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    # every zpic symbol the executable needs is resolved by our library (nothing left for a reference object)
    undefined = subprocess.run(["nm", "-u", str(exe)], stdout=subprocess.PIPE, text=True).stdout
    wanted = set(re.findall(r"\bU (\w+)", undefined))
    assert "sim_iter" in wanted and "sim_new" in wanted
    if _have_cuda():
        return                                  # with a GPU the deck would run to its tmax: not a unit test
    run = subprocess.run([str(exe)], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert run.returncode != 0
    assert "no usable CUDA device" in run.stderr and "no CPU path" in run.stderr
    # the host-side initial state may have been reported (iteration 0 lives in host memory); no step was taken
    assert all(p.name.endswith("-000000.zdf") for p in tmp_path.rglob("*.zdf"))
