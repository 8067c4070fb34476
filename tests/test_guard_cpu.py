"""Guarded host mirrors (zpic_b200/csrc/host/common/zb_guard.c) without a GPU: a small C program stands in for the
device side (tests/guard_check.c) - stale reads served by one fill, first writes reported, growth, foreign faults
chained to the handler that was installed before."""
import os
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_guarded_mirrors(tmp_path):
    exe = str(tmp_path / "guard_check")
    src = os.path.join(REPO, "zpic_b200", "csrc", "host", "common")
    subprocess.check_call(["gcc", "-O1", "-std=gnu99", "-I" + src, os.path.join(REPO, "tests", "guard_check.c"),
                           os.path.join(src, "zb_guard.c"), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "guard ok: 3 fills, 2 dirties" in r.stdout


def test_guards_can_be_switched_off(tmp_path):
    """ZPIC_GUARD=0: plain allocations, every state call is a no-op"""
    exe = str(tmp_path / "guard_off")
    src = os.path.join(REPO, "zpic_b200", "csrc", "host", "common")
    prog = tmp_path / "off.c"
    prog.write_text('#include <stdio.h>\n#include "zb_guard.h"\nint main(void){ float* p = zb_guard_alloc(4096);'
                    ' zb_guard_set(p, ZB_G_NONE); p[3] = 1.0f; int st = zb_guard_state(p); zb_guard_free(p);'
                    ' printf("%d %d\\n", zb_guard_enabled(), st); return 0; }\n')
    subprocess.check_call(["gcc", "-O1", "-std=gnu99", "-I" + src, str(prog), os.path.join(src, "zb_guard.c"), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60, env=dict(os.environ, ZPIC_GUARD="0"))
    assert r.returncode == 0 and r.stdout.split() == ["0", "-1"], r.stdout + r.stderr
