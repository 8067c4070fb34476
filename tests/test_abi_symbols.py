"""The C-ABI libraries load without a GPU and export every function the public headers declare
(include/em2d, include/em1d = the reference API; include/zpic_dev.h = the device seam; include/zpic_b200.h =
coherence extras).  No compute call is made here."""
import ctypes as C
import os
import re

import pytest

from zpic_b200 import build as zbuild

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROTO = re.compile(r"^[A-Za-z_][\w\s\*]*?[\s\*]([A-Za-z_]\w*)\s*\(", re.M)
NOT_FUNCTIONS = {"if", "for", "while", "switch", "return", "sizeof", "defined", "PHASESPACE"}


def declared(path):
    """names of the functions a header declares (prototypes at file scope; typedef'd callbacks skipped)"""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)              # comments
    src = re.sub(r"//[^\n]*", "", src)
    src = re.sub(r"^\s*#.*?$", "", src, flags=re.M)              # preprocessor lines
    src = src.replace('extern "C" {', "")                        # the C++ guard is not a body
    src = re.sub(r"typedef[^;{]*\([^;]*;", "", src)              # function-pointer typedefs
    src = re.sub(r"\{[^{}]*\}", "", src)                         # struct / enum / inline bodies (one level)
    src = re.sub(r"\{[^{}]*\}", "", src)
    names = set()
    for stmt in src.split(";"):
        stmt = stmt.strip()
        m = PROTO.match(stmt)
        if m and "(" in stmt and not stmt.startswith("typedef"):
            names.add(m.group(1))
    return names - NOT_FUNCTIONS


def headers(sub):
    d = os.path.join(REPO, "include", sub)
    return [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith(".h")]


@pytest.mark.parametrize("code", ["em2d", "em1d"])
def test_library_exports_every_declared_symbol(code):
    lib_path = zbuild.lib_path(code)
    assert os.path.exists(lib_path), "run python __graft_entry__.py first"
    lib = C.CDLL(lib_path)
    dim = "2d" if code == "em2d" else "1d"
    other = "1d" if code == "em2d" else "2d"
    wanted = set()
    for h in headers(code):
        wanted |= declared(h)
    only_2d = re.compile(r"2d|^zdev_(current|emf|yee|smooth)_|^zdev_host_forget$")
    only_1d = re.compile(r"1d")
    for h in (os.path.join(REPO, "include", "zpic_dev.h"), os.path.join(REPO, "include", "zpic_b200.h")):
        for name in declared(h):
            # the seam header covers both codes; the runtime entry points (zdev_init, zdev_sync, ...) are in both
            if code == "em1d" and only_2d.search(name) and not only_1d.search(name):
                continue
            if code == "em2d" and only_1d.search(name):
                continue
            wanted.add(name)
    assert len(wanted) > 60
    missing = sorted(n for n in wanted if not hasattr(lib, n))
    # sim_init / sim_report are IMPORTED from the deck, like in the reference (em2d/main.c:32-36)
    missing = [n for n in missing if n not in ("sim_init", "sim_report")]
    assert not missing, missing


@pytest.mark.parametrize("code", ["em2d", "em1d"])
def test_headers_declare_the_whole_reference_api(code):
    """every function a reference header of the code directory declares is declared by our header of the same
    name (and therefore exported, by the test above); runs where the reference tree is present"""
    ref_dir = os.path.join(os.environ.get("ZPIC_REFERENCE", "/root/reference"), code)
    if not os.path.isdir(ref_dir):
        pytest.skip("reference tree not present")
    checked = 0
    for h in headers(code):
        ref_h = os.path.join(ref_dir, os.path.basename(h))
        assert os.path.exists(ref_h), ref_h
        missing = declared(ref_h) - declared(h)
        assert not missing, (os.path.basename(h), sorted(missing))
        checked += len(declared(ref_h))
    assert checked > 60


def test_host_copy_of_the_charge_deposit_covers_every_byte():
    """zdev_spec2d_par_memcpy splits a large copy over a few threads: every byte must arrive whatever the size
    (67 141 636 bytes = the (4096+1)^2 charge grid, whose size is 4 more than 8 page-aligned chunks - the
    original chunking dropped the last node)"""
    import numpy as np
    lib = C.CDLL(zbuild.lib_path("em2d"))
    lib.zdev_spec2d_par_memcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.zdev_spec2d_par_memcpy.restype = None
    rng = np.random.default_rng(5)
    for nbytes in (0, 1, 4095, 4 << 20, (8 << 20) + 1, 4097 * 4097 * 4, 8 * 2049 * 4096 + 4, 33554432 + 4095, 50000001):
        src = rng.integers(1, 255, nbytes, dtype=np.uint8)
        dst = np.zeros(nbytes + 64, dtype=np.uint8)
        lib.zdev_spec2d_par_memcpy(dst.ctypes.data, src.ctypes.data, nbytes)
        assert np.array_equal(dst[:nbytes], src), nbytes
        assert not dst[nbytes:].any(), nbytes


def test_regrow_plan_of_the_particle_tiles():
    """zdev_spec2d_plan_regrow (host only): the tile layout after a density spike filled a tile - the full tile and the
    tiles up to two away get twice its need, tiles within 10 % of their capacity 1.5 x theirs, nothing shrinks, nothing
    grows past what a push CTA can hold (then the run stops instead), y wraps and x does not"""
    import numpy as np
    lib = C.CDLL(zbuild.lib_path("em2d"))
    f = lib.zdev_spec2d_plan_regrow
    f.restype = C.c_int64
    f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    ntx, nty, cap0 = 8, 6, 5184                                    # 16x16 tiles at 16 per cell, slack 1.25
    off = (np.arange(ntx * nty + 1) * cap0).astype(np.int64)

    def plan(npt, ovf):
        new = np.zeros(ntx * nty + 1, dtype=np.int64)
        bad = C.c_int(-1)
        m = f(ntx, nty, 16, 16, off.ctypes.data, npt.ctypes.data, ovf.ctypes.data, new.ctypes.data, C.byref(bad))
        return m, np.diff(new).reshape(nty, ntx), bad.value

    npt = np.full(ntx * nty, 4096, dtype=np.int32)                 # every tile at 79 % (the nominal fill): nothing to do
    ovf = np.zeros(ntx * nty, dtype=np.int32)
    m, cap, _ = plan(npt, ovf)
    assert m == cap0 and (cap == cap0).all()
    npt[3 + 0 * ntx] = 4800                                        # one tile within 10 % of its capacity: 1.5 x its need, alone
    m, cap, _ = plan(npt, ovf)
    assert cap[0, 3] == (4800 + 2400 + 256 + 31) // 32 * 32 and (cap.ravel() != cap0).sum() == 1
    npt[3 + 0 * ntx] = 4096
    npt[0 + 5 * ntx], ovf[0 + 5 * ntx] = 5184, 100                 # a full tile in the corner x = 0, y = last row
    m, cap, _ = plan(npt, ovf)
    grown = (2 * 5284 + 256 + 31) // 32 * 32
    assert m == grown
    want = np.full((nty, ntx), cap0)
    for y in (3, 4, 5, 0, 1):                                      # two rows either side, wrapping in y
        want[y, 0:3] = grown                                       # ... and two columns to the right only (x does not wrap)
    assert (cap == want).all()
    assert (cap % 32 == 0).all() and (cap >= cap0).all()
    npt[10], ovf[10] = 30000, 50                                   # twice its need would pass the limit: clamped, not fatal
    m, cap, bad = plan(npt, ovf)
    limit = m
    assert 30050 + 32 <= limit < 2 * 30050 and cap.ravel()[10] == limit and limit % 32 == 0
    npt[10] = limit                                                # a tile that NEEDS more than a CTA can take stops the run
    m, cap, bad = plan(npt, ovf)
    assert m == -1 and bad == 10
