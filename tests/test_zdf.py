"""ZDF output: files written by the product must be byte-compatible with the reference writer and
readable by our reader (and by the reference's own python/lib/zdf.py when it is around)."""
import ctypes as C
import filecmp
import os
import sys

import numpy as np
import pytest

from tests import helpers as H
from zpic_b200 import abi_em2d as A
from zpic_b200 import zdf


def _report_set(deck):
    """the diagnostics of the shipped Weibel deck that need no device (reference input/weibel.c:44-57)"""
    L = deck.lib
    L.emf_report(C.byref(deck.sim.emf), bytes([A.BFLD]), 0)
    L.emf_report(C.byref(deck.sim.emf), bytes([A.EFLD]), 2)
    L.current_report(C.byref(deck.sim.current), 2)
    L.spec_report(deck.species[0], 0x3000, None, None)          # PARTICLES
    nx = (C.c_int * 2)(64, 32)
    rng = ((C.c_float * 2) * 2)((C.c_float * 2)(0.0, 3.2), (C.c_float * 2)(-1.0, 1.0))
    L.spec_report(deck.species[1], 0x2000 + 1 + 16 * 6, nx, rng)  # PHASESPACE(X1, U3)


def test_iteration0_files_identical_to_reference(ours, ref, tmp_path):
    cwd = os.getcwd()
    dirs = {}
    try:
        for name, lib in (("ours", ours), ("ref", ref)):
            d = tmp_path / name
            d.mkdir()
            os.chdir(d)
            deck = H.weibel(lib, n=32, ppc=(2, 2))
            _report_set(deck)
            dirs[name] = d
    finally:
        os.chdir(cwd)
    files = sorted(p.relative_to(dirs["ref"]) for p in dirs["ref"].rglob("*.zdf"))
    assert len(files) == 5
    for f in files:
        assert (dirs["ours"] / f).exists(), f
        assert filecmp.cmp(dirs["ours"] / f, dirs["ref"] / f, shallow=False), f


def test_reader_round_trip(ours, tmp_path):
    cwd = os.getcwd()
    try:
        os.chdir(tmp_path)
        deck = H.weibel(ours, n=32, ppc=(2, 2))
        _report_set(deck)
    finally:
        os.chdir(cwd)
    data, info = zdf.read(str(tmp_path / "PARTICLES" / "electrons" / "particles-electrons-000000.zdf"))
    assert info.type == "particles" and info.particles.nparts == 32 * 32 * 4
    p = deck.parts(0)
    dx = deck.sim.emf.dx[0]
    assert np.array_equal(data["ux"], p["ux"])
    assert np.array_equal(data["x"], ((p["ix"] + p["x"]) * np.float32(dx)).astype(np.float32))
    data, info = zdf.read(str(tmp_path / "PHASESPACE" / "positrons" / "positrons-x1u3-000000.zdf"))
    assert info.type == "grid" and data.shape == (32, 64)
    assert info.grid.axis[1].min == -1.0 and info.iteration.n == 0
    assert 0.9 * 1024 < data.sum() <= 1024.0          # total charge, minus the weight falling off the grid edges
    ref_reader = "/root/reference/python/lib"
    if os.path.isdir(ref_reader):
        sys.path.insert(0, ref_reader)
        import zdf as refzdf          # the reference's reader must parse our files too
        d2, i2 = refzdf.read(str(tmp_path / "PHASESPACE" / "positrons" / "positrons-x1u3-000000.zdf"))
        sys.path.remove(ref_reader)
        assert np.array_equal(d2, data) and i2.grid.nx[0] == 64
