"""CPU-only: pin the restatement in oracle/ against the unmodified reference (oracle/_ref) and
against the golden vectors in tests/golden (generated from the reference by
tests/golden/make_golden.py).  Everything must be bit-identical."""
import ctypes as C
import os

import numpy as np
import pytest

from tests import helpers as H
from tests import oracle as O
from zpic_b200 import abi_em2d as A

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _same(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


def test_oracle_weibel_matches_reference_bit_for_bit(ref):
    """40 steps of the Weibel deck incl. two particle sorts (iter 16, 32)"""
    d = H.weibel(ref, n=32, ppc=(2, 2))
    o = O.OracleSim(d)
    for step in (1, 15, 16, 40):
        d.iter(step - d.sim.emf.iter)
        o.iter(step - o.sim.iter)
        for k in range(2):
            assert _same(o.part(k), d.parts(k)), (step, k)
            assert o.energy(k) == d.species[k].energy
        assert _same(o.E, d.E()) and _same(o.B, d.B()) and _same(o.J, d.J()), step


def test_oracle_moving_window_matches_reference(ref):
    """laser + moving window + absorbing boundary + compensated smoothing; the slab ends inside
    the box so the window injects nothing (injection needs the host random stream)"""
    dens = dict(type=A.SLAB, start=1.0, end=3.0)
    sp = [dict(name="e", m_q=-1.0, ppc=(2, 2), uth=(0.01, 0.01, 0.01), density=dens, n_sort=0)]
    # equal cell sizes: the reference sizes the SLAB buffer with dx[1] (em2d/particles.c:390)
    d = H.Deck(ref, (200, 48), (4.0, 0.96), 0.012, sp)
    d.add_laser(type=A.GAUSSIAN, start=3.8, fwhm=1.0, a0=1.5, omega0=10.0, W0=0.3, focus=5.0, axis=0.48,
                polarization=np.pi / 2)
    d.set_moving_window()
    d.set_smooth(xtype=A.COMPENSATED, xlevel=4)
    o = O.OracleSim(d)
    n0 = d.species[0].np
    for step in (1, 60, 150):
        d.iter(step - d.sim.emf.iter)
        o.iter(step - o.sim.iter)
        assert o.sim.n_move == d.sim.emf.n_move
        assert o.spec[0].np == d.species[0].np
        assert _same(o.part(0), d.parts(0)), step
        assert _same(o.E, d.E()) and _same(o.B, d.B()) and _same(o.J, d.J()), step
    assert d.sim.emf.n_move > 50 and d.species[0].np < n0


@pytest.mark.parametrize("smooth", [(1, 1, 2, 2), (2, 2, 3, 1), (0, 1, 0, 1), (2, 0, 1, 0)])
def test_oracle_current_update_matches_reference(ref, smooth):
    rng = np.random.default_rng(5)
    nx, ny = 37, 29
    for mw in (0, 1):
        d = H.Deck(ref, (nx, ny), (3.7, 2.9), 0.05)
        j = rng.standard_normal(d.J().shape).astype(np.float32)
        d.J()[...] = j
        d.sim.current.moving_window = mw
        d.sim.current.smooth = A.Smooth(*smooth)
        ref.current_update(C.byref(d.sim.current))
        mine = j.copy()
        L = O.lib()
        L.orc2d_current_gc(mine.ctypes.data_as(C.c_void_p), nx, ny, mw)
        L.orc2d_current_smooth(mine.ctypes.data_as(C.c_void_p), nx, ny, mw, *smooth)
        assert _same(mine, d.J())


def test_oracle_charge_and_energy_match_reference(ref):
    d = H.weibel(ref, n=32, ppc=(2, 2), n_sort=0)
    d.iter(5)
    L = O.lib()
    for k in range(2):
        p = d.parts(k).copy()
        rho = np.zeros((33, 33), dtype=np.float32)
        L.orc2d_deposit_charge(p.ctypes.data_as(C.c_void_p), len(p), C.c_float(d.species[k].q), 32, 32, 0,
                               rho.ctypes.data_as(C.c_void_p))
        assert _same(rho, d.charge(k))
    e = (C.c_double * 6)()
    Ec, Bc = d.E().copy(), d.B().copy()
    L.orc2d_emf_energy(Ec.ctypes.data_as(C.c_void_p), Bc.ctypes.data_as(C.c_void_p), 32, 32, e)
    want = d.emf_energy()
    got = np.array(e[:]) * ((0.5 * d.sim.emf.dx[0]) * d.sim.emf.dx[1])     # the reference's association
    assert np.array_equal(got, want)


def test_oracle_against_golden_vectors():
    """no reference needed: vectors committed under tests/golden (made from the reference)"""
    path = os.path.join(GOLD, "weibel_16x16.npz")
    g = np.load(path)
    L = O.lib()
    nx = ny = int(g["nx"])
    E, B = g["E0"].copy(), g["B0"].copy()
    J = np.zeros_like(E)
    parts = [g["part0_s0"].copy(), g["part0_s1"].copy()]
    spec = (O.OrcSpecies * 2)()
    for k in range(2):
        spec[k].part = parts[k].ctypes.data
        spec[k].np = len(parts[k])
        spec[k].m_q = float(g["m_q"][k])
        spec[k].q = float(g["q"][k])
        spec[k].n_sort = 16
    sim = O.OrcSim(nx, ny, float(g["dx"]), float(g["dx"]), float(g["dt"]), E.ctypes.data, B.ctypes.data,
                   J.ctypes.data, 0, 0, 0, 0, 0, 0, 0, 2, spec)
    for _ in range(int(g["steps"])):
        L.orc2d_sim_iter(C.byref(sim))
    assert _same(E, g["E1"]) and _same(B, g["B1"]) and _same(J, g["J1"])
    for k in range(2):
        assert _same(parts[k], g["part1_s%d" % k])
        assert spec[k].energy == float(g["energy"][k])
