"""CPU-only: the host side of the product (injector, random stream, laser launch,
struct layout) against the reference build.  No device call is made."""
import ctypes as C

import numpy as np

from tests import helpers as H
from zpic_b200 import abi_em2d as A


def test_struct_sizes_match_reference_headers():
    # sizes measured from the reference headers with gcc (x86-64)
    assert C.sizeof(A.Part) == 28
    assert C.sizeof(A.Species) == 224
    assert C.sizeof(A.Simulation) == 312


def test_weibel_initial_particles_bit_exact(ours, ref):
    a = H.weibel(ours)
    b = H.weibel(ref)
    for k in range(2):
        pa, pb = a.parts(k), b.parts(k)
        assert a.species[k].np == b.species[k].np == 65536
        assert a.species[k].q == b.species[k].q
        assert np.array_equal(pa.view(np.uint8), pb.view(np.uint8))


def test_lwfa_initial_fields_bit_exact(ours, ref):
    a = H.lwfa(ours, nx=(300, 64), box=(6.0, 12.8), laser_start=5.0)
    b = H.lwfa(ref, nx=(300, 64), box=(6.0, 12.8), laser_start=5.0)
    assert a.species[0].np == b.species[0].np == 0
    assert np.abs(b.E()).max() > 1.0
    assert np.array_equal(a.E().view(np.uint32), b.E().view(np.uint32))
    assert np.array_equal(a.B().view(np.uint32), b.B().view(np.uint32))


def test_step_and_slab_injection_bit_exact(ours, ref):
    for dens in (dict(type=A.STEP, start=3.3), dict(type=A.SLAB, start=2.05, end=4.75), dict(type=A.UNIFORM, n=2.0)):
        sp = [dict(name="e", m_q=-1.0, ppc=(3, 2), uth=(0.05, 0.02, 0.01), ufl=(0.1, 0, 0), density=dens)]
        a = H.Deck(ours, (64, 32), (6.4, 3.2), 0.05, sp)
        b = H.Deck(ref, (64, 32), (6.4, 3.2), 0.05, sp)
        assert a.species[0].np == b.species[0].np > 0
        assert np.array_equal(a.parts(0).view(np.uint8), b.parts(0).view(np.uint8))


def _kh_species(lower):
    """Kelvin-Helmholtz shear (BASELINE config 4): each species fills one half of the box in y through a
    CUSTOM density (step function along y) and drifts along +x / -x"""
    def step_y(y, data, lower=lower):
        return 1.0 if ((y < 6.4) == lower) else 0.0
    fn = A.DENSITY_FN(step_y)
    dens = dict(type=A.CUSTOM, custom_y=fn)
    return dict(name="lower" if lower else "upper", m_q=-1.0, ppc=(4, 2), ufl=(0.2 if lower else -0.2, 0, 0),
                uth=(0.01, 0.01, 0.01), density=dens, n_sort=0), fn


def kh_deck(lib, n=64):
    a, fa = _kh_species(True)
    b, fb = _kh_species(False)
    d = H.Deck(lib, (n, n), (12.8, 12.8), 0.07, [a, b])
    d._callbacks = (fa, fb)
    d.set_smooth(xtype=A.BINOMIAL, ytype=A.BINOMIAL, xlevel=2, ylevel=2)
    return d


def test_custom_density_injection_bit_exact(ours, ref):
    a, b = kh_deck(ours), kh_deck(ref)
    for k in range(2):
        assert a.species[k].np == b.species[k].np > 15000
        assert np.array_equal(a.parts(k).view(np.uint8), b.parts(k).view(np.uint8))
    # the trapezoidal inverse-CDF injector smears the step over the boundary cell row
    assert a.parts(0)["iy"].max() == 30 and a.parts(1)["iy"].min() == 31
