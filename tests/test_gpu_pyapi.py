"""zpic_b200.em2d - the reference's Python API (python/source/em2d.pyx: same classes, arguments and properties) over the
CUDA library - against the unmodified reference driven through its C API with the same deck."""
import ctypes as C

import numpy as np
import pytest

from tests import helpers as H
from zpic_b200 import abi_em2d as A

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture()
def em2d(ours):
    ours.zpic_b200_set_option(b"track_ids", 1)
    ours.zpic_b200_set_option(b"lazy", 0)
    ours.zpic_b200_set_option(b"coherent", 0)
    assert ours.zdev_init(-1) == 0
    from zpic_b200 import em2d as mod
    return mod


def _weibel(em2d, n=64, ppc=(2, 2)):
    sp = [em2d.Species("electrons", -1.0, ppc, ufl=[0, 0, 0.6], uth=[0.1, 0.1, 0.1], n_sort=0),
          em2d.Species("positrons", +1.0, ppc, ufl=[0, 0, -0.6], uth=[0.1, 0.1, 0.1], n_sort=0)]
    return em2d.Simulation([n, n], [n * 0.1, n * 0.1], 0.07, species=sp)


def test_weibel_notebook_style_run_matches_the_reference(em2d, ref):
    sim = _weibel(em2d)
    b = H.weibel(ref, n=64, ppc=(2, 2), n_sort=0)
    # iteration 0: the host-side set-up is the reference's, bit for bit
    for k in range(2):
        assert np.array_equal(sim.species[k].particles.view(np.uint8), b.parts(k).view(np.uint8))
    for _ in range(10):
        sim.iter()
    b.iter(10)
    assert sim.n == 10 and abs(sim.t - 10 * 0.07) < 1e-6
    g = b.B()
    assert H.rel_l2(sim.emf.Bz, g[1:65, 1:65, 2]) < TOL
    assert H.rel_l2(sim.emf.Ex, b.E()[1:65, 1:65, 0]) < TOL
    assert H.rel_l2(sim.current.Jz, b.J()[1:65, 1:65, 2]) < TOL
    for k in range(2):
        pa, pb = sim.species[k].particles, b.parts(k)
        assert len(pa) == len(pb)
        assert ((pa["ix"] != pb["ix"]) | (pa["iy"] != pb["iy"])).sum() <= 3
        assert H.rel_l2(pa["uz"], pb["uz"]) < TOL
        assert abs(sim.species[k].energy - b.species[k].energy) <= 1e-6 * abs(b.species[k].energy)
        assert H.rel_l2(sim.species[k].charge(), b.charge(k)[:64, :64]) < 1e-5
    ea, eb = sim.emf.get_energy(), b.emf_energy()
    assert abs(ea.sum() - eb.sum()) <= 5e-6 * eb.sum()


def test_in_place_edits_of_the_views_reach_the_device(em2d, ref):
    """notebooks do `sim.species[0].particles['ux'] += ...` and `sim.emf.Ez[...] = ...` between iterations
    (em2d.pyx:305-312, 1044-1296): the views are host mirrors, the getters mark them as edited"""
    sim = _weibel(em2d, n=32)
    b = H.weibel(ref, n=32, ppc=(2, 2), n_sort=0)
    for step in range(3):
        sim.iter()
        b.iter(1)
        sim.species[0].particles["ux"][::5] += np.float32(0.02)
        b.parts(0)["ux"][::5] += np.float32(0.02)
        sim.emf.Ez[4:9, 6:11] += np.float32(2e-3)
        b.E()[5:10, 7:12, 2] += np.float32(2e-3)
    sim.iter()
    b.iter(1)
    pa, pb = sim.species[0].particles, b.parts(0)
    assert np.array_equal(pa["ix"], pb["ix"]) and H.rel_l2(pa["ux"], pb["ux"]) < 1e-6
    assert H.rel_l2(sim.emf.Ez, b.E()[1:33, 1:33, 2]) < TOL


def test_species_add(em2d, ref):
    """Species.add appends one particle to the buffer (em2d.pyx:238-265)"""
    sim = _weibel(em2d, n=32)
    b = H.weibel(ref, n=32, ppc=(2, 2), n_sort=0)
    sim.iter()
    b.iter(1)
    sim.species[1].add([3, 7], [0.25, 0.5], [0.1, -0.2, 0.3])
    ref.spec_grow_buffer(C.byref(b.species[1]), b.species[1].np + 1)
    b.species[1].part[b.species[1].np] = A.Part(3, 7, 0.25, 0.5, 0.1, -0.2, 0.3)
    b.species[1].np += 1
    sim.iter()
    b.iter(1)
    pa, pb = H.canon(sim.species[1].particles), H.canon(b.parts(1))
    assert len(pa) == len(pb) == 32 * 32 * 4 + 1
    assert np.array_equal(pa["ix"], pb["ix"]) and H.rel_l2(pa["ux"], pb["ux"]) < 1e-6


def test_custom_external_field_from_a_python_callable(em2d, ours, ref):
    """ExternalField(B_type='custom', B_custom=f) with a Python f(ix, dx, iy, dy) (em2d.pyx:481-640): evaluated once per
    cell and handed to the C side as a table; the reference gets the same table through the same C callback"""
    def wire(ix, dx, iy, dy):
        x, y = ix * dx - 3.2, (iy + 0.5) * dy - 3.2
        bx = -y / (x * x + y * y)
        x, y = (ix + 0.5) * dx - 3.2, iy * dy - 3.2
        return (bx, x / (x * x + y * y), 0.0)

    sp = em2d.Species("electrons", -1.0, [2, 2], uth=[0.01, 0.01, 0.01], n_sort=0)
    ext = em2d.ExternalField(B_type="custom", B_custom=wire)
    sim = em2d.Simulation([32, 32], [6.4, 6.4], 0.07, species=sp, ext_fld=ext)
    b = H.Deck(ref, (32, 32), (6.4, 6.4), 0.07, [dict(name="electrons", m_q=-1.0, ppc=(2, 2), uth=(0.01, 0.01, 0.01), n_sort=0)])
    rext = A.ExtField()
    rext.B_type = A.EMF_FLD_TYPE_CUSTOM
    rext.B_custom = ext._c.B_custom                 # the table reader of our library, called by the reference
    rext.B_custom_data = ext._c.B_custom_data
    ref.sim_set_ext_fld(C.byref(b.sim), C.byref(rext))
    for _ in range(10):
        sim.iter()
    b.iter(10)
    gb = A.grid_view(b.sim.emf.ext_fld.B_part_buf, 32, 32)
    assert np.abs(gb).max() > 1.0
    assert H.rel_l2(sim.emf.By_part, gb[1:33, 1:33, 1]) < TOL
    pa, pb = sim.species[0].particles, b.parts(0)
    assert ((pa["ix"] != pb["ix"]) | (pa["iy"] != pb["iy"])).sum() <= 3
    assert H.rel_l2(pa["ux"], pb["ux"]) < TOL
