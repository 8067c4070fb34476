"""Guarded host mirrors on the GPU (csrc/host/common/zb_guard.c): callers that read and write the raw buffers of the
reference API between iterations - sim->emf.E_buf, species[i].part, as the reference's decks and its Cython module
do - see reference semantics WITHOUT any zpic_b200_sync_* / touch_* call and without the round trip of
ZPIC_COHERENT=1: a stale mirror is downloaded when it is first touched, a modified one goes up before the next step."""
import ctypes as C

import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu


def _counts(lib):
    lib.zb_guard_fills.restype = lib.zb_guard_dirties.restype = C.c_ulong
    return lib.zb_guard_fills(), lib.zb_guard_dirties()


@pytest.mark.parametrize("n,ppc", [(48, (2, 2)), (320, (1, 1))])       # 320: the E, B, J mirrors are page-locked (>= 1 MB)
def test_raw_buffers_follow_the_device_without_sync_calls(ours, ref, n, ppc):
    assert ours.zdev_init(-1) == 0
    assert ours.zb_guard_enabled() == 1
    ours.zpic_b200_set_option(b"lazy", 0)
    ours.zpic_b200_set_option(b"coherent", 0)
    a = H.weibel(ours, n=n, ppc=ppc, n_sort=0)
    b = H.weibel(ref, n=n, ppc=ppc, n_sort=0)
    f0, d0 = _counts(ours)
    for step in range(3):
        a.iter(2)
        b.iter(2)
        # no sync: the views are the raw buffers
        for name, got, want in (("E", a.E(), b.E()), ("B", a.B(), b.B()), ("J", a.J(), b.J())):
            assert H.rel_l2(got, want) < 1e-5, (step, name)
        for k in range(2):
            pa, pb = H.canon(a.parts(k).copy()), H.canon(b.parts(k).copy())
            assert len(pa) == len(pb)
            assert (pa["ix"] != pb["ix"]).sum() + (pa["iy"] != pb["iy"]).sum() == 0
    f1, d1 = _counts(ours)
    assert f1 - f0 == 3 * 4             # per visit: E+B together, J, and the two species
    assert d1 == d0                     # nobody wrote
    # steps whose mirrors nobody looks at transfer nothing
    a.iter(5)
    b.iter(5)
    assert _counts(ours)[0] == f1
    assert H.rel_l2(a.E(), b.E()) < 1e-5
    a.delete()
    b.delete()


def test_in_place_edits_of_raw_buffers_reach_the_device(ours, ref):
    """what a notebook does: sim.emf.Ez[...] += ..., particles['ux'] *= ... between iterations"""
    assert ours.zdev_init(-1) == 0
    ours.zpic_b200_set_option(b"lazy", 0)
    ours.zpic_b200_set_option(b"coherent", 0)
    a = H.weibel(ours, n=48, ppc=(2, 2), n_sort=0)
    b = H.weibel(ref, n=48, ppc=(2, 2), n_sort=0)
    a.iter(3)
    b.iter(3)
    _, d0 = _counts(ours)
    yy, xx = np.mgrid[0:51, 0:51]
    bump = (0.05 * np.sin(0.3 * xx) * np.cos(0.2 * yy)).astype(np.float32)
    for d in (a, b):
        d.E()[:, :, 2] += bump                      # read-modify-write of a stale mirror
        p = d.parts(1)
        p["ux"] *= np.float32(1.5)
    assert _counts(ours)[1] - d0 == 2               # the E mirror and one species became dirty
    a.iter(4)
    b.iter(4)
    for name, got, want in (("E", a.E(), b.E()), ("B", a.B(), b.B())):
        assert H.rel_l2(got, want) < 1e-5, name
    for k in range(2):
        pa, pb = H.canon(a.parts(k).copy()), H.canon(b.parts(k).copy())
        assert np.array_equal(pa["ix"], pb["ix"]) and np.array_equal(pa["iy"], pb["iy"])
        assert H.rel_l2(pa["ux"], pb["ux"]) < 1e-5
    a.delete()
    b.delete()


def test_window_run_with_a_growing_population(ours, ref):
    """moving window: the particle mirror has to grow between the steps (a fault handler cannot move it)"""
    assert ours.zdev_init(-1) == 0
    ours.zpic_b200_set_option(b"lazy", 0)
    ours.zpic_b200_set_option(b"coherent", 0)
    kw = dict(nx=(256, 32), box=(5.12, 6.4), dt=0.014, ppc=(2, 2), start=5.12, laser_start=4.2, a0=1.0)
    a, b = H.lwfa(ours, **kw), H.lwfa(ref, **kw)
    for _ in range(6):
        a.iter(25)
        b.iter(25)
        assert a.species[0].np == b.species[0].np
        pa, pb = H.canon(a.parts(0).copy()), H.canon(b.parts(0).copy())
        assert len(pa) == len(pb) and np.array_equal(pa["ix"], pb["ix"]) and np.array_equal(pa["iy"], pb["iy"])
        assert H.rel_l2(a.E(), b.E()) < 1e-5
    assert a.species[0].np > 0
    a.delete()
    b.delete()
