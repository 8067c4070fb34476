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


# ------------------------------------------------------------------ em1d

@pytest.fixture(scope="module")
def ours1():
    from zpic_b200 import load
    return load("em1d")


@pytest.fixture(scope="module")
def ref1():
    from tests import helpers1d as H1
    lib = H1.load_ref()
    if lib is None:
        pytest.skip("oracle/_ref not built")
    return lib


def test_em1d_raw_buffers_and_in_place_edits(ours1, ref1):
    """the em1d twin: the two-stream deck read and edited through the raw buffers of the API, no sync calls"""
    from tests import helpers1d as H1
    assert ours1.zdev_init(-1) == 0
    ours1.zpic_b200_set_option(b"lazy", 0)
    ours1.zpic_b200_set_option(b"coherent", 0)
    a, b = H1.twostream(ours1, nx=120, ppc=64, n_sort=0), H1.twostream(ref1, nx=120, ppc=64, n_sort=0)
    # the two beams' currents cancel: E and J are summation-order noise on the scale of one beam (0.2), see
    # tests/test_gpu_em1d.py::test_twostream_shipped_deck_100_steps
    for chunk in range(3):
        a.iter(5)
        b.iter(5)
        t = 0.1 * 5 * (chunk + 1)
        assert np.abs(a.J() - b.J()).max() < 1e-5 * 0.2
        assert np.abs(a.E() - b.E()).max() < 1e-5 * 0.2 * t
        assert np.abs(a.B() - b.B()).max() < 1e-5 * 0.2 * t
        for k in range(2):
            pa, pb = a.parts(k).copy(), b.parts(k).copy()
            assert len(pa) == len(pb) and np.array_equal(np.sort(pa["ix"]), np.sort(pb["ix"]))
    for d in (a, b):
        d.E()[:, 0] += np.float32(0.01)
        p = d.parts(0)
        p["ux"] += np.float32(0.05)
    a.iter(5)
    b.iter(5)
    assert H.rel_l2(a.E(), b.E()) < 1e-5               # now a real field: the 0.01 offset and the kicked beam's current
    for k in range(2):
        pa, pb = a.parts(k).copy(), b.parts(k).copy()
        assert np.array_equal(np.sort(pa["ix"]), np.sort(pb["ix"]))
        assert abs(float(pa["ux"].astype(np.float64).sum()) - float(pb["ux"].astype(np.float64).sum())) < 1e-3
    a.delete()
    b.delete()


def test_em1d_cython_module_steps_like_the_reference(ours1, ref1):
    """the reference's unmodified em1d.pyx linked to the CUDA library (never run on a GPU before): two-stream deck,
    50 steps, its numpy views against the reference build"""
    import glob
    import importlib.util
    import os
    from tests import helpers1d as H1
    hits = glob.glob(os.path.join(H.REPO, "zpic_b200", "cython", "_build", "em1d.*.so"))
    if not hits:
        pytest.skip("Cython module not built (python -m zpic_b200.cython.build_modules)")
    spec = importlib.util.spec_from_file_location("em1d", hits[0])
    em1d = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(em1d)
    assert ours1.zdev_init(-1) == 0
    ours1.zpic_b200_set_option(b"lazy", 0)
    ours1.zpic_b200_set_option(b"coherent", 0)
    nx, box = 120, float(np.float32(4 * np.pi))
    sp = [em1d.Species("right", -1.0, 64, ufl=[0.2, 0.0, 0.0], uth=[0.001, 0.001, 0.001], n_sort=0),
          em1d.Species("left", -1.0, 64, ufl=[-0.2, 0.0, 0.0], uth=[0.001, 0.001, 0.001], n_sort=0)]
    sim = em1d.Simulation(nx, box, 0.1, species=sp)
    b = H1.twostream(ref1, nx=nx, ppc=64, n_sort=0)
    for chunk in range(5):
        for _ in range(10):
            sim.iter()
        b.iter(10)
        got, want = np.asarray(sim.emf.Ex), b.E()[1:-2, 0]
        assert np.abs(got - want).max() <= 1e-5 * 0.2 * (chunk + 1), chunk      # (cancelling beams: the scale is one beam's)
        for k in range(2):
            pa, pb = np.asarray(sp[k].particles), b.parts(k)
            assert len(pa) == len(pb) and np.array_equal(np.sort(pa["ix"]), np.sort(pb["ix"]))
    assert sim.n == 50
