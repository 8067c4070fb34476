"""The reference's random stream continued on the device (zdev_refrng.cu): the jump-ahead of the two
multiply-with-carry generators (CPU), the deviates against the host generator, and whole species generated on the
device (ZPIC_DEVICE_INIT=2) against the reference's host injector: same particles bit for bit, same stream state
afterwards, same window columns later on."""
import ctypes as C

import numpy as np
import pytest

from tests import helpers as H
from zpic_b200 import abi_em2d as A


def _state(lib):
    z, w, have, spare = C.c_uint32(), C.c_uint32(), C.c_int(), C.c_double()
    lib.zb_rand_get_state(C.byref(z), C.byref(w), C.byref(have), C.byref(spare))
    return z.value, w.value, have.value, spare.value


@pytest.mark.parametrize("seed", [(12345, 67890), (1, 1), (0x464ffffe, 0x9068fffe), (987654321, 2422800382)])
def test_jump_ahead_equals_sequential_draws(ours, seed):
    """z' = a (z & 0xffff) + (z >> 16) is z a mod (a 2^16 - 1): the state after k draws in one modular power
    (reference em2d/random.c:48-53)"""
    ours.rand_uint32.restype = C.c_uint32
    for k in (0, 1, 2, 3, 1000, 65537, 1234567):
        ours.set_rand_seed(*seed)
        for _ in range(k):
            ours.rand_uint32()
        want = _state(ours)[:2]
        ours.set_rand_seed(*seed)
        z, w = C.c_uint32(seed[1]), C.c_uint32(seed[0])        # set_rand_seed(first -> m_w, second -> m_z)
        assert ours.zdev_ref_jump(C.byref(z), C.byref(w), C.c_ulonglong(k)) == 0
        assert (z.value, w.value) == want, k


def test_jump_refuses_states_outside_the_linear_range(ours):
    for z0, w0 in ((0, 5), (5, 0), (36969 * 65536 - 1, 5), (5, 18000 * 65536 - 1), (0xffffffff, 5)):
        z, w = C.c_uint32(z0), C.c_uint32(w0)
        assert ours.zdev_ref_jump(C.byref(z), C.byref(w), C.c_ulonglong(10)) == 1


def _device_normals(lib, count, scale):
    """count deviates from the current host state through the device generator; the host state moves along"""
    import torch
    z, w, have, spare = _state(lib)
    cz, cw, ch, cs = C.c_uint32(z), C.c_uint32(w), C.c_int(have), C.c_double(spare)
    out = torch.zeros(max(count, 1), dtype=torch.float32, device="cuda")
    sc = (C.c_float * 3)(*scale)
    rc = lib.zdev_ref_normals(C.byref(cz), C.byref(cw), C.byref(ch), C.byref(cs), C.c_longlong(count), sc,
                              C.c_void_p(out.data_ptr()))
    assert rc == 0
    lib.zdev_sync()
    lib.zb_rand_set_state(cz, cw, ch, cs)
    return out.cpu().numpy()[:count]


@pytest.mark.gpu
@pytest.mark.parametrize("count", [1, 2, 3, 7, 64, 4097, 300001])
def test_device_deviates_equal_the_host_generator(ours, count):
    """(float) (scale * rand_norm()) for `count` consecutive deviates, twice in a row (the second call starts with
    or without a cached deviate depending on the parity of the first), and the stream state afterwards"""
    assert ours.zdev_init(-1) == 0
    ours.rand_norm.restype = C.c_double
    scale = (0.1, 0.25, 3.0)
    for seed in ((12345, 67890), (777, 31337)):
        ours.set_rand_seed(*seed)
        ours.zb_rand_set_state(C.c_uint32(seed[1]), C.c_uint32(seed[0]), 0, C.c_double(0.0))
        want = []
        for rep in range(2):
            want.append(np.array([np.float32(np.float64(np.float32(scale[m % 3])) * ours.rand_norm()) for m in range(count)],
                                 dtype=np.float32))
        end = _state(ours)
        ours.zb_rand_set_state(C.c_uint32(seed[1]), C.c_uint32(seed[0]), 0, C.c_double(0.0))
        for rep in range(2):
            got = _device_normals(ours, count, scale)
            assert np.array_equal(got.view(np.uint32), want[rep].view(np.uint32)), (seed, rep, int((got != want[rep]).sum()))
        got_end = _state(ours)
        assert got_end[:3] == end[:3]
        if end[2]:
            assert got_end[3] == end[3]


def _same_species(a, b, k):
    pa, pb = a.parts(k), b.parts(k)
    assert len(pa) == len(pb)
    assert np.array_equal(H.canon(pa.copy()).view(np.uint8), H.canon(pb.copy()).view(np.uint8))


@pytest.mark.gpu
@pytest.mark.parametrize("n,ppc", [(96, (3, 2)), (512, (4, 4))])
def test_weibel_species_generated_on_the_device_equal_the_reference(ours, ref, n, ppc):
    """two warm species in a square box (em2d/input/weibel.c; 8.4 M particles in the larger case): every particle
    of both species bit-identical to the reference's host injector, the host stream ends where the reference's
    does, and after one step every particle is still bit-identical (nothing was ever uploaded)"""
    assert ours.zdev_init(-1) == 0
    ours.zpic_b200_set_option(b"device_init", 2)
    ours.zpic_b200_set_option(b"lazy", 0)
    try:
        a = H.weibel(ours, n=n, ppc=ppc, n_sort=0)
        b = H.weibel(ref, n=n, ppc=ppc, n_sort=0)
        for k in range(2):
            assert a.species[k].np == b.species[k].np
            assert not a.species[k].part                       # nothing was generated on the host
        ours.rand_uint32.restype = ref.rand_uint32.restype = C.c_uint32
        assert ours.rand_uint32() == ref.rand_uint32()        # the stream is where the reference left it
        if n <= 96:
            a.sync()
            for k in range(2):
                _same_species(a, b, k)
        a.iter(1)
        b.iter(1)
        a.sync()
        for k in range(2):
            _same_species(a, b, k)
        a.iter(2)
        b.iter(2)
        s = a.snapshot()
        for k in range(2):
            assert s["np"][k] == b.species[k].np
        for name, want in (("E", b.E()), ("B", b.B()), ("J", b.J())):
            assert H.rel_l2(s[name], want) < 1e-5, name
        a.delete()
        b.delete()
    finally:
        ours.zpic_b200_set_option(b"device_init", 0)


@pytest.mark.gpu
@pytest.mark.parametrize("dens", [dict(type=A.STEP, start=3.13), dict(type=A.SLAB, start=2.07, end=9.61), dict(type=A.UNIFORM)])
def test_clipped_cold_plasma_and_window_columns_follow_the_reference(ours, ref, dens):
    """a cold plasma clipped inside a cell (STEP / SLAB) in a non-square box under a moving window: the initial
    population comes from the device, the window columns from the host injector on the same stream"""
    assert ours.zdev_init(-1) == 0
    ours.zpic_b200_set_option(b"device_init", 2)
    ours.zpic_b200_set_option(b"lazy", 0)
    try:
        decks = []
        for lib in (ours, ref):
            sp = [dict(name="electrons", m_q=-1.0, ppc=(3, 2), density=dict(dens))]
            # dx[0] == dx[1]: the reference sizes a SLAB's buffer with dx[1] (particles.c:390) and overruns it otherwise
            d = H.Deck(lib, (200, 48), (12.0, 2.88), 0.03, sp)
            d.add_laser(type=A.PLANE, start=11.0, fwhm=2.0, a0=1.0, omega0=8.0, polarization=np.pi / 2)
            d.set_moving_window()
            d.set_smooth(xtype=A.COMPENSATED, xlevel=2)
            decks.append(d)
        a, b = decks
        assert a.species[0].np == b.species[0].np
        assert not a.species[0].part
        a.sync()
        _same_species(a, b, 0)
        a.iter(25)
        b.iter(25)
        s = a.snapshot()
        assert s["np"][0] == b.species[0].np
        pa, pb = H.canon(s["parts"][0]), H.canon(b.parts(0).copy())
        assert np.array_equal(pa["ix"], pb["ix"]) and np.array_equal(pa["iy"], pb["iy"])
        for name, want in (("E", b.E()), ("B", b.B())):
            assert H.rel_l2(s[name], want) < 1e-5, name
        ours.rand_uint32.restype = ref.rand_uint32.restype = C.c_uint32
        assert ours.rand_uint32() == ref.rand_uint32()
        a.delete()
        b.delete()
    finally:
        ours.zpic_b200_set_option(b"device_init", 0)


@pytest.mark.gpu
def test_warm_plasma_in_a_non_square_box_takes_the_host_injector(ours, ref):
    """the reference's cell means mix cells when nx[0] != nx[1] (its accumulator index uses nx[1] as stride): such a
    species is generated on the host even with device_init = 2, and still equals the reference"""
    assert ours.zdev_init(-1) == 0
    ours.zpic_b200_set_option(b"device_init", 2)
    try:
        sp = [dict(name="warm", m_q=-1.0, ppc=(2, 2), uth=(0.01, 0.02, 0.03), ufl=(0.1, 0, 0))]
        a = H.Deck(ours, (64, 32), (6.4, 3.2), 0.05, sp)
        b = H.Deck(ref, (64, 32), (6.4, 3.2), 0.05, sp)
        assert a.species[0].part
        assert np.array_equal(a.parts(0).view(np.uint8), b.parts(0).view(np.uint8))
        a.delete()
        b.delete()
    finally:
        ours.zpic_b200_set_option(b"device_init", 0)


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


@pytest.mark.gpu
def test_em1d_twostream_beams_generated_on_the_device_equal_the_reference(ours1, ref1):
    """em1d/input/twostream.c: both beams (2 x 30 000 warm particles) bit-identical to the reference's host injector,
    stream position equal afterwards, still bit-identical after one step"""
    from tests import helpers1d as H1
    assert ours1.zdev_init(-1) == 0
    ours1.zpic_b200_set_option(b"device_init", 2)
    ours1.zpic_b200_set_option(b"lazy", 0)
    try:
        a, b = H1.twostream(ours1, nx=120, ppc=250, n_sort=0), H1.twostream(ref1, nx=120, ppc=250, n_sort=0)
        for k in range(2):
            assert a.species[k].np == b.species[k].np and not a.species[k].part
        ours1.rand_uint32.restype = ref1.rand_uint32.restype = C.c_uint32
        assert ours1.rand_uint32() == ref1.rand_uint32()
        a.sync()
        order = ["ix", "x", "ux", "uy", "uz"]
        for k in range(2):
            pa, pb = np.sort(a.parts(k).copy(), order=order), np.sort(b.parts(k).copy(), order=order)
            assert np.array_equal(pa.view(np.uint8), pb.view(np.uint8))
        a.iter(1)
        b.iter(1)
        a.sync()
        for k in range(2):
            pa, pb = np.sort(a.parts(k).copy(), order=order), np.sort(b.parts(k).copy(), order=order)
            assert np.array_equal(pa.view(np.uint8), pb.view(np.uint8))
        a.delete()
        b.delete()
    finally:
        ours1.zpic_b200_set_option(b"device_init", 0)


@pytest.mark.gpu
def test_em1d_clipped_plasma_under_a_moving_window(ours1, ref1):
    """em1d/input/movwindow.c pattern: a STEP plasma that starts inside a cell, window columns from the host injector"""
    from tests import helpers1d as H1
    from zpic_b200 import abi_em1d as A1
    assert ours1.zdev_init(-1) == 0
    ours1.zpic_b200_set_option(b"device_init", 2)
    ours1.zpic_b200_set_option(b"lazy", 0)
    try:
        decks = []
        for lib in (ours1, ref1):
            sp = dict(name="electrons", m_q=-1.0, ppc=32, density=dict(type=A1.STEP, start=20.013), n_sort=0)
            d = H1.Deck1D(lib, 512, 41.0, 0.07, [sp], tmax=300.0, ndump=50)
            d.add_laser(start=25.0, fwhm=7.0, a0=0.5, omega0=10.0, polarization=np.pi / 2)
            d.set_moving_window()
            decks.append(d)
        a, b = decks
        assert a.species[0].np == b.species[0].np and not a.species[0].part
        a.iter(60)
        b.iter(60)
        a.sync()
        assert a.species[0].np == b.species[0].np and a.sim.emf.n_move == b.sim.emf.n_move > 3
        pa, pb = np.sort(a.parts(0).copy(), order=["ix", "x", "ux"]), np.sort(b.parts(0).copy(), order=["ix", "x", "ux"])
        assert np.array_equal(pa["ix"], pb["ix"])
        assert H.rel_l2(a.E(), b.E()) < 1e-5
        ours1.rand_uint32.restype = ref1.rand_uint32.restype = C.c_uint32
        assert ours1.rand_uint32() == ref1.rand_uint32()
        a.delete()
        b.delete()
    finally:
        ours1.zpic_b200_set_option(b"device_init", 0)
