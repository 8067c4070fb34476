"""GPU: several slabs on ONE device (LoopbackComm) against the single-domain CUDA run and the reference.
Exercises the CUDA side of the slab decomposition: export lists, column pack/unpack, pass-wise smoothing,
window shift without zeroing, host injection on the last slab."""
import ctypes as C

import numpy as np
import pytest

from tests import helpers as H
from zpic_b200 import abi_em2d as A
from zpic_b200 import parallel as P

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _setup(ours):
    import torch
    assert ours.zdev_init(-1) == 0
    ours.zpic_b200_set_option(b"track_ids", 0)
    ours.zpic_b200_set_option(b"lazy", 0)
    stream = P.share_stream_with_torch(ours)
    yield
    ours.zdev_sync()
    ours.zdev_set_stream(None)
    torch.cuda.set_stream(torch.cuda.default_stream())
    del stream


def _slabs_from_deck(lib, deck, nranks, window, smooth):
    nx, ny = deck.nx
    cfg = [dict(m_q=deck.species[k].m_q, q=deck.species[k].q, ppc=tuple(deck.species[k].ppc)) for k in range(deck.n_species)]
    slabs = []
    for r in range(nranks):
        g = P.Geometry(nx, ny, nranks, r, moving_window=window)
        s = P.CudaSlab(lib, g, deck.sim.dt, deck.sim.emf.dx[0], deck.sim.emf.dx[1], cfg, smooth)
        s.upload_grid(P.E, P.split_grid(deck.E(), g))
        s.upload_grid(P.B, P.split_grid(deck.B(), g))
        for k in range(deck.n_species):
            s.upload_particles(k, P.split_particles(deck.parts(k), g))
        slabs.append(s)
    return slabs


def _global_parts(slabs, k):
    out = []
    for s in slabs:
        p = s.download_particles(k).copy()
        p["ix"] += s.g.x0
        out.append(p)
    return H.canon(np.concatenate(out))


@pytest.mark.parametrize("nranks", [2, 4])
def test_periodic_slabs_vs_reference(ours, ref, nranks):
    steps = 30
    host = H.weibel(ours, n=64, ppc=(2, 2), n_sort=0)        # host-side init only (no device use)
    b = H.weibel(ref, n=64, ppc=(2, 2), n_sort=0)
    slabs = _slabs_from_deck(ours, host, nranks, False, (0, 0, 0, 0))
    hub = P.LoopbackComm.Hub(nranks)
    comms = [P.LoopbackComm(s.g, hub) for s in slabs]
    for _ in range(steps):
        P.step_all(slabs, comms)
    b.iter(steps)
    for which, want in ((P.E, b.E()), (P.B, b.B()), (P.J, b.J())):
        got = P.join_grids([s.download_grid(which) for s in slabs], nranks)
        assert H.rel_l2(got, want[1:-2, 1:-2, :]) < 1e-5, which
    for k in range(2):
        a, r = _global_parts(slabs, k), H.canon(b.parts(k).copy())
        assert len(a) == len(r) == 64 * 64 * 4
        assert (a["ix"] != r["ix"]).sum() + (a["iy"] != r["iy"]).sum() <= 4
    for s in slabs:
        s.destroy()


def test_window_chain_with_injection_vs_reference(ours, ref):
    """laser + moving window + STEP plasma entering through the right edge, 2 slabs (open chain)"""
    kw = dict(nx=(256, 64), box=(5.12, 12.8), dt=0.014, ppc=(2, 2), start=4.0, laser_start=3.5, a0=1.0, n_sort=0)
    host = H.lwfa(ours, **kw)
    b = H.lwfa(ref, **kw)
    nranks = 2
    slabs = _slabs_from_deck(ours, host, nranks, True, (A.COMPENSATED, 0, 4, 0))
    hub = P.LoopbackComm.Hub(nranks)
    comms = [P.LoopbackComm(s.g, hub) for s in slabs]
    inject = P.HostColumnInjector(ours, host.species, slabs[-1].g)
    steps = 120
    for _ in range(steps):
        P.step_all(slabs, comms, inject)
    b.iter(steps)
    assert all(s.n_move == b.sim.emf.n_move for s in slabs) and b.sim.emf.n_move > 0
    a, r = _global_parts(slabs, 0), H.canon(b.parts(0).copy())
    assert len(a) == len(r) > 0
    assert np.array_equal(a["ix"], r["ix"]) and np.array_equal(a["iy"], r["iy"])
    for which, want in ((P.E, b.E()), (P.B, b.B()), (P.J, b.J())):
        got = P.join_grids([s.download_grid(which) for s in slabs], nranks)
        assert H.rel_l2(got, want[1:-2, 1:-2, :]) < 1e-5, which
    for s in slabs:
        s.destroy()
