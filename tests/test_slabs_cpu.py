"""CPU-only: the slab-decomposition logic of zpic_b200.parallel, exercised with the oracle as the local
"device" - in one process through LoopbackComm, and with two real processes over gloo (world_size 2)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import helpers as H
from tests import oracle as O
from tests.oracle_slab import OracleSlab
from zpic_b200 import abi_em2d as A
from zpic_b200 import parallel as P


def _species_cfg(deck):
    return [dict(m_q=deck.species[k].m_q, q=deck.species[k].q, ppc=tuple(deck.species[k].ppc)) for k in range(deck.n_species)]


def _make_slabs(deck, nranks, window, smooth=(0, 0, 0, 0), ranks=None):
    nx, ny = deck.nx
    slabs = []
    for r in (ranks if ranks is not None else range(nranks)):
        g = P.Geometry(nx, ny, nranks, r, moving_window=window)
        s = OracleSlab(g, deck.sim.dt, deck.sim.emf.dx[0], deck.sim.emf.dx[1], _species_cfg(deck), smooth)
        s.upload_grid(P.E, P.split_grid(deck.E(), g))
        s.upload_grid(P.B, P.split_grid(deck.B(), g))
        for k in range(deck.n_species):
            s.upload_particles(k, P.split_particles(deck.parts(k), g))
        slabs.append(s)
    return slabs


def _global_particles(slabs, k):
    parts = []
    for s in slabs:
        p = s.download_particles(k).copy()
        p["ix"] += s.g.x0
        parts.append(p)
    return H.canon(np.concatenate(parts))


def _check_against_single(deck_factory, nranks, steps, window, smooth=(0, 0, 0, 0)):
    ref = H.load_ref("em2d")
    if ref is None:
        pytest.skip("oracle/_ref not built")
    deck = deck_factory(ref)
    single = O.OracleSim(deck, n_sort=0)
    slabs = _make_slabs(deck, nranks, window, smooth)
    hub = P.LoopbackComm.Hub(nranks)
    comms = [P.LoopbackComm(s.g, hub) for s in slabs]
    for _ in range(steps):
        single.iter(1)
        P.step_all(slabs, comms)
    nxl = deck.nx[0] // nranks
    for which, want in ((P.E, single.E), (P.B, single.B), (P.J, single.J)):
        got = P.join_grids([s.download_grid(which) for s in slabs], nranks)
        want = want[1:-2, 1:-2, :]
        # cells next to a slab edge collect current from two ranks: summation order only
        assert H.rel_l2(got, want) < 1e-5, which
        # everywhere else the decomposed run is bit-identical to the single domain
        inner = np.ones(deck.nx[0], dtype=bool)
        for r in range(nranks + 1):
            inner[max(r * nxl - 3, 0): r * nxl + 3] = False
        assert np.array_equal(got[:, inner, :], want[:, inner, :]), which
    for k in range(deck.n_species):
        a, b = _global_particles(slabs, k), H.canon(single.part(k).copy())
        assert len(a) == len(b)
        assert np.array_equal(a["ix"], b["ix"]) and np.array_equal(a["iy"], b["iy"])
        assert H.rel_l2(a["ux"], b["ux"]) < 1e-6


def _weibel(ref):
    return H.weibel(ref, n=32, ppc=(2, 2), n_sort=0)


def _window_deck(ref):
    dens = dict(type=A.SLAB, start=1.0, end=3.0)
    sp = [dict(name="e", m_q=-1.0, ppc=(2, 2), uth=(0.02, 0.02, 0.02), density=dens, n_sort=0)]
    d = H.Deck(ref, (192, 32), (3.84, 0.64), 0.012, sp)
    d.add_laser(type=A.GAUSSIAN, start=3.6, fwhm=1.0, a0=1.5, omega0=10.0, W0=0.2, focus=5.0, axis=0.32,
                polarization=np.pi / 2)
    d.set_moving_window()
    d.set_smooth(xtype=A.COMPENSATED, xlevel=4)
    return d


@pytest.mark.parametrize("nranks", [2, 4])
def test_periodic_slabs_equal_single_domain(nranks):
    # one step: J of the first step is zero-field driven, particles have not diverged: exact cells
    _check_against_single(_weibel, nranks, 1, window=False)


def test_periodic_slabs_many_steps_with_smoothing():
    def deck(ref):
        d = _weibel(ref)
        d.set_smooth(xtype=A.BINOMIAL, ytype=A.BINOMIAL, xlevel=2, ylevel=2)
        return d
    _check_against_single_loose(deck, 4, 25, window=False, smooth=(1, 1, 2, 2))


def _check_against_single_loose(deck_factory, nranks, steps, window, smooth):
    """after many steps J (hence E, B, u) differs by summation order: tolerance instead of equality"""
    ref = H.load_ref("em2d")
    if ref is None:
        pytest.skip("oracle/_ref not built")
    deck = deck_factory(ref)
    single = O.OracleSim(deck, n_sort=0)
    slabs = _make_slabs(deck, nranks, window, smooth)
    hub = P.LoopbackComm.Hub(nranks)
    comms = [P.LoopbackComm(s.g, hub) for s in slabs]
    for _ in range(steps):
        single.iter(1)
        P.step_all(slabs, comms)
    scale = np.sqrt((single.E.astype(np.float64) ** 2).sum())
    for which, want in ((P.E, single.E), (P.B, single.B), (P.J, single.J)):
        got = P.join_grids([s.download_grid(which) for s in slabs], nranks)
        want = want[1:-2, 1:-2, :]
        if which == P.B:
            # B is still orders of magnitude below E this early: measure its error on the field scale
            err = np.sqrt(((got.astype(np.float64) - want) ** 2).sum()) / scale
        else:
            err = H.rel_l2(got, want)
        assert err < 1e-5, (which, err)
    for k in range(deck.n_species):
        assert sum(len(s.download_particles(k)) for s in slabs) == len(single.part(k))
    assert all(s.n_move == single.sim.n_move for s in slabs)


def test_moving_window_chain_equals_single_domain():
    _check_against_single_loose(_window_deck, 3, 120, window=True, smooth=(2, 0, 4, 0))


# ------------------------------------------------------------------ two real processes over gloo

def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, steps, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ref = H.load_ref("em2d")
    deck = _weibel(ref)
    slab = _make_slabs(deck, world, False, ranks=[rank])[0]
    comm = P.TorchComm(slab.g)
    for _ in range(steps):
        P.slab_step(slab, comm)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), E=slab.download_grid(P.E), B=slab.download_grid(P.B),
             J=slab.download_grid(P.J), p0=slab.download_particles(0), p1=slab.download_particles(1))
    dist.barrier()
    dist.destroy_process_group()


def test_two_processes_over_gloo(tmp_path):
    ref = H.load_ref("em2d")
    if ref is None:
        pytest.skip("oracle/_ref not built")
    steps, world = 6, 2
    port = _free_port()
    mp.spawn(_gloo_worker, args=(world, port, steps, str(tmp_path)), nprocs=world, join=True)
    deck = _weibel(ref)
    single = O.OracleSim(deck, n_sort=0)
    single.iter(steps)
    res = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    scale = np.sqrt((single.E.astype(np.float64) ** 2).sum())
    for name, want in (("E", single.E), ("B", single.B), ("J", single.J)):
        got = P.join_grids([r[name] for r in res], world)
        want = want[1:-2, 1:-2, :]
        err = np.sqrt(((got.astype(np.float64) - want) ** 2).sum()) / (scale if name == "B" else np.sqrt((want.astype(np.float64) ** 2).sum()))
        assert err < 1e-6, (name, err)
    for k in range(2):
        assert sum(len(r["p%d" % k]) for r in res) == len(single.part(k))
