"""One rank of a slab-decomposed run, driven through the reference's C API (tests only).
usage (per rank; ZPIC_RANK / ZPIC_NRANKS / ZPIC_JOB in the environment):
    python tests/slab_worker.py DECK CHECKPOINTS OUT.npz
Every rank builds the same deck and makes the same calls (SPMD); after each checkpoint the host mirrors are
synchronised (a gather over the ranks) and rank 0 stores them."""
import ctypes as C
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from tests import helpers as H  # noqa: E402


def build(lib, name):
    if name == "weibel":
        return H.weibel(lib, n=64, ppc=(2, 2), n_sort=0)      # as test_weibel_cells_cross_after_a_few_steps
    if name == "weibel_smooth":
        d = H.weibel(lib, n=64, ppc=(2, 2), n_sort=0)
        d.set_smooth(xtype=1, ytype=2, xlevel=2, ylevel=2)
        return d
    if name == "lwfa":
        return H.lwfa(lib, nx=(256, 64), box=(5.12, 12.8), dt=0.014, ppc=(2, 2), start=4.0, laser_start=3.5, a0=1.0, n_sort=0)
    if name == "kh":
        from tests.test_host_init import kh_deck
        return kh_deck(lib)
    raise SystemExit("unknown deck " + name)


def run(lib, name, checkpoints, rank0=True):
    d = build(lib, name)
    out = {}
    for cp in checkpoints:
        d.iter(cp - d.sim.emf.iter)
        np_step = [d.species[k].np for k in range(d.n_species)]          # as the step reports it (no sync)
        s = d.snapshot()
        en = d.emf_energy()
        rho = [d.charge(k) for k in range(d.n_species)]
        if rank0:
            for q in ("E", "B", "J"):
                out["%s_%d" % (q, cp)] = s[q]
            out["np_%d" % cp] = np.array(s["np"])
            out["np_step_%d" % cp] = np.array(np_step)
            out["energy_%d" % cp] = np.array(s["energy"])
            out["emf_energy_%d" % cp] = en
            out["n_move_%d" % cp] = np.array([d.sim.emf.n_move])
            for k in range(d.n_species):
                out["parts%d_%d" % (k, cp)] = H.canon(s["parts"][k])
                out["rho%d_%d" % (k, cp)] = rho[k]
    d.delete()
    return out


if __name__ == "__main__":
    from zpic_b200 import load
    name, cps, path = sys.argv[1], [int(x) for x in sys.argv[2].split(",")], sys.argv[3]
    lib = load("em2d")
    assert lib.zdev_init(-1) == 0
    lib.zpic_b200_set_option(b"lazy", int(os.environ.get("ZPIC_TEST_LAZY", "0")))
    rank = int(os.environ.get("ZPIC_RANK", os.environ.get("RANK", "0")))
    res = run(lib, name, cps, rank0=(rank == 0))
    if rank == 0:
        np.savez(path, **res)
