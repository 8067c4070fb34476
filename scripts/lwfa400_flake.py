"""Repeat the comparison of tests/test_gpu_em2d.py::test_lwfa_shipped_deck_400_steps and print the numbers its assertions
look at (the J summation order of the GPU run differs from run to run, so they vary a little): usage python scripts/lwfa400_flake.py [reps]"""
import sys, os
import numpy as np
sys.path.insert(0, os.getcwd())
from tests import helpers as H
from tests.test_gpu_em2d import _match_in_cells
from zpic_b200 import load
ours = load("em2d"); assert ours.zdev_init(-1) == 0

ref = H.load_ref("em2d")
b = H.lwfa(ref, n_sort=0); b.iter(400); sb = b.snapshot(); eb = b.emf_energy()
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    a = H.lwfa(ours, n_sort=0); a.iter(400); sa = a.snapshot(); ea = a.emf_energy()
    pa, pb = _match_in_cells(sa["parts"][0], sb["parts"][0])
    same = (pa["ix"] == pb["ix"]) & (pa["iy"] == pb["iy"])
    d2 = sum((pa[q].astype(np.float64) - pb[q]) ** 2 for q in ("ux", "uy", "uz"))
    n2 = sum(pb[q].astype(np.float64) ** 2 for q in ("ux", "uy", "uz"))
    bad = d2 > 1e-6 * max(n2.max(), 1e-30)
    err = np.sqrt(d2[~bad].sum() / max(n2[~bad].sum(), 1e-300))
    tot_a, tot_b = ea.sum() + sa["energy"][0], eb.sum() + sb["energy"][0]
    print("rep %d: np %d/%d cells differ %d, bad momenta %d, u err %.2e, E %.2e B %.2e J %.2e, energy rel %.2e" % (
        rep, sa["np"][0], sb["np"][0], (~same).sum(), bad.sum(), err, H.rel_l2(sa["E"], sb["E"]), H.rel_l2(sa["B"], sb["B"]),
        H.rel_l2(sa["J"], sb["J"]), abs(tot_a - tot_b) / abs(tot_b)))
    a.delete()
