"""Slab decomposition behind the reference's C API: N processes (one per slab; here all on the GPU of the test
box - the links are CUDA IPC mappings either way) build the same deck through sim_new / sim_iter, exchange guard
cells and particles GPU to GPU (csrc/dev/zdev_slab.cuh), and rank 0's gathered mirrors are compared with the
unmodified reference running the whole box."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests import helpers as H
from tests import slab_worker as W

pytestmark = pytest.mark.gpu
REPO = H.REPO
TOL = 1e-5


def launch(name, cps, nranks, tmp_path, lazy=0, timeout=240):
    out = str(tmp_path / ("%s_%d.npz" % (name, nranks)))
    procs = []
    for r in range(nranks):
        env = dict(os.environ, ZPIC_RANK=str(r), ZPIC_NRANKS=str(nranks), ZPIC_JOB="t%d_%s" % (os.getpid(), name),
                   ZPIC_TEST_LAZY=str(lazy), PYTHONPATH=REPO, ZPIC_DEVICE=str(r))      # one GPU each where the box has several
        env.pop("RANK", None)
        env.pop("WORLD_SIZE", None)
        procs.append(subprocess.Popen([sys.executable, os.path.join(REPO, "tests", "slab_worker.py"), name,
                                       ",".join(map(str, cps)), out], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    logs = []
    for p in procs:
        try:
            logs.append(p.communicate(timeout=timeout)[0].decode()[-2000:])
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    return dict(np.load(out))


def compare(got, want, cps, nsp, nx, nranks, bit_exact_interior=False):
    nxl = nx // nranks
    for cp in cps:
        assert got["n_move_%d" % cp][0] == want["n_move_%d" % cp][0]
        assert np.array_equal(got["np_%d" % cp], want["np_%d" % cp]), cp
        assert np.array_equal(got["np_step_%d" % cp], want["np_step_%d" % cp]), cp
        # E and B are measured on the scale of the electromagnetic field as a whole (same units): after one step of
        # a deck that starts from zero fields B is ~1e-3 of E and pure summation-order noise relative to itself
        em = max(np.sqrt((want["E_%d" % cp].astype(np.float64) ** 2).sum()), np.sqrt((want["B_%d" % cp].astype(np.float64) ** 2).sum()))
        for q in ("E", "B", "J"):
            a, b = got["%s_%d" % (q, cp)].astype(np.float64), want["%s_%d" % (q, cp)].astype(np.float64)
            den = em if q != "J" else np.sqrt((b ** 2).sum())
            err = np.sqrt(((a - b) ** 2).sum()) / max(den, 1e-300)
            assert err < TOL, (cp, q, err)
        for k in range(nsp):
            pa, pb = got["parts%d_%d" % (k, cp)], want["parts%d_%d" % (k, cp)]
            same = (pa["ix"] == pb["ix"]) & (pa["iy"] == pb["iy"])
            assert (~same).sum() <= (0 if cp == 1 else 3), (cp, k, (~same).sum())
            if cp == 1:
                assert np.array_equal(pa.view(np.uint8), pb.view(np.uint8))      # one step: bit-exact particles
            ea, eb = got["energy_%d" % cp][k], want["energy_%d" % cp][k]
            assert abs(ea - eb) <= 1e-6 * abs(eb), (cp, k)
            ra, rb = got["rho%d_%d" % (k, cp)], want["rho%d_%d" % (k, cp)]
            assert np.abs(ra - rb).max() <= 2e-5 * max(np.abs(rb).max(), 1e-30), (cp, k)
        fa, fb = got["emf_energy_%d" % cp].sum(), want["emf_energy_%d" % cp].sum()
        assert abs(fa - fb) <= 5e-6 * max(fb, 1e-30), cp


@pytest.mark.parametrize("name,cps,nsp,nx,nranks", [
    ("weibel", (1, 12), 2, 64, 2),
    ("weibel", (1, 12), 2, 64, 4),
    ("weibel_smooth", (1, 30), 2, 64, 2),
    ("kh", (1, 40), 2, 64, 2),
    ("lwfa", (1, 40, 120), 1, 256, 2),
    ("lwfa", (1, 40, 120), 1, 256, 4),
])
def test_slabs_through_the_c_api_match_the_reference(ref, tmp_path, name, cps, nsp, nx, nranks):
    got = launch(name, cps, nranks, tmp_path)
    want = W.run(ref, name, cps)
    compare(got, want, cps, nsp, nx, nranks)


def test_slabs_lazy_mode_never_waits_and_still_matches(ref, tmp_path):
    """ZPIC_LAZY: no per-step fetch, the control block is only looked at one step later"""
    got = launch("weibel", (1, 12), 2, tmp_path, lazy=1)
    want = W.run(ref, "weibel", (1, 12))
    for cp in (1, 12):
        for q in ("E", "B", "J"):
            assert H.rel_l2(got["%s_%d" % (q, cp)], want["%s_%d" % (q, cp)]) < TOL
        assert np.array_equal(got["np_%d" % cp], want["np_%d" % cp])
