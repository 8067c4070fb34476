"""The reference's own Cython module (python/source/em2d.pyx, unmodified) compiled against include/em2d and
linked to libzpic_b200_em2d.so (zpic_b200/cython/build_modules.py): the drop-in at the Python boundary.
Skipped when the module was not built (it needs /root/reference at build time; the built .so travels)."""
import ctypes as C
import glob
import importlib.util
import os

import numpy as np
import pytest

from tests import helpers as H

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(code):
    hits = glob.glob(os.path.join(REPO, "zpic_b200", "cython", "_build", code + ".*.so"))
    if not hits:
        pytest.skip("Cython module not built (python -m zpic_b200.cython.build_modules)")
    spec = importlib.util.spec_from_file_location(code, hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _weibel(em2d, n=32, ppc=(2, 2)):
    sp = [em2d.Species("electrons", -1.0, list(ppc), ufl=[0.0, 0.0, 0.6], uth=[0.1, 0.1, 0.1]),
          em2d.Species("positrons", +1.0, list(ppc), ufl=[0.0, 0.0, -0.6], uth=[0.1, 0.1, 0.1])]
    return em2d.Simulation([n, n], [n * 0.1, n * 0.1], 0.07, species=sp), sp


def test_module_builds_a_simulation_on_the_host(ours):
    """no GPU needed: Species / Simulation construction runs the library's host layer (reference random stream)"""
    em2d = _load("em2d")
    sim, sp = _weibel(em2d)
    deck = H.weibel(ours, n=32, ppc=(2, 2))
    for k in range(2):
        a, b = np.asarray(sp[k].particles), deck.parts(k)
        assert a.shape == b.shape == (32 * 32 * 4,)
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
    assert sim.emf.Ex.shape == (32, 32) and sim.n == 0


@pytest.mark.gpu
@pytest.mark.parametrize("coherent", [0, 1])
def test_module_steps_like_the_reference(ours, ref, coherent):
    """the unmodified module on the CUDA library against the reference build driven through ctypes: same deck, 10
    steps.  coherent = 0: the module's numpy views are guarded mirrors (zb_guard.h) - what it reads is downloaded
    when it reads it; coherent = 1: ZPIC_COHERENT semantics, every mirror refreshed around every sim_iter"""
    em2d = _load("em2d")
    assert ours.zdev_init(-1) == 0
    ours.zpic_b200_set_option(b"coherent", coherent)
    ours.zpic_b200_set_option(b"lazy", 0)
    try:
        sim, sp = _weibel(em2d)
        b = H.weibel(ref, n=32, ppc=(2, 2), n_sort=0)
        for _ in range(10):
            sim.iter()
        b.iter(10)
        assert sim.n == 10
        for name, comp, grid in (("Ez", 2, b.E()), ("Bx", 0, b.B()), ("By", 1, b.B())):
            got = np.asarray(getattr(sim.emf, name))
            want = grid[1:-2, 1:-2, comp]
            assert np.abs(got - want).max() <= 1e-5 * max(np.abs(want).max(), 0.6 * 10 * 0.07), name
        for k in range(2):
            pa = H.canon(np.asarray(sp[k].particles).copy())
            pb = H.canon(b.parts(k).copy())
            assert len(pa) == len(pb)
            assert (pa["ix"] != pb["ix"]).sum() + (pa["iy"] != pb["iy"]).sum() <= 2
    finally:
        ours.zpic_b200_set_option(b"coherent", 0)
