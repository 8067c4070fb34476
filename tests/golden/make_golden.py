"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref strict build).
Run in the build container (needs /root/reference -> `make -C oracle ref`):

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests import helpers as H  # noqa: E402

ref = H.load_ref("em2d")
assert ref is not None, "build oracle/_ref first"

# Weibel deck at 16x16 cells, 2x2 ppc: initial state and the state after 20 steps (one sort)
d = H.weibel(ref, n=16, ppc=(2, 2))
out = dict(nx=16, dx=d.sim.emf.dx[0], dt=d.sim.dt, steps=20,
           m_q=np.array([d.species[k].m_q for k in range(2)], dtype=np.float32),
           q=np.array([d.species[k].q for k in range(2)], dtype=np.float32),
           E0=d.E().copy(), B0=d.B().copy(), part0_s0=d.parts(0).copy(), part0_s1=d.parts(1).copy())
d.iter(20)
out.update(E1=d.E().copy(), B1=d.B().copy(), J1=d.J().copy(), part1_s0=d.parts(0).copy(), part1_s1=d.parts(1).copy(),
           energy=np.array([d.species[k].energy for k in range(2)]), emf_energy=d.emf_energy())
np.savez_compressed(os.path.join(HERE, "weibel_16x16.npz"), **out)
print("wrote weibel_16x16.npz", os.path.getsize(os.path.join(HERE, "weibel_16x16.npz")), "bytes")
