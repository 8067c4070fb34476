"""Quick device-only throughput probe of the em1d path (not the bench): two-stream plasma, N steps.
usage: python scripts/quick_push_probe1d.py [log2 cells] [ppc] [steps]"""
import ctypes as C
import sys

import numpy as np

from zpic_b200 import load
from zpic_b200._lib import PushParams1D

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ppc = int(sys.argv[2]) if len(sys.argv) > 2 else 256
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
n = 1 << lg
lib = load("em1d")
assert lib.zdev_init(-1) == 0
g = lib.zdev_grid1d_create(n)
dx = np.float32(4 * np.pi / 120)          # em1d/input/twostream.c:15-20
dt = np.float32(0.1)
specs = []
for k, sign in enumerate((1.0, -1.0)):
    s = lib.zdev_spec1d_create(n, ppc, 0)
    ufl = (C.c_float * 3)(0.2 * sign, 0, 0)
    uth = (C.c_float * 3)(0.001, 0.001, 0.001)
    lib.zdev_spec1d_inject_uniform(s, ppc, ufl, uth, 4321 + k)
    q = np.float32(-1.0) / np.float32(ppc)
    prm = PushParams1D(float(np.float32(0.5 * float(dt) / -1.0)), float(dt / dx), float(q * dx / dt), float(q), 0, 0)
    specs.append((s, prm))
lib.zdev_sync()
npart = 2 * n * ppc
print("em1d grid 2^%d cells, ppc %d: %d particles" % (lg, ppc, npart))


def step():
    lib.zdev_current1d_zero(g)
    for s, prm in specs:
        lib.zdev_spec1d_advance(s, g, g, C.byref(prm))
    lib.zdev_current1d_update(g, 1, 0, 0)
    lib.zdev_emf1d_advance(g, g, float(dt), float(dx), 0, 0)


for _ in range(3):
    step()
e0, e1 = lib.zdev_event_create(), lib.zdev_event_create()
lib.zdev_event_record(e0)
for _ in range(steps):
    step()
lib.zdev_event_record(e1)
ms = lib.zdev_event_elapsed_ms(e0, e1)
en, npn = C.c_double(), C.c_int64()
lib.zdev_spec1d_fetch(specs[0][0], C.byref(en), C.byref(npn))
print("%.3f ms/step, %.2f Gpush/s, %.1f GB/s at 40 B/push; np[0]=%d energy_sum=%g" %
      (ms / steps, npart * steps / ms / 1e6, 40 * npart * steps / ms / 1e6, npn.value, en.value))
