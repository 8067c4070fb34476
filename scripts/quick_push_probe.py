"""Quick device-only throughput probe (not the bench): uniform thermal plasma, N steps."""
import ctypes as C
import sys
import time

import numpy as np

from zpic_b200 import load
from zpic_b200._lib import PushParams2D

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ppc_arg = sys.argv[2] if len(sys.argv) > 2 else "8"               # "8" = 8x8 per cell, "4x8" = 4 along x, 8 along y
ppcx, ppcy = (int(v) for v in ppc_arg.split("x")) if "x" in ppc_arg else (int(ppc_arg), int(ppc_arg))
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
lib = load("em2d")
assert lib.zdev_init(-1) == 0
g = lib.zdev_grid2d_create(n, n)
dx = np.float32(0.1)
dt = np.float32(0.07)
specs = []
for k, sign in enumerate((-1.0, 1.0)):
    s = lib.zdev_spec2d_create(n, n, ppcx * ppcy, 0)
    ufl = (C.c_float * 3)(0, 0, 0.6 * sign)
    uth = (C.c_float * 3)(0.1, 0.1, 0.1)
    lib.zdev_spec2d_inject_uniform(s, ppcx, ppcy, ufl, uth, 1234 + k)
    q = np.float32(sign) / np.float32(ppcx * ppcy)
    prm = PushParams2D(float(np.float32(0.5 * float(dt) / sign)), float(dt / dx), float(dt / dx),
                       float(q * dx / dt), float(q * dx / dt), float(q), 0, 0)
    specs.append((s, prm))
lib.zdev_sync()
tx, ty, nt, cap = C.c_int(), C.c_int(), C.c_int(), C.c_int64()
lib.zdev_spec2d_tile_info(specs[0][0], C.byref(tx), C.byref(ty), C.byref(nt), C.byref(cap))
npart = 2 * n * n * ppcx * ppcy
print("grid %d^2 ppc %d: %d particles, tile %dx%d, %d tiles, capacity %d" % (n, ppcx * ppcy, npart, tx.value, ty.value, nt.value, cap.value))


def step():
    lib.zdev_current_zero(g)
    for s, prm in specs:
        lib.zdev_spec2d_advance(s, g, g, C.byref(prm))
    lib.zdev_current_update(g, 0, 0, 0, 0, 0)
    lib.zdev_emf_advance(g, g, float(dt), float(dx), float(dx), 0, 0)


for _ in range(3):
    step()
e0, e1 = lib.zdev_event_create(), lib.zdev_event_create()
lib.zdev_event_record(e0)
for _ in range(steps):
    step()
lib.zdev_event_record(e1)
ms = lib.zdev_event_elapsed_ms(e0, e1)
en = C.c_double()
npn = C.c_int64()
lib.zdev_spec2d_fetch(specs[0][0], C.byref(en), C.byref(npn))
print("%.3f ms/step, %.2f Gpush/s, %.1f GB/s at 56 B/push; np[0]=%d energy_sum=%g" %
      (ms / steps, npart * steps / ms / 1e6, 56 * npart * steps / ms / 1e6, npn.value, en.value))
sums = (C.c_double * 6)()
lib.zdev_emf_energy(g, sums)
print("field sums", list(sums))
