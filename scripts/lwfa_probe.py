"""LWFA-type step timing on one GPU (config 3 physics at reduced size): laser, moving window, STEP plasma
filling the box, compensated smoothing level 4, 16 ppc.  Prints ms/step and pushes/s."""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from tests import helpers as H
from zpic_b200 import load

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ny = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 50
lib = load("em2d")
assert lib.zdev_init(-1) == 0
lib.zpic_b200_set_option(b"lazy", 1)
box = (nx * 0.01, ny * 0.05)
t0 = time.time()
d = H.lwfa(lib, nx=(nx, ny), box=box, dt=0.009, ppc=(4, 4), start=0.5, laser_start=box[0] - 3.0, a0=3.0)
print("host init %.1f s, np = %d" % (time.time() - t0, d.species[0].np))
d.iter(5)
lib.zdev_sync()
e0, e1 = lib.zdev_event_create(), lib.zdev_event_create()
lib.zdev_event_record(e0)
d.iter(steps)
lib.zdev_event_record(e1)
ms = lib.zdev_event_elapsed_ms(e0, e1)
lib.zpic_b200_set_option(b"lazy", 0)
npart = d.species[0].np
print("%.3f ms/step, %.2f Gpush/s, %.2f Gcell/s (n_move %d)" % (ms / steps, npart * steps / ms / 1e6, nx * ny * steps / ms / 1e6, d.sim.emf.n_move))
