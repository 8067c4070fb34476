"""Device-only probe of the grid half of an em2d step (not the bench): current_zero + current_update +
emf_advance on an n x n grid, with the one-pass field kernel and with the three separate stencils.
usage: python scripts/grid_probe.py [n] [reps]"""
import sys

from zpic_b200 import load

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
lib = load("em2d")
assert lib.zdev_init(-1) == 0
g = lib.zdev_grid2d_create(n, n)
dt, dx = 0.07, 0.1


def grid_half():
    lib.zdev_current_zero(g)
    lib.zdev_current_update(g, 0, 0, 0, 0, 0)
    lib.zdev_emf_advance(g, g, dt, dx, dx, 0, 0)


for fused in (1, 0, 2, 1, 0, 2):      # 1: tile kernel (shared memory), 0: three stencils, 2: rows marched in registers
    lib.zdev_yee_set_fused(fused)
    for _ in range(5):
        grid_half()
    e0, e1 = lib.zdev_event_create(), lib.zdev_event_create()
    lib.zdev_sync()
    lib.zdev_event_record(e0)
    for _ in range(reps):
        grid_half()
    lib.zdev_event_record(e1)
    ms = lib.zdev_event_elapsed_ms(e0, e1) / reps
    print("fused=%d  %.4f ms per grid half, %.2f G cell-updates/s, %.0f GB/s at 72 B/cell" %
          (fused, ms, n * n / ms / 1e6, 72.0 * n * n / ms / 1e6))
lib.zdev_grid2d_destroy(g)
