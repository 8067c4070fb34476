import ctypes as C, time, sys, os
import numpy as np
sys.path.insert(0, os.getcwd())
import bench
from zpic_b200 import abi_em2d as A, load
lib = load("em2d"); assert lib.zdev_init(0) == 0
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
os.environ.setdefault("ZPIC_TILE_SLACK", "1.25")
lib.zpic_b200_set_option(b"device_init", 2 if n > 1024 else 0); lib.zpic_b200_set_option(b"lazy", 0); lib.zpic_b200_set_option(b"coherent", 0)
t0 = time.perf_counter()
sim, species, _ = bench.build_weibel(lib, A, n, n, (8, 8))
print("host init %.1f s" % (time.perf_counter() - t0))
rho = [np.zeros((n + 1, n + 1), dtype=np.float32) for _ in range(2)]
en6 = (C.c_double * 6)()
def T(f, reps=5):
    lib.zdev_sync(); t0 = time.perf_counter()
    for _ in range(reps): f()
    lib.zdev_sync(); return (time.perf_counter() - t0) / reps * 1e3
for _ in range(3): lib.sim_iter(C.byref(sim))
print("sim_iter (non-lazy)      %.2f ms" % T(lambda: lib.sim_iter(C.byref(sim))))
print("emf_get_energy           %.2f ms" % T(lambda: lib.emf_get_energy(C.byref(sim.emf), en6)))
def f():
    lib.sim_iter(C.byref(sim)); lib.zpic_b200_sync_emf(C.byref(sim.emf))
print("sim_iter + sync_emf      %.2f ms" % T(f))
def f():
    lib.sim_iter(C.byref(sim)); lib.zpic_b200_sync_current(C.byref(sim.current))
print("sim_iter + sync_current  %.2f ms" % T(f))
def f():
    for s in range(2):
        rho[s][...] = 0
        lib.spec_deposit_charge(C.byref(species[s]), rho[s].ctypes.data_as(C.POINTER(C.c_float)))
print("2 x spec_deposit_charge  %.2f ms" % T(f))
lib.zpic_b200_set_option(b"lazy", 1)
print("sim_iter (lazy)          %.2f ms" % T(lambda: lib.sim_iter(C.byref(sim)), 10))
lib.zpic_b200_set_option(b"lazy", 0)
for rep in range(6):
    lib.zdev_sync(); t0 = time.perf_counter()
    lib.spec_deposit_charge(C.byref(species[rep % 2]), rho[rep % 2].ctypes.data_as(C.POINTER(C.c_float)))
    print("deposit_charge call %d: %.2f ms" % (rep, (time.perf_counter() - t0) * 1e3))
for rep in range(4):
    lib.sim_iter(C.byref(sim)); lib.zdev_sync(); t0 = time.perf_counter()
    lib.zpic_b200_sync_emf(C.byref(sim.emf))
    t1 = time.perf_counter()
    lib.zpic_b200_sync_current(C.byref(sim.current))
    print("sync_emf %.2f ms, sync_current %.2f ms" % ((t1 - t0) * 1e3, (time.perf_counter() - t1) * 1e3))
