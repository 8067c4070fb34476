"""em1d counterpart of tests/helpers.py: one ctypes call sequence, two libraries."""
import ctypes as C
import math
import os

import numpy as np

from zpic_b200 import abi_em1d as A

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_ref(fast=False):
    path = os.path.join(REPO, "oracle", "_ref", "libzpic_ref_em1d%s.so" % ("_fast" if fast else ""))
    if not os.path.exists(path):
        return None
    return A.declare(C.CDLL(path, mode=C.RTLD_LOCAL))


def is_ours(lib):
    return hasattr(lib, "zpic_b200_sync_host")


class Deck1D:
    def __init__(self, lib, nx, box, dt, species=(), tmax=0.0, ndump=0, seed=(12345, 67890)):
        self.lib, self.nx = lib, nx
        lib.set_rand_seed(*seed)
        libc = C.CDLL(None)
        libc.calloc.restype = C.c_void_p
        libc.calloc.argtypes = [C.c_size_t, C.c_size_t]
        n = len(species)
        self._keep = []
        self.species = C.cast(libc.calloc(max(n, 1), C.sizeof(A.Species)), C.POINTER(A.Species)) if n else C.POINTER(A.Species)()
        for k, sp in enumerate(species):
            ufl = (C.c_float * 3)(*sp.get("ufl", (0, 0, 0)))
            uth = (C.c_float * 3)(*sp.get("uth", (0, 0, 0)))
            dens = None
            if "density" in sp:
                d = A.Density()
                for key, val in sp["density"].items():
                    if key == "ramp":
                        d.ramp[0], d.ramp[1] = val
                    else:
                        setattr(d, key, val)
                self._keep.append(d)
                dens = C.byref(d)
            lib.spec_new(C.byref(self.species[k]), sp["name"].encode(), sp["m_q"], sp["ppc"], ufl, uth, nx, box, dt, dens)
            if "n_sort" in sp:
                self.species[k].n_sort = sp["n_sort"]
            if "bc_type" in sp:
                self.species[k].bc_type = sp["bc_type"]
        self.sim = A.Simulation()
        lib.sim_new(C.byref(self.sim), nx, box, dt, tmax, ndump, self.species, n)
        self.n_species = n

    def add_laser(self, **kw):
        laser = A.Laser()
        for k, v in kw.items():
            setattr(laser, k, v)
        self.lib.sim_add_laser(C.byref(self.sim), C.byref(laser))

    def set_moving_window(self):
        self.lib.sim_set_moving_window(C.byref(self.sim))

    def set_smooth(self, xtype, xlevel):
        s = A.Smooth(xtype, xlevel)
        self.lib.sim_set_smooth(C.byref(self.sim), C.byref(s))

    def iter(self, n=1):
        for _ in range(n):
            self.lib.sim_iter(C.byref(self.sim))

    def sync(self):
        if is_ours(self.lib):
            self.lib.zpic_b200_sync_host(C.byref(self.sim))

    def E(self):
        return A.grid_view(self.sim.emf.E_buf, self.nx)

    def B(self):
        return A.grid_view(self.sim.emf.B_buf, self.nx)

    def J(self):
        return A.grid_view(self.sim.current.J_buf, self.nx)

    def parts(self, k):
        return A.part_view(self.species[k])

    def emf_energy(self):
        e = (C.c_double * 6)()
        self.lib.emf_get_energy(C.byref(self.sim.emf), e)
        return np.array(e[:])

    def charge(self, k):
        rho = np.zeros(self.nx + 1, dtype=np.float32)
        self.lib.spec_deposit_charge(C.byref(self.species[k]), rho.ctypes.data_as(C.POINTER(C.c_float)))
        return rho

    def snapshot(self):
        self.sync()
        return {"E": self.E().copy(), "B": self.B().copy(), "J": self.J().copy(),
                "np": [self.species[k].np for k in range(self.n_species)],
                "energy": [self.species[k].energy for k in range(self.n_species)],
                "parts": [self.parts(k).copy() for k in range(self.n_species)]}

    def delete(self):
        self.lib.sim_delete(C.byref(self.sim))


def twostream(lib, nx=120, ppc=500, n_sort=None, uth=(0.001, 0.001, 0.001)):
    """em1d/input/twostream.c as shipped (reference input/twostream.c:12-36)"""
    sp = []
    for name, u in (("right", 0.2), ("left", -0.2)):
        s = dict(name=name, m_q=-1.0, ppc=ppc, ufl=(u, 0.0, 0.0), uth=uth)
        if n_sort is not None:
            s["n_sort"] = n_sort
        sp.append(s)
    return Deck1D(lib, nx, np.float32(4 * math.pi) * nx / 120, 0.1, sp, tmax=50.0, ndump=10)


def absorbing(lib, nx=1000):
    """em1d/input/absorbing.c: laser in vacuum with Mur open boundaries"""
    d = Deck1D(lib, nx, 20.0, 0.019, [], tmax=40.0, ndump=50)
    d.add_laser(start=17.0, fwhm=2.0, a0=2.0, omega0=10.0, polarization=math.pi / 2)
    d.sim.emf.bc_type = A.EMF_BC_OPEN
    return d


def movwindow(lib, nx=512, ppc=32, n_sort=None):
    """em1d/input/movwindow.c pattern: STEP plasma entering a moving window, three lasers"""
    sp = dict(name="electrons", m_q=-1.0, ppc=ppc, density=dict(type=A.STEP, start=39.0))
    if n_sort is not None:
        sp["n_sort"] = n_sort
    d = Deck1D(lib, nx, 41.0, 0.07, [sp], tmax=300.0, ndump=50)
    d.add_laser(start=25.0, fwhm=7.0, a0=0.5, omega0=10.0, polarization=math.pi / 2)
    d.add_laser(start=25.0, fwhm=7.0, a0=0.05, omega0=11.0, polarization=math.pi / 2)
    d.set_moving_window()
    return d
