"""Shared test plumbing: the same ctypes call sequence ("deck") is run against the
product library and against the reference build in oracle/_ref, then raw buffers are
compared.  Nothing here is imported by the product."""
import ctypes as C
import os

import numpy as np

from zpic_b200 import abi_em2d as A

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_ref(code="em2d", fast=False):
    path = os.path.join(REPO, "oracle", "_ref", "libzpic_ref_%s%s.so" % (code, "_fast" if fast else ""))
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    if code == "em2d":
        A.declare(lib)
    return lib


def is_ours(lib):
    return hasattr(lib, "zpic_b200_sync_host")


class Deck:
    """A simulation built through the C API of `lib` (ours or the reference)."""

    def __init__(self, lib, nx, box, dt, species=(), tmax=0.0, ndump=0, seed=(12345, 67890)):
        self.lib = lib
        self.nx = tuple(nx)
        lib.set_rand_seed(*seed)
        self._nx = (C.c_int * 2)(*nx)
        self._box = (C.c_float * 2)(*box)
        n = len(species)
        # the reference frees sim->species with free(): allocate with the C allocator
        libc = C.CDLL(None)
        libc.calloc.restype = C.c_void_p
        libc.calloc.argtypes = [C.c_size_t, C.c_size_t]
        self._keep = []
        if n:
            raw = libc.calloc(n, C.sizeof(A.Species))
            self.species = C.cast(raw, C.POINTER(A.Species))
        else:
            self.species = C.POINTER(A.Species)()
        for k, sp in enumerate(species):
            ppc = (C.c_int * 2)(*sp["ppc"])
            ufl = (C.c_float * 3)(*sp.get("ufl", (0, 0, 0)))
            uth = (C.c_float * 3)(*sp.get("uth", (0, 0, 0)))
            dens = None
            if "density" in sp:
                dens = A.Density()
                for key, val in sp["density"].items():
                    setattr(dens, key, val)
                self._keep.append(dens)
                dens = C.byref(dens)
            lib.spec_new(C.byref(self.species[k]), sp["name"].encode(), sp["m_q"], ppc, ufl, uth,
                         self._nx, self._box, dt, dens)
            if "n_sort" in sp:
                self.species[k].n_sort = sp["n_sort"]
        self.sim = A.Simulation()
        lib.sim_new(C.byref(self.sim), self._nx, self._box, dt, tmax, ndump, self.species, n)
        self.n_species = n

    # --- configuration -------------------------------------------------
    def add_laser(self, **kw):
        laser = A.Laser()
        for k, v in kw.items():
            setattr(laser, k, v)
        self.lib.sim_add_laser(C.byref(self.sim), C.byref(laser))

    def set_moving_window(self):
        self.lib.sim_set_moving_window(C.byref(self.sim))

    def set_smooth(self, xtype=0, ytype=0, xlevel=0, ylevel=0):
        s = A.Smooth(xtype, ytype, xlevel, ylevel)
        self.lib.sim_set_smooth(C.byref(self.sim), C.byref(s))

    def set_ext_uniform(self, E0=None, B0=None):
        ext = A.ExtField()
        if E0 is not None:
            ext.E_type = A.EMF_FLD_TYPE_UNIFORM
            ext.E_0 = A.Float3(*E0)
        if B0 is not None:
            ext.B_type = A.EMF_FLD_TYPE_UNIFORM
            ext.B_0 = A.Float3(*B0)
        self.lib.sim_set_ext_fld(C.byref(self.sim), C.byref(ext))

    # --- stepping -------------------------------------------------------
    def iter(self, n=1):
        for _ in range(n):
            self.lib.sim_iter(C.byref(self.sim))

    def sync(self):
        if is_ours(self.lib):
            self.lib.zpic_b200_sync_host(C.byref(self.sim))

    def touch(self):
        if is_ours(self.lib):
            self.lib.zpic_b200_touch_host(C.byref(self.sim))

    # --- raw state ------------------------------------------------------
    def E(self):
        return A.grid_view(self.sim.emf.E_buf, *self.nx)

    def B(self):
        return A.grid_view(self.sim.emf.B_buf, *self.nx)

    def J(self):
        return A.grid_view(self.sim.current.J_buf, *self.nx)

    def parts(self, k):
        return A.part_view(self.species[k])

    def emf_energy(self):
        e = (C.c_double * 6)()
        self.lib.emf_get_energy(C.byref(self.sim.emf), e)
        return np.array(e[:])

    def charge(self, k):
        nx, ny = self.nx
        rho = np.zeros((ny + 1, nx + 1), dtype=np.float32)
        self.lib.spec_deposit_charge(C.byref(self.species[k]), rho.ctypes.data_as(C.POINTER(C.c_float)))
        return rho

    def snapshot(self):
        """copies of everything, after making the host mirrors current"""
        self.sync()
        out = {"E": self.E().copy(), "B": self.B().copy(), "J": self.J().copy(),
               "np": [self.species[k].np for k in range(self.n_species)],
               "energy": [self.species[k].energy for k in range(self.n_species)],
               "parts": [self.parts(k).copy() for k in range(self.n_species)]}
        return out

    def delete(self):
        self.lib.sim_delete(C.byref(self.sim))


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.sqrt((b * b).sum())
    num = np.sqrt(((a - b) ** 2).sum())
    return num / den if den > 0 else num


def canon(parts):
    """particles in a canonical order (for comparisons that must ignore buffer order)"""
    order = np.lexsort((parts["uz"], parts["uy"], parts["ux"], parts["y"], parts["x"], parts["iy"], parts["ix"]))
    return parts[order]


WEIBEL_SPECIES = (
    dict(name="electrons", m_q=-1.0, ppc=(2, 2), ufl=(0, 0, 0.6), uth=(0.1, 0.1, 0.1)),
    dict(name="positrons", m_q=+1.0, ppc=(2, 2), ufl=(0, 0, -0.6), uth=(0.1, 0.1, 0.1)),
)


def weibel(lib, n=128, ppc=(2, 2), n_sort=None, dt=0.07, cell=0.1):
    """em2d/input/weibel.c as shipped (reference input/weibel.c:13-40), size / ppc adjustable"""
    sp = []
    for s in WEIBEL_SPECIES:
        s = dict(s, ppc=ppc)
        if n_sort is not None:
            s["n_sort"] = n_sort
        sp.append(s)
    return Deck(lib, (n, n), (n * cell, n * cell), dt, sp, tmax=35.0, ndump=10)


def lwfa(lib, nx=(1500, 128), box=(30.0, 25.6), dt=0.014, ppc=(4, 2), start=None, n_sort=None,
         laser_start=27.0, a0=2.0):
    """em2d/input/lwfa.c as shipped (reference input/lwfa.c:15-63), geometry adjustable"""
    dens = dict(type=A.STEP, start=box[0] if start is None else start)
    sp = dict(name="electrons", m_q=-1.0, ppc=ppc, density=dens)
    if n_sort is not None:
        sp["n_sort"] = n_sort
    d = Deck(lib, nx, box, dt, [sp], tmax=40.6, ndump=10)
    d.add_laser(type=A.GAUSSIAN, start=laser_start, fwhm=2.0, a0=a0, omega0=10.0, W0=4.0,
                focus=20.0, axis=box[1] / 2, polarization=np.pi / 2)
    d.set_moving_window()
    d.set_smooth(xtype=A.COMPENSATED, xlevel=4)
    return d
