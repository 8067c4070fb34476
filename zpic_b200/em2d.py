"""em2d - the reference's Python API (python/source/em2d.pyx) on top of the CUDA library.

Same classes, constructor arguments, properties and methods as the Cython module of the
reference, so notebooks written for `import em2d` run with `from zpic_b200 import em2d`:

    sim = em2d.Simulation(nx=[128,128], box=[12.8,12.8], dt=0.07, species=[electrons, positrons])
    sim.run(35.0)
    plt.imshow(sim.emf.Bx)

Every array property returns a numpy view of the HOST mirror of the corresponding C
buffer (em2d.pyx:305-312, 1044-1296, 1567-1619), as in the reference.  Because the device
copy is authoritative while stepping, each getter first refreshes the mirror and marks it
as "possibly modified by the caller", so in-place edits such as
`species.particles['ux'] += ...` reach the device before the next iteration (SURVEY.md 8b,
coherence contract).
"""
import ctypes as C

import numpy as np

from . import abi_em2d as A
from ._lib import load

_lib = None


def _L():
    global _lib
    if _lib is None:
        _lib = load("em2d")
    return _lib


_libc = C.CDLL(None)
_libc.calloc.restype = C.c_void_p
_libc.calloc.argtypes = [C.c_size_t, C.c_size_t]
_libc.free.argtypes = [C.c_void_p]


class Density:
    """Density profile (em2d.pyx:25-170)"""
    _types = {"uniform": A.UNIFORM, "empty": A.EMPTY, "step": A.STEP, "slab": A.SLAB, "custom": A.CUSTOM}

    def __init__(self, *, type="uniform", start=0.0, end=0.0, n=1.0, custom_x=None, custom_y=None):
        self._c = A.Density()
        self._c.type = self._types[type]
        self._c.n = n
        self._c.start = start
        self._c.end = end
        self._type = type
        self._fx = self._fy = None
        self.custom_x, self.custom_y = custom_x, custom_y
        if custom_x:
            self._fx = A.DENSITY_FN(lambda x, data: float(custom_x(x)))
            self._c.custom_x = self._fx
        if custom_y:
            self._fy = A.DENSITY_FN(lambda y, data: float(custom_y(y)))
            self._c.custom_y = self._fy

    def copy(self):
        return Density(type=self._type, start=self._c.start, end=self._c.end, n=self._c.n,
                       custom_x=self.custom_x, custom_y=self.custom_y)

    n = property(lambda s: s._c.n, lambda s, v: setattr(s._c, "n", v))
    type = property(lambda s: s._type)
    start = property(lambda s: s._c.start, lambda s, v: setattr(s._c, "start", v))
    end = property(lambda s: s._c.end, lambda s, v: setattr(s._c, "end", v))


class Species:
    """Particle species (em2d.pyx:171-480)"""
    _diag_types = {"charge": 0x1000, "pha": 0x2000, "particles": 0x3000}
    _pha_quants = {"x1": 1, "x2": 2, "u1": 4, "u2": 5, "u3": 6}

    def __init__(self, name, m_q, ppc=(1, 1), *, ufl=(0., 0., 0.), uth=(0., 0., 0.), density=None, n_sort=16):
        self._name = name
        self._m_q = m_q
        self._ppc = list(ppc)
        self._ufl = list(ufl)
        self._uth = list(uth)
        self._n_sort = n_sort
        self._density = density.copy() if density else Density()
        self._p = None          # POINTER(Species) once attached to a simulation

    def _new(self, ptr, nx, box, dt):
        self._p = ptr
        _L().spec_new(ptr, self._name.encode(), self._m_q, (C.c_int * 2)(*self._ppc),
                      (C.c_float * 3)(*self._ufl), (C.c_float * 3)(*self._uth), nx, box, dt,
                      C.byref(self._density._c))
        ptr.contents.n_sort = self._n_sort

    @property
    def _s(self):
        return self._p.contents

    def add(self, ix, x, u):
        """append one particle (em2d.pyx:238-265)"""
        L = _L()
        L.zpic_b200_touch_species(self._p)
        s = self._s
        L.spec_grow_buffer(self._p, s.np + 1)
        s.part[s.np] = A.Part(int(ix[0]), int(ix[1]), x[0], x[1], u[0], u[1], u[2])
        s.np = s.np + 1

    def report(self, type, *, quants=(), pha_nx=(), pha_range=()):
        """save diagnostics to a ZDF file (em2d.pyx:268-303)"""
        rep = self._diag_types[type]
        if type == "pha":
            rep += self._pha_quants[quants[0]] + 16 * self._pha_quants[quants[1]]
            nxa = (C.c_int * 2)(*pha_nx)
            rng = ((C.c_float * 2) * 2)((C.c_float * 2)(*pha_range[0]), (C.c_float * 2)(*pha_range[1]))
            _L().spec_report(self._p, rep, nxa, rng)
        else:
            _L().spec_report(self._p, rep, None, None)

    @property
    def particles(self):
        """structured ndarray view of the particle buffer, read/write (em2d.pyx:305-312)"""
        _L().zpic_b200_touch_species(self._p)
        return A.part_view(self._s)

    def charge(self):
        """charge density of the species, shape (ny, nx) (em2d.pyx:314-331)"""
        s = self._s
        rho = np.zeros((s.nx[1] + 1, s.nx[0] + 1), dtype=np.float32)
        _L().spec_deposit_charge(self._p, rho.ctypes.data_as(C.POINTER(C.c_float)))
        return rho[0:s.nx[1], 0:s.nx[0]]

    def phasespace(self, quants, pha_nx, pha_range):
        """phasespace density of the species (em2d.pyx:333-368)"""
        rep = self._pha_quants[quants[0]] + 16 * self._pha_quants[quants[1]] + 0x2000
        pha = np.zeros((pha_nx[1], pha_nx[0]), dtype=np.float32)
        nxa = (C.c_int * 2)(*pha_nx)
        rng = ((C.c_float * 2) * 2)((C.c_float * 2)(*pha_range[0]), (C.c_float * 2)(*pha_range[1]))
        fn = _L().spec_deposit_pha
        fn.restype = None
        fn.argtypes = [C.POINTER(A.Species), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        fn(self._p, rep, nxa, rng, pha.ctypes.data)
        return pha

    dx = property(lambda s: np.array(s._s.dx[:], dtype=np.float32))
    dt = property(lambda s: s._s.dt)
    iter = property(lambda s: s._s.iter)
    ppc = property(lambda s: np.array(s._s.ppc[:], dtype=np.int32))
    n_move = property(lambda s: s._s.n_move)
    name = property(lambda s: s._name)
    energy = property(lambda s: s._s.energy)

    @property
    def n_sort(self):
        return self._s.n_sort if self._p else self._n_sort

    @n_sort.setter
    def n_sort(self, value):
        if value < 0:
            raise ValueError("n_sort must be >= 0")
        self._n_sort = value
        if self._p:
            self._s.n_sort = value


def _fld_type(name):
    return {"none": A.EMF_FLD_TYPE_NONE, "uniform": A.EMF_FLD_TYPE_UNIFORM, "custom": A.EMF_FLD_TYPE_CUSTOM}[name]


class _FieldTable(C.Structure):          # include/zpic_b200.h zpic_b200_field_table
    _fields_ = [("nrow", C.c_int), ("table", C.POINTER(C.c_float))]


class _FieldSpec:
    """common part of ExternalField / InitialField (em2d.pyx:481-878).  Custom fields are Python callables
    f(ix, dx, iy, dy) -> (fx, fy, fz) like in the reference; ctypes cannot return a struct from a callback, so they
    are evaluated once for every cell of the buffer (the reference calls them with exactly these arguments,
    em2d/emf.c:862-863, 924-980) and handed to the C side as a table read by zpic_b200_table_field."""
    _struct = None

    def __init__(self, *, E_type="none", B_type="none", E_0=(0., 0., 0.), B_0=(0., 0., 0.),
                 E_custom=None, B_custom=None):
        self._c = self._struct()
        self._c.E_type = _fld_type(E_type)
        self._c.B_type = _fld_type(B_type)
        self._c.E_0 = A.Float3(*E_0)
        self._c.B_0 = A.Float3(*B_0)
        self._fns = {"E": E_custom, "B": B_custom}
        self._keep = []

    def _bind(self, nx, ny, dx, dy):
        """evaluate the custom callables on the grid of the simulation this object is being attached to"""
        self._keep = []
        tramp = C.cast(_L().zpic_b200_table_field, A.FIELD_FN)
        for f, fn in self._fns.items():
            if not fn or getattr(self._c, f + "_type") != A.EMF_FLD_TYPE_CUSTOM:
                continue
            tab = np.empty((ny + 3, nx + 3, 3), dtype=np.float32)
            for j in range(-1, ny + 2):
                for i in range(-1, nx + 2):
                    tab[j + 1, i + 1] = fn(i, dx, j, dy)
            desc = _FieldTable(nx + 3, tab.ctypes.data_as(C.POINTER(C.c_float)))
            self._keep += [tab, desc]
            setattr(self._c, f + "_custom", tramp)
            setattr(self._c, f + "_custom_data", C.cast(C.pointer(desc), C.c_void_p))


class ExternalField(_FieldSpec):
    _struct = A.ExtField


class InitialField(_FieldSpec):
    _struct = A.InitField


class EMF:
    """EM fields of a simulation (em2d.pyx:879-1296)"""

    def __init__(self, ptr):
        self._p = ptr

    @property
    def _e(self):
        return self._p.contents

    def report(self, type, fc):
        _L().emf_report(self._p, bytes([{"E": A.EFLD, "B": A.BFLD, "Epart": A.EPART, "Bpart": A.BPART}[type]]), fc)

    def get_energy(self):
        e = (C.c_double * 6)()
        _L().emf_get_energy(self._p, e)
        return np.array(e[:])

    def init_fld(self, init_fld):
        e = self._e
        init_fld._bind(e.nx[0], e.nx[1], e.dx[0], e.dx[1])
        _L().emf_init_fld(self._p, C.byref(init_fld._c))

    def set_ext_fld(self, ext_fld):
        self._ext = ext_fld
        e = self._e
        ext_fld._bind(e.nx[0], e.nx[1], e.dx[0], e.dx[1])
        _L().emf_set_ext_fld(self._p, C.byref(ext_fld._c))

    nx = property(lambda s: np.array(s._e.nx[:], dtype=np.int32))
    dx = property(lambda s: np.array(s._e.dx[:], dtype=np.float32))
    box = property(lambda s: np.array(s._e.box[:], dtype=np.float32))
    n_move = property(lambda s: s._e.n_move)

    def _view(self, buf, comp):
        _L().zpic_b200_touch_emf(self._p)
        nx, ny = self._e.nx[0], self._e.nx[1]
        g = A.grid_view(buf, nx, ny)
        return g[1:ny + 1, 1:nx + 1, comp]

    Ex = property(lambda s: s._view(s._e.E_buf, 0))
    Ey = property(lambda s: s._view(s._e.E_buf, 1))
    Ez = property(lambda s: s._view(s._e.E_buf, 2))
    Bx = property(lambda s: s._view(s._e.B_buf, 0))
    By = property(lambda s: s._view(s._e.B_buf, 1))
    Bz = property(lambda s: s._view(s._e.B_buf, 2))

    def _part_view(self, is_b, comp):
        e = self._e
        ext_on = (e.ext_fld.B_type if is_b else e.ext_fld.E_type) != A.EMF_FLD_TYPE_NONE
        if not ext_on:
            return self._view(e.B_buf if is_b else e.E_buf, comp)
        _L().zpic_b200_sync_emf(self._p)
        nx, ny = e.nx[0], e.nx[1]
        g = A.grid_view(e.ext_fld.B_part_buf if is_b else e.ext_fld.E_part_buf, nx, ny)
        return g[1:ny + 1, 1:nx + 1, comp]

    Ex_part = property(lambda s: s._part_view(False, 0))
    Ey_part = property(lambda s: s._part_view(False, 1))
    Ez_part = property(lambda s: s._part_view(False, 2))
    Bx_part = property(lambda s: s._part_view(True, 0))
    By_part = property(lambda s: s._part_view(True, 1))
    Bz_part = property(lambda s: s._part_view(True, 2))


class Laser:
    """Laser pulse (em2d.pyx:1297-1534)"""
    _types = {"plane": A.PLANE, "gaussian": A.GAUSSIAN}

    def __init__(self, *, type="plane", start=0.0, fwhm=0.0, rise=0.0, flat=0.0, fall=0.0, a0=0.0,
                 omega0=0.0, polarization=0.0, W0=0.0, focus=0.0, axis=0.0):
        self._c = A.Laser(self._types[type], start, fwhm, rise, flat, fall, a0, omega0, polarization, W0, focus, axis)


for _name in ("start", "fwhm", "rise", "flat", "fall", "a0", "omega0", "polarization", "W0", "focus", "axis"):
    setattr(Laser, _name, property(lambda s, n=_name: getattr(s._c, n), lambda s, v, n=_name: setattr(s._c, n, v)))


class Current:
    """Electric current density of a simulation (em2d.pyx:1535-1621)"""

    def __init__(self, ptr):
        self._p = ptr

    def report(self, jc):
        _L().current_report(self._p, jc)

    def _view(self, comp):
        _L().zpic_b200_sync_current(self._p)
        c = self._p.contents
        nx, ny = c.nx[0], c.nx[1]
        return A.grid_view(c.J_buf, nx, ny)[1:ny + 1, 1:nx + 1, comp]

    Jx = property(lambda s: s._view(0))
    Jy = property(lambda s: s._view(1))
    Jz = property(lambda s: s._view(2))


class Smooth:
    """Digital filtering parameters (em2d.pyx:1622-1725)"""
    _types = {"none": A.SMOOTH_NONE, "binomial": A.BINOMIAL, "compensated": A.COMPENSATED}

    def __init__(self, *, xtype="none", ytype="none", xlevel=0, ylevel=0):
        self._c = A.Smooth(self._types[xtype], self._types[ytype], xlevel, ylevel)

    xlevel = property(lambda s: s._c.xlevel, lambda s, v: setattr(s._c, "xlevel", v))
    ylevel = property(lambda s: s._c.ylevel, lambda s, v: setattr(s._c, "ylevel", v))


class Simulation:
    """EM2D simulation (em2d.pyx:1726-2074)"""

    def __init__(self, nx, box, dt, *, species=None, report=None, mov_window=False, smooth=None,
                 init_fld=None, ext_fld=None):
        L = _L()
        self._sim = A.Simulation()
        self._p = C.pointer(self._sim)
        L.set_rand_seed(12345, 67890)          # as the reference does (em2d.pyx:1781)
        self._nx = (C.c_int * 2)(*[int(v) for v in nx])
        self._box = (C.c_float * 2)(*box)
        if isinstance(species, Species):
            species = [species]
        self._species = list(species) if species else []
        n = len(self._species)
        if n:
            raw = _libc.calloc(n, C.sizeof(A.Species))      # freed by sim_delete with free()
            arr = C.cast(raw, C.POINTER(A.Species))
            for k, s in enumerate(self._species):
                s._new(C.pointer(arr[k]), self._nx, self._box, dt)
        else:
            arr = C.POINTER(A.Species)()
        self.report = report
        L.sim_new(self._p, self._nx, self._box, dt, 0.0, 0, arr, n)
        self.n = 0
        self.t = 0.0
        self.emf = EMF(C.pointer(self._sim.emf))
        self.current = Current(C.pointer(self._sim.current))
        if mov_window:
            L.sim_set_moving_window(self._p)
        if smooth:
            L.sim_set_smooth(self._p, C.byref(smooth._c))
        if init_fld:
            self.emf.init_fld(init_fld)
        if ext_fld:
            self.emf.set_ext_fld(ext_fld)
        self._alive = True

    def __del__(self):
        if getattr(self, "_alive", False):
            self._alive = False
            _L().sim_delete(self._p)

    def set_moving_window(self):
        _L().sim_set_moving_window(self._p)

    def set_smooth(self, smooth):
        _L().sim_set_smooth(self._p, C.byref(smooth._c))

    def add_laser(self, laser):
        _L().sim_add_laser(self._p, C.byref(laser._c))

    def iter(self):
        """advance one iteration (em2d.pyx:1895-1902)"""
        _L().sim_iter(self._p)
        self.n += 1
        self.t = self.n * self._sim.dt

    def run(self, tmax):
        """advance up to time tmax, calling `report` before every iteration (em2d.pyx:1904-1940)"""
        if tmax < self.t:
            print("Simulation is already at t = {:g}".format(self.t))
            return
        print("\nRunning simulation up to t = {:g} ...".format(tmax))
        while self.t <= tmax:
            print("n = {:d}, t = {:g}".format(self.n, self.t), end="\r")
            if self.report:
                self.report(self)
            self.iter()
        print("n = {:d}, t = {:g}".format(self.n, self.t), end="\r")
        print("\nDone.")

    def sync(self):
        """extension: make every host mirror current (device -> host)"""
        _L().zpic_b200_sync_host(self._p)

    species = property(lambda s: s._species)
    dt = property(lambda s: s._sim.dt)
    nx = property(lambda s: np.array(s._sim.emf.nx[:], dtype=np.int32))
    dx = property(lambda s: np.array(s._sim.emf.dx[:], dtype=np.float32))
    box = property(lambda s: np.array(s._sim.emf.box[:], dtype=np.float32))
