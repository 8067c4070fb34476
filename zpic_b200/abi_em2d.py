"""ctypes view of the em2d C API (include/em2d/*.h == reference em2d/*.h).

The same declarations drive BOTH shared objects that export this API:
  * zpic_b200/lib/libzpic_b200_em2d.so  - the product (host C + CUDA kernels)
  * oracle/_ref/libzpic_ref_em2d.so     - the unmodified reference (tests only)
so a parity test is literally the same call sequence run on two libraries.
"""
import ctypes as C
import numpy as np

MAX_SPNAME_LEN = 32


class Float3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class Part(C.Structure):          # t_part, 28 bytes (em2d/particles.h:29-37)
    _fields_ = [("ix", C.c_int), ("iy", C.c_int), ("x", C.c_float), ("y", C.c_float),
                ("ux", C.c_float), ("uy", C.c_float), ("uz", C.c_float)]


PART_DTYPE = np.dtype([("ix", "<i4"), ("iy", "<i4"), ("x", "<f4"), ("y", "<f4"),
                       ("ux", "<f4"), ("uy", "<f4"), ("uz", "<f4")])

DENSITY_FN = C.CFUNCTYPE(C.c_float, C.c_float, C.c_void_p)
FIELD_FN = C.CFUNCTYPE(Float3, C.c_int, C.c_float, C.c_int, C.c_float, C.c_void_p)

UNIFORM, EMPTY, STEP, SLAB, CUSTOM = range(5)
SMOOTH_NONE, BINOMIAL, COMPENSATED = range(3)
EMF_FLD_TYPE_NONE, EMF_FLD_TYPE_UNIFORM, EMF_FLD_TYPE_CUSTOM = range(3)
PLANE, GAUSSIAN = range(2)
EFLD, BFLD, EPART, BPART = range(4)


class Density(C.Structure):       # em2d/particles.h:55-78
    _fields_ = [("n", C.c_float), ("type", C.c_int), ("start", C.c_float), ("end", C.c_float),
                ("custom_x", DENSITY_FN), ("custom_data_x", C.c_void_p),
                ("custom_y", DENSITY_FN), ("custom_data_y", C.c_void_p),
                ("custom_x_total_part", C.c_ulong), ("custom_x_total_q", C.c_double)]


class Species(C.Structure):       # em2d/particles.h:85-132
    _fields_ = [("name", C.c_char * (MAX_SPNAME_LEN + 1)),
                ("part", C.POINTER(Part)), ("np", C.c_int), ("np_max", C.c_int),
                ("m_q", C.c_float), ("energy", C.c_double), ("q", C.c_float),
                ("ppc", C.c_int * 2), ("density", Density),
                ("ufl", C.c_float * 3), ("uth", C.c_float * 3),
                ("nx", C.c_int * 2), ("dx", C.c_float * 2), ("box", C.c_float * 2),
                ("dt", C.c_float), ("iter", C.c_int),
                ("moving_window", C.c_int), ("n_move", C.c_int), ("n_sort", C.c_int)]


class Smooth(C.Structure):        # em2d/current.h:29-34
    _fields_ = [("xtype", C.c_int), ("ytype", C.c_int), ("xlevel", C.c_int), ("ylevel", C.c_int)]


class Current(C.Structure):       # em2d/current.h:41-70
    _fields_ = [("J", C.POINTER(Float3)), ("J_buf", C.POINTER(Float3)),
                ("nx", C.c_int * 2), ("nrow", C.c_int), ("gc", (C.c_int * 2) * 2),
                ("box", C.c_float * 2), ("dx", C.c_float * 2), ("smooth", Smooth),
                ("dt", C.c_float), ("iter", C.c_int), ("moving_window", C.c_int)]


class ExtField(C.Structure):      # em2d/emf.h:28-44
    _fields_ = [("E_type", C.c_int), ("B_type", C.c_int), ("E_0", Float3), ("B_0", Float3),
                ("E_custom", FIELD_FN), ("B_custom", FIELD_FN),
                ("E_custom_data", C.c_void_p), ("B_custom_data", C.c_void_p),
                ("E_part_buf", C.POINTER(Float3)), ("B_part_buf", C.POINTER(Float3))]


class InitField(C.Structure):     # em2d/emf.h:50-65
    _fields_ = [("E_type", C.c_int), ("B_type", C.c_int), ("E_0", Float3), ("B_0", Float3),
                ("E_custom", FIELD_FN), ("B_custom", FIELD_FN),
                ("E_custom_data", C.c_void_p), ("B_custom_data", C.c_void_p)]


class EMF(C.Structure):           # em2d/emf.h:83-120
    _fields_ = [("E", C.POINTER(Float3)), ("B", C.POINTER(Float3)),
                ("E_buf", C.POINTER(Float3)), ("B_buf", C.POINTER(Float3)),
                ("E_part", C.POINTER(Float3)), ("B_part", C.POINTER(Float3)),
                ("nx", C.c_int * 2), ("nrow", C.c_int), ("gc", (C.c_int * 2) * 2),
                ("box", C.c_float * 2), ("dx", C.c_float * 2), ("dt", C.c_float),
                ("iter", C.c_int), ("moving_window", C.c_int), ("n_move", C.c_int),
                ("ext_fld", ExtField)]


class Laser(C.Structure):         # em2d/emf.h:135-157
    _fields_ = [("type", C.c_int), ("start", C.c_float), ("fwhm", C.c_float),
                ("rise", C.c_float), ("flat", C.c_float), ("fall", C.c_float),
                ("a0", C.c_float), ("omega0", C.c_float), ("polarization", C.c_float),
                ("W0", C.c_float), ("focus", C.c_float), ("axis", C.c_float)]


class Simulation(C.Structure):    # em2d/simulation.h:13-29
    _fields_ = [("dt", C.c_float), ("tmax", C.c_float), ("ndump", C.c_int),
                ("n_species", C.c_int), ("species", C.POINTER(Species)),
                ("emf", EMF), ("current", Current), ("moving_window", C.c_int)]


def declare(lib):
    """Attach argument / result types of the em2d API to a loaded library."""
    P = C.POINTER
    i2, f2 = C.c_int * 2, C.c_float * 2
    lib.spec_new.argtypes = [P(Species), C.c_char_p, C.c_float, P(C.c_int), P(C.c_float), P(C.c_float),
                             P(C.c_int), P(C.c_float), C.c_float, P(Density)]
    lib.spec_new.restype = None
    lib.spec_delete.argtypes = [P(Species)]
    lib.spec_grow_buffer.argtypes = [P(Species), C.c_int]
    lib.spec_advance.argtypes = [P(Species), P(EMF), P(Current)]
    lib.spec_deposit_charge.argtypes = [P(Species), P(C.c_float)]
    lib.spec_report.argtypes = [P(Species), C.c_int, C.c_void_p, C.c_void_p]
    lib.spec_npush.restype = C.c_uint64
    lib.spec_time.restype = C.c_double
    lib.spec_perf.restype = C.c_double
    lib.emf_new.argtypes = [P(EMF), P(C.c_int), P(C.c_float), C.c_float]
    lib.emf_delete.argtypes = [P(EMF)]
    lib.emf_advance.argtypes = [P(EMF), P(Current)]
    lib.emf_add_laser.argtypes = [P(EMF), P(Laser)]
    lib.emf_get_energy.argtypes = [P(EMF), P(C.c_double)]
    lib.emf_set_ext_fld.argtypes = [P(EMF), P(ExtField)]
    lib.emf_init_fld.argtypes = [P(EMF), P(InitField)]
    lib.emf_report.argtypes = [P(EMF), C.c_char, C.c_int]
    lib.emf_time.restype = C.c_double
    lib.current_new.argtypes = [P(Current), P(C.c_int), P(C.c_float), C.c_float]
    lib.current_delete.argtypes = [P(Current)]
    lib.current_zero.argtypes = [P(Current)]
    lib.current_update.argtypes = [P(Current)]
    lib.current_report.argtypes = [P(Current), C.c_int]
    lib.sim_new.argtypes = [P(Simulation), P(C.c_int), P(C.c_float), C.c_float, C.c_float, C.c_int,
                            P(Species), C.c_int]
    lib.sim_iter.argtypes = [P(Simulation)]
    lib.sim_delete.argtypes = [P(Simulation)]
    lib.sim_add_laser.argtypes = [P(Simulation), P(Laser)]
    lib.sim_set_smooth.argtypes = [P(Simulation), P(Smooth)]
    lib.sim_set_moving_window.argtypes = [P(Simulation)]
    lib.sim_set_ext_fld.argtypes = [P(Simulation), P(ExtField)]
    lib.sim_report_energy.argtypes = [P(Simulation)]
    lib.set_rand_seed.argtypes = [C.c_uint32, C.c_uint32]
    lib.rand_norm.restype = C.c_double
    for name in ("spec_delete", "spec_grow_buffer", "spec_advance", "spec_deposit_charge", "spec_report",
                 "emf_new", "emf_delete", "emf_advance", "emf_add_laser", "emf_get_energy",
                 "emf_set_ext_fld", "emf_init_fld", "emf_report", "current_new", "current_delete",
                 "current_zero", "current_update", "current_report", "sim_new", "sim_iter", "sim_delete",
                 "sim_add_laser", "sim_set_smooth", "sim_set_moving_window", "sim_set_ext_fld",
                 "sim_report_energy", "set_rand_seed"):
        getattr(lib, name).restype = None
    return lib


# ---------------------------------------------------------------- numpy views of raw buffers

def grid_view(ptr_buf, nx, ny):
    """(ny+3, nx+3, 3) float32 view of an E_buf / B_buf / J_buf (guards included)."""
    n = (nx + 3) * (ny + 3) * 3
    a = np.ctypeslib.as_array(C.cast(ptr_buf, C.POINTER(C.c_float)), shape=(n,))
    return a.reshape(ny + 3, nx + 3, 3)


def part_view(spec):
    """structured view of spec.part[0:np]"""
    n = spec.np
    if n <= 0:
        return np.zeros(0, dtype=PART_DTYPE)
    raw = np.ctypeslib.as_array(C.cast(spec.part, C.POINTER(C.c_uint8)), shape=(n * 28,))
    return raw.view(PART_DTYPE)
