"""ctypes view of the em1d C API (include/em1d/*.h == reference em1d/*.h); used for the product
library and for oracle/_ref/libzpic_ref_em1d.so alike."""
import ctypes as C

import numpy as np

from .abi_em2d import Float3

MAX_SPNAME_LEN = 32


class Part(C.Structure):          # t_part, 20 bytes (em1d/particles.h:29-35)
    _fields_ = [("ix", C.c_int), ("x", C.c_float), ("ux", C.c_float), ("uy", C.c_float), ("uz", C.c_float)]


PART_DTYPE = np.dtype([("ix", "<i4"), ("x", "<f4"), ("ux", "<f4"), ("uy", "<f4"), ("uz", "<f4")])
DENSITY_FN = C.CFUNCTYPE(C.c_float, C.c_float, C.c_void_p)
FIELD_FN = C.CFUNCTYPE(Float3, C.c_int, C.c_float, C.c_void_p)

UNIFORM, EMPTY, STEP, SLAB, RAMP, CUSTOM = range(6)
SMOOTH_NONE, BINOMIAL, COMPENSATED = range(3)
EMF_FLD_TYPE_NONE, EMF_FLD_TYPE_UNIFORM, EMF_FLD_TYPE_CUSTOM = range(3)
EMF_BC_NONE, EMF_BC_PERIODIC, EMF_BC_OPEN = range(3)
CURRENT_BC_NONE, CURRENT_BC_PERIODIC = range(2)
PART_BC_NONE, PART_BC_PERIODIC, PART_BC_OPEN = range(3)
EFLD, BFLD, EPART, BPART = range(4)


class Density(C.Structure):       # em1d/particles.h:52-70
    _fields_ = [("n", C.c_float), ("type", C.c_int), ("start", C.c_float), ("end", C.c_float),
                ("ramp", C.c_float * 2), ("custom", DENSITY_FN), ("custom_data", C.c_void_p),
                ("total_np_inj", C.c_ulong), ("custom_q_inj", C.c_double)]


class Species(C.Structure):       # em1d/particles.h:86-135
    _fields_ = [("name", C.c_char * (MAX_SPNAME_LEN + 1)),
                ("part", C.POINTER(Part)), ("np", C.c_int), ("np_max", C.c_int),
                ("m_q", C.c_float), ("energy", C.c_double), ("q", C.c_float), ("ppc", C.c_int),
                ("density", Density), ("ufl", C.c_float * 3), ("uth", C.c_float * 3),
                ("nx", C.c_int), ("dx", C.c_float), ("box", C.c_float), ("dt", C.c_float), ("iter", C.c_int),
                ("moving_window", C.c_int), ("n_move", C.c_int), ("bc_type", C.c_int), ("n_sort", C.c_int)]


class Smooth(C.Structure):
    _fields_ = [("xtype", C.c_int), ("xlevel", C.c_int)]


class Current(C.Structure):       # em1d/current.h:47-75
    _fields_ = [("J", C.POINTER(Float3)), ("J_buf", C.POINTER(Float3)), ("nx", C.c_int), ("gc", C.c_int * 2),
                ("box", C.c_float), ("dx", C.c_float), ("smooth", Smooth), ("dt", C.c_float), ("iter", C.c_int),
                ("bc_type", C.c_int)]


class ExtField(C.Structure):
    _fields_ = [("E_type", C.c_int), ("B_type", C.c_int), ("E_0", Float3), ("B_0", Float3),
                ("E_custom", FIELD_FN), ("B_custom", FIELD_FN), ("E_custom_data", C.c_void_p), ("B_custom_data", C.c_void_p),
                ("E_part_buf", C.POINTER(Float3)), ("B_part_buf", C.POINTER(Float3))]


class EMF(C.Structure):           # em1d/emf.h:94-132
    _fields_ = [("E", C.POINTER(Float3)), ("B", C.POINTER(Float3)), ("E_buf", C.POINTER(Float3)), ("B_buf", C.POINTER(Float3)),
                ("E_part", C.POINTER(Float3)), ("B_part", C.POINTER(Float3)),
                ("nx", C.c_int), ("gc", C.c_int * 2), ("box", C.c_float), ("dx", C.c_float), ("dt", C.c_float),
                ("iter", C.c_int), ("moving_window", C.c_int), ("n_move", C.c_int), ("bc_type", C.c_int),
                ("mur_fld", Float3 * 2), ("mur_tmp", Float3 * 2), ("ext_fld", ExtField)]


class Laser(C.Structure):         # em1d/emf.h:139-154
    _fields_ = [("start", C.c_float), ("fwhm", C.c_float), ("rise", C.c_float), ("flat", C.c_float), ("fall", C.c_float),
                ("a0", C.c_float), ("omega0", C.c_float), ("polarization", C.c_float)]


class Simulation(C.Structure):
    _fields_ = [("dt", C.c_float), ("tmax", C.c_float), ("ndump", C.c_int), ("n_species", C.c_int),
                ("species", C.POINTER(Species)), ("emf", EMF), ("current", Current), ("moving_window", C.c_int)]


def declare(lib):
    P = C.POINTER
    lib.spec_new.argtypes = [P(Species), C.c_char_p, C.c_float, C.c_int, P(C.c_float), P(C.c_float),
                             C.c_int, C.c_float, C.c_float, P(Density)]
    lib.spec_delete.argtypes = [P(Species)]
    lib.spec_advance.argtypes = [P(Species), P(EMF), P(Current)]
    lib.spec_deposit_charge.argtypes = [P(Species), P(C.c_float)]
    lib.spec_report.argtypes = [P(Species), C.c_int, C.c_void_p, C.c_void_p]
    lib.spec_npush.restype = C.c_uint64
    lib.spec_time.restype = C.c_double
    lib.emf_new.argtypes = [P(EMF), C.c_int, C.c_float, C.c_float]
    lib.emf_advance.argtypes = [P(EMF), P(Current)]
    lib.emf_add_laser.argtypes = [P(EMF), P(Laser)]
    lib.emf_get_energy.argtypes = [P(EMF), P(C.c_double)]
    lib.emf_set_ext_fld.argtypes = [P(EMF), P(ExtField)]
    lib.emf_report.argtypes = [P(EMF), C.c_char, C.c_int]
    lib.current_new.argtypes = [P(Current), C.c_int, C.c_float, C.c_float]
    lib.current_update.argtypes = [P(Current)]
    lib.current_report.argtypes = [P(Current), C.c_int]
    lib.sim_new.argtypes = [P(Simulation), C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, P(Species), C.c_int]
    lib.sim_iter.argtypes = [P(Simulation)]
    lib.sim_delete.argtypes = [P(Simulation)]
    lib.sim_add_laser.argtypes = [P(Simulation), P(Laser)]
    lib.sim_set_smooth.argtypes = [P(Simulation), P(Smooth)]
    lib.sim_set_moving_window.argtypes = [P(Simulation)]
    lib.sim_set_ext_fld.argtypes = [P(Simulation), P(ExtField)]
    lib.set_rand_seed.argtypes = [C.c_uint32, C.c_uint32]
    for name in ("spec_new", "spec_delete", "spec_advance", "spec_deposit_charge", "spec_report", "emf_new", "emf_advance",
                 "emf_add_laser", "emf_get_energy", "emf_set_ext_fld", "emf_report", "current_new", "current_update",
                 "current_report", "sim_new", "sim_iter", "sim_delete", "sim_add_laser", "sim_set_smooth",
                 "sim_set_moving_window", "sim_set_ext_fld", "set_rand_seed"):
        getattr(lib, name).restype = None
    return lib


def grid_view(ptr_buf, nx):
    a = np.ctypeslib.as_array(C.cast(ptr_buf, C.POINTER(C.c_float)), shape=((nx + 3) * 3,))
    return a.reshape(nx + 3, 3)


def part_view(spec):
    n = spec.np
    if n <= 0:
        return np.zeros(0, dtype=PART_DTYPE)
    raw = np.ctypeslib.as_array(C.cast(spec.part, C.POINTER(C.c_uint8)), shape=(n * 20,))
    return raw.view(PART_DTYPE)
