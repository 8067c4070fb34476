"""Locate and load the native libraries.  There is no Python / CPU fallback: if the
CUDA library is missing the import fails loudly."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_cache = {}


def lib_path(code="em2d"):
    # ZPIC_LIB_SUFFIX selects an experimental build variant (see zpic_b200/build.py)
    return os.path.join(_HERE, "lib", "libzpic_b200_%s%s.so" % (code, os.environ.get("ZPIC_LIB_SUFFIX", "")))


def load(code="em2d"):
    """ctypes handle of libzpic_b200_<code>.so (built by zpic_b200.build)."""
    if code in _cache:
        return _cache[code]
    path = lib_path(code)
    if not os.path.exists(path):
        raise ImportError(
            "zpic_b200: native library %s not found. Build it with `python -m zpic_b200.build` "
            "(needs nvcc); there is no CPU fallback." % path)
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    _declare_dev(lib)
    if code == "em2d":
        from . import abi_em2d
        abi_em2d.declare(lib)
    elif code == "em1d":
        from . import abi_em1d
        abi_em1d.declare(lib)
    _cache[code] = lib
    return lib


class PushParams1D(C.Structure):   # zdev_push1d_params (include/zpic_dev.h)
    _fields_ = [("tem", C.c_float), ("dt_dx", C.c_float), ("qnx", C.c_float), ("q", C.c_float),
                ("absorbing", C.c_int), ("shift_window", C.c_int)]


class PushParams2D(C.Structure):   # zdev_push2d_params (include/zpic_dev.h)
    _fields_ = [("tem", C.c_float), ("dt_dx", C.c_float), ("dt_dy", C.c_float),
                ("qnx", C.c_float), ("qny", C.c_float), ("q", C.c_float),
                ("moving_window", C.c_int), ("shift_window", C.c_int),
                ("slab_left", C.c_int), ("slab_right", C.c_int)]


def _declare_dev(lib):
    """argument types of the device seam (include/zpic_dev.h)"""
    vp, fp, i, f = C.c_void_p, C.POINTER(C.c_float), C.c_int, C.c_float
    sig = {
        "zdev_init": (i, [i]), "zdev_ready": (i, []), "zdev_sync": (None, []),
        "zdev_stream": (vp, []), "zdev_launch_count": (C.c_uint64, []),
        "zdev_event_create": (vp, []), "zdev_event_record": (None, [vp]),
        "zdev_event_elapsed_ms": (f, [vp, vp]), "zdev_event_destroy": (None, [vp]),
        "zdev_mem_info": (None, [C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
        "zdev_flush_l2": (None, []),
    }
    sig2d = {
        "zdev_grid2d_create": (vp, [i, i]), "zdev_grid2d_destroy": (None, [vp]),
        "zdev_grid2d_upload": (None, [vp, i, vp]), "zdev_grid2d_download": (None, [vp, i, vp]),
        "zdev_grid2d_ptr": (vp, [vp, i]),
        "zdev_current_zero": (None, [vp]),
        "zdev_current_update": (None, [vp, i, i, i, i, i]),
        "zdev_current_update_gc": (None, [vp, i]),
        "zdev_current_smooth": (None, [vp, i, i, i, i, i]),
        "zdev_emf_set_ext_uniform": (None, [vp, i, fp, i, fp]),
        "zdev_emf_set_ext_grid": (None, [vp, vp, vp]),
        "zdev_emf_advance": (None, [vp, vp, f, f, f, i, i]),
        "zdev_yee_set_fused": (None, [i]),
        "zdev_yee_b": (None, [vp, f, f]), "zdev_yee_e": (None, [vp, vp, f, f, f]),
        "zdev_emf_update_gc": (None, [vp, i]), "zdev_emf_move_window": (None, [vp]),
        "zdev_emf_energy": (None, [vp, C.POINTER(C.c_double)]),
        "zdev_spec2d_create": (vp, [i, i, i, i]), "zdev_spec2d_destroy": (None, [vp]),
        "zdev_spec2d_upload": (None, [vp, vp, C.c_int64]),
        "zdev_spec2d_append": (None, [vp, vp, C.c_int64]),
        "zdev_spec2d_download": (C.c_int64, [vp, vp, C.c_int64]),
        "zdev_spec2d_np": (C.c_int64, [vp]),
        "zdev_spec2d_inject_uniform": (None, [vp, i, i, fp, fp, C.c_uint64]),
        "zdev_spec2d_inject_band": (None, [vp, i, i, fp, fp, C.c_uint64, i, i]),
        "zdev_spec2d_inject_rect": (None, [vp, i, i, fp, fp, C.c_uint64, i, i, i, i]),
        "zdev_spec2d_advance": (None, [vp, vp, vp, C.POINTER(PushParams2D)]),
        "zdev_spec2d_fetch": (None, [vp, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
        "zdev_spec2d_deposit_charge": (None, [vp, f, i, fp]),
        "zdev_spec2d_tile_info": (None, [vp, C.POINTER(i), C.POINTER(i), C.POINTER(i), C.POINTER(C.c_int64)]),
        "zdev_spec2d_export_counts": (None, [vp, C.POINTER(C.c_int64)]),
        "zdev_spec2d_export_ptr": (vp, [vp, i]),
        "zdev_spec2d_append_device": (None, [vp, vp, C.c_int64]),
        "zdev_grid2d_pack_cols": (None, [vp, i, i, i, i, i, vp]),
        "zdev_grid2d_unpack_cols": (None, [vp, i, i, i, i, i, vp, i]),
        "zdev_current_fold_y": (None, [vp]),
        "zdev_smooth_plan": (i, [i, i, i, i, C.POINTER(i), fp, fp]),
        "zdev_smooth_pass": (None, [vp, i, f, f, i]),
        "zdev_emf_shift": (None, [vp, i]),
        "zdev_emf_update_part_fld": (None, [vp]),
        "zdev_set_stream": (None, [vp]),
    }
    sig1d = {
        "zdev_grid1d_create": (vp, [i]), "zdev_grid1d_destroy": (None, [vp]),
        "zdev_grid1d_upload": (None, [vp, i, vp]), "zdev_grid1d_download": (None, [vp, i, vp]),
        "zdev_current1d_zero": (None, [vp]), "zdev_current1d_update": (None, [vp, i, i, i]),
        "zdev_emf1d_advance": (None, [vp, vp, f, f, i, i]),
        "zdev_emf1d_set_mur": (None, [vp, fp]), "zdev_emf1d_get_mur": (None, [vp, fp]),
        "zdev_emf1d_energy": (None, [vp, C.POINTER(C.c_double)]),
        "zdev_spec1d_create": (vp, [i, i, i]), "zdev_spec1d_destroy": (None, [vp]),
        "zdev_spec1d_upload": (None, [vp, vp, C.c_int64]), "zdev_spec1d_append": (None, [vp, vp, C.c_int64]),
        "zdev_spec1d_download": (C.c_int64, [vp, vp, C.c_int64]), "zdev_spec1d_np": (C.c_int64, [vp]),
        "zdev_spec1d_inject_uniform": (None, [vp, i, fp, fp, C.c_uint64]),
        "zdev_spec1d_advance": (None, [vp, vp, vp, C.POINTER(PushParams1D)]),
        "zdev_spec1d_fetch": (None, [vp, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
        "zdev_spec1d_deposit_charge": (None, [vp, f, i, fp]),
        "zdev_spec1d_push_timing": (None, [vp, C.POINTER(C.c_double), C.POINTER(C.c_int64), i]),
    }
    for table in (sig, sig2d, sig1d):
        for name, (res, args) in table.items():
            if hasattr(lib, name):
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
    for name in ("zpic_b200_sync_host", "zpic_b200_touch_host", "zpic_b200_sync_species",
                 "zpic_b200_sync_emf", "zpic_b200_sync_current", "zpic_b200_touch_species",
                 "zpic_b200_touch_emf"):
        if hasattr(lib, name):
            getattr(lib, name).restype = None
            getattr(lib, name).argtypes = [vp]
    for name in ("zpic_b200_species_handle", "zpic_b200_grid_handle"):
        if hasattr(lib, name):
            getattr(lib, name).restype = vp
            getattr(lib, name).argtypes = [vp]
    for name, (res, args) in {"zdev_set_push_timing": (None, [i]),
                              "zdev_spec2d_push_timing": (None, [vp, C.POINTER(C.c_double), C.POINTER(C.c_int64), i])}.items():
        if hasattr(lib, name):
            getattr(lib, name).restype = res
            getattr(lib, name).argtypes = args
    if hasattr(lib, "zpic_b200_set_option"):
        lib.zpic_b200_set_option.restype = None
        lib.zpic_b200_set_option.argtypes = [C.c_char_p, i]


def spec_handle(lib, spec_ptr):
    """zdev_spec2d* behind a t_species (uploads / creates the device copy if needed)"""
    return lib.zpic_b200_species_handle(spec_ptr)
