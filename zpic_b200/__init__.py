"""zpic_b200 - B200-native em2d / em1d PIC time step behind the ZPIC C API.

The product is the native library (zpic_b200/lib/libzpic_b200_<code>.so): host C that
exports the reference's own API (include/em2d, include/em1d) and drives hand-written
sm_100a CUDA kernels through the C-ABI seam in include/zpic_dev.h.  This package only
loads that library (`zpic_b200.load`) and mirrors the reference's Python classes on
top of it (`zpic_b200.em2d`).
"""
from ._lib import load, lib_path  # noqa: F401
