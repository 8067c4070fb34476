"""ctypes front end of the CPU restatement oracle/liboracle_em2d.so (tests only)."""
import ctypes as C
import os
import subprocess

import numpy as np

from zpic_b200.abi_em2d import PART_DTYPE

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


class OrcSpecies(C.Structure):
    _fields_ = [("part", C.c_void_p), ("np", C.c_int), ("m_q", C.c_float), ("q", C.c_float),
                ("energy", C.c_double), ("iter", C.c_int), ("n_move", C.c_int), ("n_sort", C.c_int),
                ("inject", C.c_void_p), ("inject_ctx", C.c_void_p)]


class OrcSim(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("dx", C.c_float), ("dy", C.c_float), ("dt", C.c_float),
                ("E", C.c_void_p), ("B", C.c_void_p), ("J", C.c_void_p),
                ("iter", C.c_int), ("n_move", C.c_int), ("moving_window", C.c_int),
                ("xtype", C.c_int), ("ytype", C.c_int), ("xlevel", C.c_int), ("ylevel", C.c_int),
                ("n_species", C.c_int), ("species", C.POINTER(OrcSpecies))]


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(REPO, "oracle", "liboracle_em2d.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", os.path.join(REPO, "oracle"), "restate"], stdout=subprocess.DEVNULL)
        _lib = C.CDLL(path)
        _lib.orc2d_spec_push.restype = C.c_double
        _lib.orc2d_spec_boundary.restype = C.c_int
    return _lib


class OracleSim:
    """The oracle's own simulation state, initialised from the raw buffers of a Deck (any library)."""

    def __init__(self, deck, n_sort=None):
        s = deck.sim
        self.nx, self.ny = deck.nx
        self.E = deck.E().copy()
        self.B = deck.B().copy()
        self.J = np.zeros_like(self.E)
        self.parts = []
        self.spec = (OrcSpecies * max(deck.n_species, 1))()
        for k in range(deck.n_species):
            sp = deck.species[k]
            buf = np.zeros(max(sp.np, 1) + 1024, dtype=PART_DTYPE)
            buf[:sp.np] = deck.parts(k)
            self.parts.append(buf)
            o = self.spec[k]
            o.part = buf.ctypes.data
            o.np = sp.np
            o.m_q = sp.m_q
            o.q = sp.q
            o.iter = sp.iter
            o.n_move = sp.n_move
            o.n_sort = sp.n_sort if n_sort is None else n_sort
        c = s.current
        self.sim = OrcSim(self.nx, self.ny, s.emf.dx[0], s.emf.dx[1], s.dt,
                          self.E.ctypes.data, self.B.ctypes.data, self.J.ctypes.data,
                          s.emf.iter, s.emf.n_move, s.emf.moving_window,
                          c.smooth.xtype, c.smooth.ytype, c.smooth.xlevel, c.smooth.ylevel,
                          deck.n_species, self.spec)

    def iter(self, n=1):
        L = lib()
        for _ in range(n):
            L.orc2d_sim_iter(C.byref(self.sim))

    def part(self, k):
        return self.parts[k][:self.spec[k].np]

    def energy(self, k):
        return self.spec[k].energy
