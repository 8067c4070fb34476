"""CPU stand-in for zpic_b200.parallel.CudaSlab built on the oracle restatement (tests only): lets the
slab-decomposition logic (which columns travel, ring vs chain, window shift, particle hand-over) run
under gloo with world_size 2 and be compared bit for bit with the single-slab oracle."""
import ctypes as C

import numpy as np
import torch

from tests import oracle as O
from zpic_b200.abi_em2d import PART_DTYPE


class OracleSlab:
    def __init__(self, geom, dt, dx, dy, species, smooth=(0, 0, 0, 0)):
        self.g = geom
        self.dt, self.dx, self.dy = np.float32(dt), np.float32(dx), np.float32(dy)
        self.smooth = smooth
        shape = (geom.ny + 3, geom.nxl + 3, 3)
        self.grids = [np.zeros(shape, dtype=np.float32) for _ in range(3)]      # E, B, J
        self.species = [dict(sp, part=np.zeros(0, dtype=PART_DTYPE), iter=0, n_move=0, energy=0.0) for sp in species]
        self.exports = [[np.zeros(0, dtype=PART_DTYPE)] * 2 for _ in species]
        self.iter, self.n_move = 0, 0
        self.L = O.lib()

    # buffers are CPU torch tensors so the same Comm classes work
    def new_grid_buffer(self, ngrids, ncols, nrows):
        return torch.empty(ngrids * ncols * nrows * 3, dtype=torch.float32)

    def new_part_buffer(self, n):
        return torch.empty(n * 7, dtype=torch.int32)

    def new_counts(self, values=None):
        t = torch.zeros(2 * len(self.species), dtype=torch.int64)
        if values is not None:
            t.copy_(torch.tensor(values, dtype=torch.int64))
        return t

    def upload_particles(self, k, part):
        self.species[k]["part"] = np.ascontiguousarray(part).copy()

    def push(self, k, shift):
        sp, g = self.species[k], self.g
        q, m_q = np.float32(sp["q"]), np.float32(sp["m_q"])
        prm = (C.c_float * 6)(np.float32(0.5 * float(self.dt) / float(m_q)), self.dt / self.dx, self.dt / self.dy,
                              q * self.dx / self.dt, q * self.dy / self.dt, q)
        p = sp["part"]
        Ee, Bb, Jj = self.grids
        e = self.L.orc2d_spec_push(p.ctypes.data_as(C.c_void_p), len(p), Ee.ctypes.data_as(C.c_void_p),
                                   Bb.ctypes.data_as(C.c_void_p), Jj.ctypes.data_as(C.c_void_p), g.nxl, g.ny, prm)
        sp["energy"] = e
        if shift:
            p["ix"] -= 1
        nx, ny = g.nxl, g.ny
        lo, hi = p["ix"] < 0, p["ix"] >= nx
        exp_l = p[lo].copy() if g.left is not None else np.zeros(0, dtype=PART_DTYPE)
        exp_r = p[hi].copy() if g.right is not None else np.zeros(0, dtype=PART_DTYPE)
        exp_l["ix"] += nx
        exp_r["ix"] -= nx
        keep = ~(lo | hi)
        if g.left is None and not g.window:
            p["ix"][lo] += nx
            keep |= lo
        if g.right is None and not g.window:
            p["ix"][hi] -= nx
            keep |= hi
        p = p[keep]
        p["iy"] += np.where(p["iy"] < 0, ny, 0) - np.where(p["iy"] >= ny, ny, 0)
        for e_ in (exp_l, exp_r):
            e_["iy"] += np.where(e_["iy"] < 0, ny, 0) - np.where(e_["iy"] >= ny, ny, 0)
        sp["part"] = p
        self.exports[k] = [exp_l, exp_r]

    def export_counts(self, k):
        return len(self.exports[k][0]), len(self.exports[k][1])

    def export_buffer(self, k, side, n):
        return torch.from_numpy(self.exports[k][side].view(np.int32).copy())

    def import_particles(self, k, buf):
        if buf.numel():
            rec = buf.numpy().view(PART_DTYPE)
            self.species[k]["part"] = np.concatenate([self.species[k]["part"], rec])

    def append_host_particles(self, k, part):
        if len(part):
            self.species[k]["part"] = np.concatenate([self.species[k]["part"], part])

    def download_particles(self, k):
        return self.species[k]["part"]

    def upload_grid(self, which, arr):
        self.grids[which][...] = arr

    def download_grid(self, which):
        return self.grids[which]

    def _cols(self, which, i0, ncols, j0, nrows):
        return self.grids[which][j0 + 1: j0 + 1 + nrows, i0 + 1: i0 + 1 + ncols, :]

    def pack(self, whichs, i0, ncols, j0, nrows, out):
        n = ncols * nrows * 3
        o = out.numpy()
        for k, w in enumerate(whichs):
            o[k * n:(k + 1) * n] = self._cols(w, i0, ncols, j0, nrows).reshape(-1)

    def unpack(self, whichs, i0, ncols, j0, nrows, buf, add):
        n = ncols * nrows * 3
        b = buf.numpy()
        for k, w in enumerate(whichs):
            v = b[k * n:(k + 1) * n].reshape(nrows, ncols, 3)
            if add:
                self._cols(w, i0, ncols, j0, nrows)[...] += v
            else:
                self._cols(w, i0, ncols, j0, nrows)[...] = v

    def _p(self, which):
        return self.grids[which].ctypes.data_as(C.c_void_p)

    def current_zero(self):
        self.grids[2][...] = 0

    def current_fold_x_local(self):
        self.L.orc2d_current_gc(self._p(2), self.g.nxl, self.g.ny, 0)

    def current_fold_y(self):
        self.L.orc2d_current_gc(self._p(2), self.g.nxl, self.g.ny, 1)

    def smooth_plan(self):
        dirs, sa, sb = (C.c_int * 64)(), (C.c_float * 64)(), (C.c_float * 64)()
        n = self.L.orc2d_smooth_plan(*self.smooth, dirs, sa, sb)
        return [(dirs[i], sa[i], sb[i]) for i in range(n)]

    def smooth_pass(self, d, sa, sb, keep_x_guards):
        self.L.orc2d_smooth_pass(self._p(2), self.g.nxl, self.g.ny, d, C.c_float(sa), C.c_float(sb), int(keep_x_guards))

    def yee_b(self):
        dth = self.dt / np.float32(2.0)
        self.L.orc2d_yee_b(self._p(1), self._p(0), self.g.nxl, self.g.ny, C.c_float(dth / self.dx), C.c_float(dth / self.dy))

    def yee_e(self):
        self.L.orc2d_yee_e(self._p(0), self._p(1), self._p(2), self.g.nxl, self.g.ny, C.c_float(self.dt / self.dx),
                           C.c_float(self.dt / self.dy), C.c_float(self.dt))

    def emf_gc(self, skip_x):
        for w in (0, 1):
            self.L.orc2d_guard_copy(self._p(w), self.g.nxl, self.g.ny, int(skip_x))

    def emf_shift(self, zero_right):
        for w in (0, 1):
            a = self.grids[w]
            a[:, :-1, :] = a[:, 1:, :].copy()
            if zero_right:
                a[:, self.g.nxl:, :] = 0          # buffer columns nxl.. = cells nxl-1, nxl, nxl+1

    def emf_part_fld(self):
        pass

    def sync(self):
        pass
