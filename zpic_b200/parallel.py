"""Slab decomposition of the em2d time step along x: one process per GPU.

Rank r owns the cell columns [r*nxl, (r+1)*nxl) of the global grid with the reference's own guard
layout (1 lower, 2 upper) acting as halo; y stays whole.  With valid halos at the start of a step the
unchanged local loop bounds of yee_b / yee_e reproduce the global interior bit for bit (the same
argument that makes the reference's periodic run work, em2d/emf.c:514-521, 548-560, 573-608), so the
decomposed run differs from the single-slab run only by the order of floating point additions in J.

Exchanges per step (ring for periodic x, open chain for the moving window, whose two physical edges
do nothing - em2d/emf.c:581, current.c:124):
  1. particles that crossed a slab edge (after all species were pushed),
  2. J guard fold: the 3 aliased columns (nxl-1, nxl, nxl+1) <-> (-1, 0, 1) are swapped and ADDED on both
     sides - addition commutes, so both neighbours hold the same sums (current.c:128-137),
  3. one halo refresh per smoothing pass along x (current.c:346-352),
  4. E/B halo refresh after the field advance, and again after a window shift (emf.c:583-607, 659-670).

`slab_step` is written against two small interfaces so that the decomposition logic is testable without
a GPU: a `Backend` (the local slab: the CUDA library through ctypes, or - in tests only - the CPU oracle)
and a `Comm` (torch.distributed over NCCL / gloo, or an in-process loop-back joining several slabs).
"""
import ctypes as C

import numpy as np

E, B, J = 0, 1, 2


class Geometry:
    def __init__(self, nx_global, ny, nranks, rank, moving_window=False):
        if nx_global % nranks:
            raise ValueError("the number of cells along x must be divisible by the number of slabs")
        self.nx_global, self.ny, self.nranks, self.rank = nx_global, ny, nranks, rank
        self.nxl = nx_global // nranks
        self.x0 = rank * self.nxl
        self.window = bool(moving_window)
        ring = not self.window
        # a slab edge is "interior" when a neighbour rank sits behind it (always, on a periodic ring
        # of >= 2 slabs; never, for a single periodic slab, which wraps locally like the reference)
        self.left = (rank - 1) % nranks if (nranks > 1 and (ring or rank > 0)) else None
        self.right = (rank + 1) % nranks if (nranks > 1 and (ring or rank < nranks - 1)) else None
        self.is_last = rank == nranks - 1
        # the edge between the last and the first slab of a periodic ring is the box boundary
        self.wrap_left = ring and rank == 0
        self.wrap_right = ring and rank == nranks - 1


# ------------------------------------------------------------------------------------------ comms

class TorchComm:
    """neighbour exchange over torch.distributed (NCCL on GPUs, gloo on CPU)"""

    def __init__(self, geom):
        import torch.distributed as dist
        self.dist = dist
        self.g = geom

    def exchange(self, send_left, send_right, recv_left, recv_right):
        """send_* / recv_* are contiguous torch tensors or None.  Messages travelling to the left carry
        tag 0, to the right tag 1; receives are posted right-neighbour first so that with two ranks
        (left == right) the pairwise order matches on backends that ignore tags."""
        d, g = self.dist, self.g
        ops = []
        if send_left is not None and g.left is not None:
            ops.append(d.P2POp(d.isend, send_left, g.left, tag=0))
        if send_right is not None and g.right is not None:
            ops.append(d.P2POp(d.isend, send_right, g.right, tag=1))
        if recv_right is not None and g.right is not None:
            ops.append(d.P2POp(d.irecv, recv_right, g.right, tag=0))
        if recv_left is not None and g.left is not None:
            ops.append(d.P2POp(d.irecv, recv_left, g.left, tag=1))
        if ops:
            for r in d.batch_isend_irecv(ops):
                r.wait()

    def allreduce_sum(self, t):
        self.dist.all_reduce(t)
        return t


class LoopbackComm:
    """Joins several slabs living in ONE process (tests, fake multi-GPU on one device).  Each slab
    calls exchange() in turn; messages are matched when the last slab of the round arrives."""

    class Hub:
        def __init__(self, n):
            self.n = n
            self.pending = []

    def __init__(self, geom, hub):
        self.g, self.hub = geom, hub

    def exchange(self, send_left, send_right, recv_left, recv_right):
        self.hub.pending.append((self.g, send_left, send_right, recv_left, recv_right))
        if len(self.hub.pending) == self.hub.n:
            by_rank = {p[0].rank: p for p in self.hub.pending}
            for g, sl, sr, rl, rr in self.hub.pending:
                if rl is not None and g.left is not None:
                    rl.copy_(by_rank[g.left][2])      # what my left neighbour sent to its right
                if rr is not None and g.right is not None:
                    rr.copy_(by_rank[g.right][1])     # what my right neighbour sent to its left
            self.hub.pending = []


# ------------------------------------------------------------------------------------------ CUDA backend

class CudaSlab:
    """One slab on one GPU, driven through the device seam (include/zpic_dev.h)."""

    def __init__(self, lib, geom, dt, dx, dy, species, smooth=(0, 0, 0, 0), device=None):
        import torch
        from ._lib import PushParams2D
        self.torch, self.lib, self.g = torch, lib, geom
        self.PushParams2D = PushParams2D
        self.dt, self.dx, self.dy = np.float32(dt), np.float32(dx), np.float32(dy)
        self.smooth = smooth
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.grid = lib.zdev_grid2d_create(geom.nxl, geom.ny)
        self.species = []                      # dicts: handle, m_q, q, iter, n_move
        for sp in species:
            h = lib.zdev_spec2d_create(geom.nxl, geom.ny, int(sp["ppc"][0] * sp["ppc"][1]), 0)
            self.species.append(dict(sp, handle=h, iter=0, n_move=0, energy=0.0))
        self.iter, self.n_move = 0, 0

    # --- buffers -----------------------------------------------------------------------------
    def new_grid_buffer(self, ngrids, ncols, nrows):
        return self.torch.empty(ngrids * ncols * nrows * 3, dtype=self.torch.float32, device=self.device)

    def new_part_buffer(self, n):
        return self.torch.empty(max(n, 1) * 7, dtype=self.torch.int32, device=self.device)[: n * 7]

    def new_counts(self, values=None):
        t = self.torch.zeros(2 * len(self.species), dtype=self.torch.int64, device=self.device)
        if values is not None:
            t.copy_(self.torch.tensor(values, dtype=self.torch.int64))
        return t

    # --- particles ---------------------------------------------------------------------------
    def upload_particles(self, k, part_aos):
        a = np.ascontiguousarray(part_aos)
        self.lib.zdev_spec2d_upload(self.species[k]["handle"], a.ctypes.data, len(a))

    def inject_uniform(self, k, ppc, ufl, uth, seed):
        self.lib.zdev_spec2d_inject_uniform(self.species[k]["handle"], ppc[0], ppc[1], (C.c_float * 3)(*ufl),
                                            (C.c_float * 3)(*uth), seed)

    def inject_band(self, k, ppc, ufl, uth, seed, iy0, iy1):
        self.lib.zdev_spec2d_inject_band(self.species[k]["handle"], ppc[0], ppc[1], (C.c_float * 3)(*ufl),
                                         (C.c_float * 3)(*uth), seed, iy0, iy1)

    def push(self, k, shift):
        sp, g = self.species[k], self.g
        q, m_q = np.float32(sp["q"]), np.float32(sp["m_q"])
        prm = self.PushParams2D(float(np.float32(0.5 * float(self.dt) / float(m_q))),
                                float(self.dt / self.dx), float(self.dt / self.dy),
                                float(q * self.dx / self.dt), float(q * self.dy / self.dt), float(q),
                                int(g.window), int(shift), int(g.left is not None), int(g.right is not None))
        self.lib.zdev_spec2d_advance(sp["handle"], self.grid, self.grid, C.byref(prm))

    def export_counts(self, k):
        c = (C.c_int64 * 2)()
        self.lib.zdev_spec2d_export_counts(self.species[k]["handle"], c)
        return int(c[0]), int(c[1])

    def export_buffer(self, k, side, n):
        """torch view (int32 words, 7 per record) of the first n exported records"""
        ptr = self.lib.zdev_spec2d_export_ptr(self.species[k]["handle"], side)
        if n == 0 or not ptr:
            return self.new_part_buffer(0)
        return _tensor_from_ptr(self.torch, ptr, n * 7, self.torch.int32, self.device)

    def import_particles(self, k, buf):
        n = buf.numel() // 7
        if n:
            self.lib.zdev_spec2d_append_device(self.species[k]["handle"], buf.data_ptr(), n)

    def append_host_particles(self, k, part_aos):
        a = np.ascontiguousarray(part_aos)
        if len(a):
            self.lib.zdev_spec2d_append(self.species[k]["handle"], a.ctypes.data, len(a))

    def fetch(self, k):
        e, n = C.c_double(), C.c_int64()
        self.lib.zdev_spec2d_fetch(self.species[k]["handle"], C.byref(e), C.byref(n))
        return e.value, n.value

    def download_particles(self, k):
        from .abi_em2d import PART_DTYPE
        n = self.lib.zdev_spec2d_np(self.species[k]["handle"])
        out = np.zeros(max(n, 1), dtype=PART_DTYPE)
        n = self.lib.zdev_spec2d_download(self.species[k]["handle"], out.ctypes.data, len(out))
        return out[:n]

    # --- grids -------------------------------------------------------------------------------
    def upload_grid(self, which, arr):
        a = np.ascontiguousarray(arr, dtype=np.float32)
        self.lib.zdev_grid2d_upload(self.grid, which, a.ctypes.data)

    def download_grid(self, which):
        out = np.empty((self.g.ny + 3, self.g.nxl + 3, 3), dtype=np.float32)
        self.lib.zdev_grid2d_download(self.grid, which, out.ctypes.data)
        return out

    def pack(self, whichs, i0, ncols, j0, nrows, out):
        n = ncols * nrows * 3
        for k, w in enumerate(whichs):
            self.lib.zdev_grid2d_pack_cols(self.grid, w, i0, ncols, j0, nrows, out.data_ptr() + 4 * n * k)

    def unpack(self, whichs, i0, ncols, j0, nrows, buf, add):
        n = ncols * nrows * 3
        for k, w in enumerate(whichs):
            self.lib.zdev_grid2d_unpack_cols(self.grid, w, i0, ncols, j0, nrows, buf.data_ptr() + 4 * n * k, int(add))

    def current_zero(self):
        self.lib.zdev_current_zero(self.grid)

    def current_fold_x_local(self):
        self.lib.zdev_current_update_gc(self.grid, 0)      # x and y folds, single periodic slab

    def current_fold_y(self):
        self.lib.zdev_current_fold_y(self.grid)

    def smooth_plan(self):
        dirs, sa, sb = (C.c_int * 64)(), (C.c_float * 64)(), (C.c_float * 64)()
        n = self.lib.zdev_smooth_plan(*self.smooth, dirs, sa, sb)
        return [(dirs[i], sa[i], sb[i]) for i in range(n)]

    def smooth_pass(self, d, sa, sb, keep_x_guards):
        self.lib.zdev_smooth_pass(self.grid, d, sa, sb, int(keep_x_guards))

    def yee_b(self):
        dth = self.dt / np.float32(2.0)
        self.lib.zdev_yee_b(self.grid, float(dth / self.dx), float(dth / self.dy))

    def yee_e(self):
        self.lib.zdev_yee_e(self.grid, self.grid, float(self.dt / self.dx), float(self.dt / self.dy), float(self.dt))

    def emf_gc(self, skip_x):
        self.lib.zdev_emf_update_gc(self.grid, int(skip_x))

    def emf_shift(self, zero_right):
        self.lib.zdev_emf_shift(self.grid, int(zero_right))

    def emf_part_fld(self):
        self.lib.zdev_emf_update_part_fld(self.grid)

    def energy_sums(self):
        e = (C.c_double * 6)()
        self.lib.zdev_emf_energy(self.grid, e)
        return np.array(e[:])

    def sync(self):
        self.lib.zdev_sync()

    def destroy(self):
        for sp in self.species:
            self.lib.zdev_spec2d_destroy(sp["handle"])
        self.lib.zdev_grid2d_destroy(self.grid)


def share_stream_with_torch(lib):
    """Run the library and torch (buffers, NCCL collectives) on ONE explicit CUDA stream so that
    kernels, pack/unpack and send/recv are ordered without host synchronisation.  The legacy default
    stream cannot be used for this (handle 0 means 'library stream' to zdev_set_stream and the library
    stream is non-blocking), so a fresh torch stream is made current.  Returns it (keep a reference)."""
    import torch
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    lib.zdev_set_stream(stream.cuda_stream)
    return stream


def _tensor_from_ptr(torch, ptr, n, dtype, device):
    """zero-copy torch view of device memory owned by the library"""
    class _Holder:
        pass
    h = _Holder()
    typestr = {torch.int32: "<i4", torch.float32: "<f4", torch.int64: "<i8"}[dtype]
    h.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(h, device=device)


# ------------------------------------------------------------------------------------------ the step

def window_due(it_plus_1, dt, dx, n_move):
    """the reference's float test (em2d/particles.c:621, emf.c:650)"""
    return bool(np.float32(it_plus_1) * np.float32(dt) > np.float32(dx) * np.float32(n_move + 1))


def slab_step(slab, comm, inject_column=None):
    """One sim_iter (em2d/simulation.c:45-56) of one slab.  Written as a generator-free straight line;
    with LoopbackComm several slabs are stepped phase by phase by `step_all`."""
    for _ in slab_step_phases(slab, comm, inject_column):
        pass


def slab_step_phases(slab, comm, inject_column=None):
    """The same step cut at every exchange, so that several slabs of one process can be interleaved
    (each `yield` is a point where all slabs must have posted their messages)."""
    g = slab.g
    nsp = len(slab.species)
    ny = g.ny
    rows_all = (-1, ny + 3)               # j0, nrows: every row incl. guards
    interior = g.left is not None or g.right is not None

    # ---- particles -----------------------------------------------------------------------------
    slab.current_zero()
    shifts = []
    for k in range(nsp):
        sp = slab.species[k]
        shift = g.window and window_due(sp["iter"] + 1, slab.dt, slab.dx, sp["n_move"])
        slab.push(k, shift)
        sp["iter"] += 1
        if shift:
            sp["n_move"] += 1
        shifts.append(shift)
    if interior:
        counts = [slab.export_counts(k) for k in range(nsp)]
        flat = [c for pair in counts for c in pair]
        mine = slab.new_counts(flat)
        from_left, from_right = slab.new_counts(), slab.new_counts()
        # counts: I tell my left neighbour what goes left (even slots), my right neighbour what goes right
        comm.exchange(mine, mine, from_left, from_right)
        yield
        fl, fr = from_left.tolist(), from_right.tolist()
        for k in range(nsp):
            nl, nr = counts[k]
            send_l = slab.export_buffer(k, 0, nl) if g.left is not None else None
            send_r = slab.export_buffer(k, 1, nr) if g.right is not None else None
            # my left neighbour sends me what it exported to ITS right (odd slot) and vice versa
            recv_l = slab.new_part_buffer(fl[2 * k + 1]) if g.left is not None else None
            recv_r = slab.new_part_buffer(fr[2 * k]) if g.right is not None else None
            comm.exchange(send_l, send_r, recv_l, recv_r)
            yield
            if recv_l is not None:
                slab.import_particles(k, recv_l)
            if recv_r is not None:
                slab.import_particles(k, recv_r)
    for k in range(nsp):
        if shifts[k] and g.is_last and inject_column is not None:
            slab.append_host_particles(k, inject_column(k, slab.species[k]["n_move"]))

    # ---- current: guard fold + smoothing ----------------------------------------------------------
    if interior:
        j0, nr = rows_all
        sl = slab.new_grid_buffer(1, 3, nr) if g.left is not None else None
        sr = slab.new_grid_buffer(1, 3, nr) if g.right is not None else None
        rl = slab.new_grid_buffer(1, 3, nr) if g.left is not None else None
        rr = slab.new_grid_buffer(1, 3, nr) if g.right is not None else None
        if sl is not None:
            slab.pack([J], -1, 3, j0, nr, sl)
        if sr is not None:
            slab.pack([J], g.nxl - 1, 3, j0, nr, sr)
        comm.exchange(sl, sr, rl, rr)
        yield
        if rl is not None:
            slab.unpack([J], -1, 3, j0, nr, rl, add=True)
        if rr is not None:
            slab.unpack([J], g.nxl - 1, 3, j0, nr, rr, add=True)
        slab.current_fold_y()
    elif g.window:
        slab.current_fold_y()
    else:
        slab.current_fold_x_local()
    y_passes = False
    for d, sa, sb in slab.smooth_plan():
        # physical edges follow the reference: refreshed locally when periodic, untouched with a window;
        # interior edges are refreshed from the neighbour after the pass
        slab.smooth_pass(d, sa, sb, keep_x_guards=(g.window or interior))
        if d == 0 and interior:
            for _ in _halo_refresh(slab, comm, [J], 0, ny):
                yield
        y_passes = y_passes or d == 1
    if y_passes and interior:
        # kernel_y leaves the x guard columns un-filtered (em2d/current.c:382-411).  Across the box
        # boundary that is what the reference feeds to yee_e, so the wrap-around edge keeps it; guards
        # that mirror interior cells of a neighbour slab must see the filtered values
        for _ in _halo_refresh(slab, comm, [J], -1, ny + 3, skip_wrap=True):
            yield

    # ---- fields ---------------------------------------------------------------------------------
    slab.yee_b()
    slab.yee_e()
    slab.yee_b()
    if interior:
        for _ in _halo_refresh(slab, comm, [E, B], -1, ny + 3):
            yield
        slab.emf_gc(skip_x=True)
    else:
        slab.emf_gc(skip_x=g.window)
    slab.emf_part_fld()
    shift = g.window and window_due(slab.iter + 1, slab.dt, slab.dx, slab.n_move)
    slab.iter += 1
    if shift:
        slab.emf_shift(zero_right=(g.right is None))
        slab.n_move += 1
        if interior:
            for _ in _halo_refresh(slab, comm, [E, B], -1, ny + 3):
                yield


def _halo_refresh(slab, comm, whichs, j0, nrows, skip_wrap=False):
    """guards <- neighbour interior: (-1) <- left's (nxl-1); (nxl, nxl+1) <- right's (0, 1)"""
    g = slab.g
    n = len(whichs)
    do_l = g.left is not None and not (skip_wrap and g.wrap_left)
    do_r = g.right is not None and not (skip_wrap and g.wrap_right)
    sl = slab.new_grid_buffer(n, 2, nrows) if do_l else None
    sr = slab.new_grid_buffer(n, 1, nrows) if do_r else None
    rl = slab.new_grid_buffer(n, 1, nrows) if do_l else None
    rr = slab.new_grid_buffer(n, 2, nrows) if do_r else None
    if sl is not None:
        slab.pack(whichs, 0, 2, j0, nrows, sl)
    if sr is not None:
        slab.pack(whichs, g.nxl - 1, 1, j0, nrows, sr)
    comm.exchange(sl, sr, rl, rr)
    yield
    if rl is not None:
        slab.unpack(whichs, -1, 1, j0, nrows, rl, add=False)
    if rr is not None:
        slab.unpack(whichs, g.nxl, 2, j0, nrows, rr, add=False)


def step_all(slabs, comms, inject_column=None):
    """advance several slabs of one process in lock step (LoopbackComm)"""
    gens = [slab_step_phases(s, c, inject_column) for s, c in zip(slabs, comms)]
    alive = True
    while alive:
        alive = False
        for gen in gens:
            try:
                next(gen)
                alive = True
            except StopIteration:
                pass


# ------------------------------------------------------------------------------------------ helpers

def split_particles(part_aos, geom):
    """particles of the global population that live in this slab, in local coordinates"""
    m = (part_aos["ix"] >= geom.x0) & (part_aos["ix"] < geom.x0 + geom.nxl)
    out = part_aos[m].copy()
    out["ix"] -= geom.x0
    return out


def split_grid(arr, geom):
    """columns of a global (ny+3, nx+3, 3) buffer that form this slab's local buffer incl. its halo"""
    return np.ascontiguousarray(arr[:, geom.x0: geom.x0 + geom.nxl + 3, :])


def join_grids(local_arrays, nranks):
    """interior columns of the slabs' local buffers -> global interior (ny, nx, 3)"""
    return np.concatenate([a[1:-2, 1:-2, :] for a in local_arrays], axis=1)


class HostColumnInjector:
    """Moving-window injection for the last slab: runs the library's host injector (the reference's
    sequential random stream, em2d/particles.c:476-492, 619-638) on the GLOBAL species description for
    the right-most cell column and returns the new particles in the slab's local coordinates."""

    def __init__(self, lib, species_array, geom):
        from . import abi_em2d as A
        self.A, self.lib, self.species, self.g = A, lib, species_array, geom
        fn = lib.spec_inject_into
        fn.restype = None
        fn.argtypes = [C.POINTER(A.Species), C.c_void_p, C.POINTER(C.POINTER(A.Part)), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        self.libc = C.CDLL(None)
        self.libc.free.argtypes = [C.c_void_p]

    def __call__(self, k, n_move):
        A, g = self.A, self.g
        sp = self.species[k]
        sp.n_move = n_move
        rng = ((C.c_int * 2) * 2)((C.c_int * 2)(g.nx_global - 1, g.nx_global - 1), (C.c_int * 2)(0, g.ny - 1))
        buf, n, nmax = C.POINTER(A.Part)(), C.c_int(0), C.c_int(0)
        self.lib.spec_inject_into(C.byref(sp), rng, C.byref(buf), C.byref(n), C.byref(nmax))
        out = np.zeros(n.value, dtype=A.PART_DTYPE)
        if n.value:
            raw = np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_uint8)), shape=(n.value * 28,))
            out[:] = raw.view(A.PART_DTYPE)
            out["ix"] -= g.x0
        if buf:
            self.libc.free(buf)
        return out
