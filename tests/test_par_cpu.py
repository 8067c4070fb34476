"""Host side of the slab decomposition without a GPU: the job segment of zb_par.h (ranks, barrier, sums, gathers,
the shared scratch area) with 2 and 4 real processes, and the slab geometry the API layer derives from it."""
import os
import subprocess
import sys

import pytest

from tests import helpers as H

WORKER = r'''
import ctypes as C, os, sys
import numpy as np
lib = C.CDLL(sys.argv[1])
n = lib.zb_par_init(); r = lib.zb_par_rank()
assert n == int(os.environ["ZPIC_NRANKS"]) and r == int(os.environ["ZPIC_RANK"])
for it in range(50):                                   # many rounds: the barrier's sense reversal, buffer reuse
    v = (C.c_double * 3)(r + 1.0, 10.0 * (r + 1) + it, 0.5)
    lib.zb_par_allreduce_sum_d(v, 3)
    assert list(v) == [n * (n + 1) / 2, 10.0 * n * (n + 1) / 2 + n * it, 0.5 * n], (it, list(v))
    w = (C.c_longlong * 2)(r, 7)
    lib.zb_par_allreduce_sum_ll(w, 2)
    assert list(w) == [n * (n - 1) // 2, 7 * n]
    mine = (C.c_int * 4)(r, it, r * r, -1)
    allv = (C.c_int * (4 * n))()
    lib.zb_par_allgather(mine, C.c_size_t(16), allv)
    assert [allv[4 * k] for k in range(n)] == list(range(n)) and all(allv[4 * k + 1] == it for k in range(n))
lib.zb_par_scratch.restype = C.c_void_p
for size in (1 << 12, 1 << 20, 1 << 16):               # grows, then a smaller request reuses the larger area
    p = lib.zb_par_scratch(C.c_size_t(size))
    a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int)), shape=(size // 4,))
    a[r::n] = r + 1
    lib.zb_par_barrier()
    assert all((a[k::n] == k + 1).all() for k in range(n))
    lib.zb_par_barrier()
f = np.full(1000, r + 1.0, dtype=np.float32)
lib.zb_par_allreduce_sum_f(f.ctypes.data_as(C.POINTER(C.c_float)), C.c_size_t(1000))
assert (f == n * (n + 1) / 2).all()
# the slab geometry of the API layer (zb_state.h): periodic ring and moving-window chain
class Slab(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("on", "rank", "nranks", "nxl", "x0", "left", "right", "wrap_left", "wrap_right", "is_last")]
lib.zb_slab_make.restype = Slab
g = lib.zb_slab_make(64 * n, 0)
assert (g.on, g.nxl, g.x0, g.left, g.right) == (1, 64, 64 * r, (r - 1) % n, (r + 1) % n)
assert (g.wrap_left, g.wrap_right, g.is_last) == (int(r == 0), int(r == n - 1), int(r == n - 1))
g = lib.zb_slab_make(64 * n, 1)
assert (g.left, g.right) == (r - 1 if r > 0 else -1, r + 1 if r < n - 1 else -1) and not g.wrap_left and not g.wrap_right
print("ok", r)
'''


@pytest.mark.parametrize("nranks", [2, 4])
def test_job_segment_collectives(nranks, tmp_path):
    from zpic_b200 import build
    lib = build.lib_path("em2d")
    procs = []
    for r in range(nranks):
        env = dict(os.environ, ZPIC_RANK=str(r), ZPIC_NRANKS=str(nranks), ZPIC_JOB="cpu%d_%d" % (os.getpid(), nranks))
        env.pop("ZPIC_SLABS", None)
        procs.append(subprocess.Popen([sys.executable, "-c", WORKER, lib], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=120)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert sorted(o.strip().splitlines()[-1] for o in outs) == ["ok %d" % r for r in range(nranks)]


def test_single_process_is_one_slab():
    import ctypes as C
    from zpic_b200 import load
    lib = load("em2d")
    env_keep = {k: os.environ.pop(k, None) for k in ("ZPIC_NRANKS", "WORLD_SIZE")}
    try:
        assert lib.zb_par_init() == 1 and lib.zb_par_rank() == 0
    finally:
        for k, v in env_keep.items():
            if v is not None:
                os.environ[k] = v
