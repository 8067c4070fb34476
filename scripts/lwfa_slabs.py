"""Config 3 (BASELINE.json): em2d LWFA with laser, moving window and compensated smoothing, slab-decomposed
along x, one process per GPU.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/lwfa_slabs.py [nx ny steps]
Every rank builds the global deck on the host through the C API (laser launch in libm double precision,
reference random stream), keeps its slab, then steps with halo / particle exchange over NCCL; the last rank
injects the window's new column with the host injector.  Prints one JSON line (rank 0)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import helpers as H
from zpic_b200 import abi_em2d as A
from zpic_b200 import load
from zpic_b200 import parallel as P

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
ny = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 100
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
import torch
import torch.distributed as dist
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = load("em2d")
assert lib.zdev_init(local) == 0
stream = P.share_stream_with_torch(lib)
box = (nx * 0.01, ny * 0.05)                               # dx = (0.01, 0.05) as input/lwfa-large.c
t0 = time.time()
host = H.lwfa(lib, nx=(nx, ny), box=box, dt=0.009, ppc=(4, 4), start=0.5, laser_start=box[0] - 3.0, a0=3.0)
t_init = time.time() - t0
geom = P.Geometry(nx, ny, world, rank, moving_window=True)
cfg = [dict(m_q=host.species[0].m_q, q=host.species[0].q, ppc=(4, 4))]
slab = P.CudaSlab(lib, geom, host.sim.dt, host.sim.emf.dx[0], host.sim.emf.dx[1], cfg, (A.COMPENSATED, 0, 4, 0))
slab.upload_grid(P.E, P.split_grid(host.E(), geom))
slab.upload_grid(P.B, P.split_grid(host.B(), geom))
slab.upload_particles(0, P.split_particles(host.parts(0), geom))
np_global = int(host.species[0].np)
comm = P.TorchComm(geom) if world > 1 else P.LoopbackComm(geom, P.LoopbackComm.Hub(1))
inject = P.HostColumnInjector(lib, host.species, geom) if rank == world - 1 else None
for _ in range(5):
    P.slab_step(slab, comm, inject)
lib.zdev_sync()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    P.slab_step(slab, comm, inject)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
cnt = torch.tensor([slab.fetch(0)[1]], device="cuda", dtype=torch.int64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(cnt)
if rank == 0:
    t = float(ms.item()) / steps
    print(json.dumps({"workload": "em2d LWFA %dx%d cells, 16 ppc, laser a0 3, moving window, compensated smoothing level 4 (BASELINE configs[2])" % (nx, ny),
                      "n_gpus": world, "steps": steps, "ms_per_step": t, "particles": int(cnt.item()), "particles_initial": np_global,
                      "pushes_per_s": float(cnt.item()) / (t * 1e-3), "cell_updates_per_s": nx * ny / (t * 1e-3),
                      "n_move": int(slab.n_move), "host_init_s": t_init}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
