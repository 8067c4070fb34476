"""Config 4 (BASELINE.json): em2d Kelvin-Helmholtz shear flow, 2 species x 32 ppc each filling half of the box
(counter-streaming along x), binomial current smoothing in x and y, periodic, slab-decomposed along x, one
process per GPU.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/kh_slabs.py [n steps warmup]
Device-side initialisation (counter-based generator, rows of the two half-box species injected by
zdev_spec2d_inject_band); tiles outside a species' half start with the minimum capacity and grow on demand as
the shear layer rolls up.  Prints one JSON line (rank 0)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zpic_b200 import abi_em2d as A
from zpic_b200 import load
from zpic_b200 import parallel as P

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
warmup = int(sys.argv[3]) if len(sys.argv) > 3 else 3
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
import torch
import torch.distributed as dist
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = load("em2d")
assert lib.zdev_init(local) == 0
os.environ.setdefault("ZPIC_TILE_SLACK", "1.25")
stream = P.share_stream_with_torch(lib)
DT, CELL, PPC = 0.07, 0.1, (8, 4)
npc = PPC[0] * PPC[1]
geom = P.Geometry(n, n, world, rank, moving_window=False)
cfg = [dict(m_q=-1.0, q=-1.0 / npc, ppc=PPC), dict(m_q=-1.0, q=-1.0 / npc, ppc=PPC)]
slab = P.CudaSlab(lib, geom, DT, CELL, CELL, cfg, (A.BINOMIAL, A.BINOMIAL, 1, 1))
slab.inject_band(0, PPC, (0.2, 0.0, 0.0), (0.01, 0.01, 0.01), 4242 + 7919 * rank, 0, n // 2)
slab.inject_band(1, PPC, (-0.2, 0.0, 0.0), (0.01, 0.01, 0.01), 4343 + 7919 * rank, n // 2, n)
comm = P.TorchComm(geom) if world > 1 else P.LoopbackComm(geom, P.LoopbackComm.Hub(1))
for _ in range(warmup):
    P.slab_step(slab, comm)
lib.zdev_sync()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    P.slab_step(slab, comm)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
cnt = torch.tensor([sum(slab.fetch(k)[1] for k in range(2))], device="cuda", dtype=torch.int64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(cnt)
expect = n * n * npc
assert int(cnt.item()) == expect, "particles were lost: %d of %d" % (int(cnt.item()), expect)
if rank == 0:
    t = float(ms.item()) / steps
    print(json.dumps({"workload": "em2d Kelvin-Helmholtz %dx%d cells, 2 species x %d ppc (half box each), binomial smoothing x,y level 1, periodic (BASELINE configs[3])" % (n, n, npc),
                      "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": t, "particles": expect,
                      "pushes_per_s": expect / (t * 1e-3), "cell_updates_per_s": n * n / (t * 1e-3)}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
