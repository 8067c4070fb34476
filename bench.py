#!/usr/bin/env python
"""bench.py - particle-pushes/s of the em2d time step on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # the CUDA implementation
    python bench.py --impl reference --gpus N --steps K ...  # the reference C code on the host cores

A "step" is one full sim_iter of the reference API (em2d/simulation.c:45-56): current_zero,
spec_advance for every species (interpolate + Boris push + current deposit + boundaries +
re-binning), current_update, emf_advance.  `value` = particle pushes per second of the whole
job, counted like the reference counts them (np per species per step, em2d/particles.c:1267),
with the state resident in HBM.  Workload at N=1: configs[1] of BASELINE.json - the Weibel deck
scaled to 4096 x 4096 cells, 2 species x 64 particles per cell (2^31 particles), periodic,
re-binned every step; per-GPU work is kept fixed for N>1 (weak scaling: the box grows along x).

Extra keys (see the task contract): `roofline` for the dominant kernel (k_push2d) from CUDA
events recorded around its launches on the library stream; `cpu_baseline` = the unmodified
reference (oracle/_ref, -Ofast as shipped) timed on one host core over a bounded sample;
`cells` = cell-updates/s of the grid half of the step (current_update + emf_advance) timed alone;
`e2e` = the same metric through the public C API (sim_new / sim_iter) with HOST buffers, i.e.
host->device upload of particles+fields and device->host download inside every timed step.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = "particle-pushes/sec (push+deposit)"
UNIT = "pushes/s"
BYTES_PER_PUSH = 56.0          # algorithmic: 28 B read + 28 B write per particle
BYTES_PER_CELL = 72.0          # algorithmic, grid half of a step: E,B read + write (48 B), J zeroed and read (24 B)
CELL = 0.1                     # dx of the shipped Weibel deck (em2d/input/weibel.c:17-18)
DT = 0.07


def peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region"""

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu = gpu
        self.samples = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.samples.append([t.strip() for t in line.split(",")])
                if self.stop_flag:
                    break
        except OSError:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        # samples under load only: the upper half (idle samples before/after the region are low)
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": load[len(load) // 2] if load else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# decks through the C API (the same calls for our library and for the reference build)

def build_weibel(lib, A, nx, ny, ppc, seed=(12345, 67890)):
    """reference em2d/input/weibel.c:13-40 with adjustable size / ppc"""
    libc = C.CDLL(None)
    libc.calloc.restype = C.c_void_p
    libc.calloc.argtypes = [C.c_size_t, C.c_size_t]
    lib.set_rand_seed(*seed)
    cnx = (C.c_int * 2)(nx, ny)
    box = (C.c_float * 2)(nx * CELL, ny * CELL)
    species = C.cast(libc.calloc(2, C.sizeof(A.Species)), C.POINTER(A.Species))
    cppc = (C.c_int * 2)(*ppc)
    uth = (C.c_float * 3)(0.1, 0.1, 0.1)
    for k, (name, m_q, uz) in enumerate(((b"electrons", -1.0, 0.6), (b"positrons", 1.0, -0.6))):
        ufl = (C.c_float * 3)(0.0, 0.0, uz)
        lib.spec_new(C.byref(species[k]), name, m_q, cppc, ufl, uth, cnx, box, DT, None)
    sim = A.Simulation()
    lib.sim_new(C.byref(sim), cnx, box, DT, 35.0, 10, species, 2)
    return sim, species, (cnx, box)


def fit_grid(lib, n, ppc_total):
    """shrink the square grid until two species fit in free device memory (keeps ppc)"""
    free_b, total_b = C.c_size_t(), C.c_size_t()
    lib.zdev_mem_info(C.byref(free_b), C.byref(total_b))
    while n > 256:
        need = 2 * (n * n * ppc_total) * 1.25 * (2 * 26 + 28 / 8.0 + 28 / 32.0) + 5 * (n + 3) ** 2 * 12      # A/B records + keys, migrants, overflow list; grids
        if need < 0.92 * free_b.value:
            break
        n //= 2
    return n, free_b.value


def time_grid_term(lib, sim, n, reps=20):
    """cell-updates/s of the grid half of a step alone - current_zero, current_update (guard fold; the Weibel
    deck does not smooth) and emf_advance (yee_b, yee_e, yee_b, guard refresh) through the public C API
    (em2d/current.c:98-183, em2d/emf.c:688-716), CUDA events on the library stream.  The grids (3 x 202 MB
    at 4096^2) exceed the L2, so consecutive repetitions do not feed each other from cache."""
    def grid_half():
        lib.current_zero(C.byref(sim.current))
        lib.current_update(C.byref(sim.current))
        lib.emf_advance(C.byref(sim.emf), C.byref(sim.current))
    for _ in range(3):
        grid_half()
    e0, e1 = lib.zdev_event_create(), lib.zdev_event_create()
    lib.zdev_sync()
    n0 = lib.zdev_launch_count()
    lib.zdev_event_record(e0)
    for _ in range(reps):
        grid_half()
    lib.zdev_event_record(e1)
    ms = lib.zdev_event_elapsed_ms(e0, e1) / reps
    peak, peak_src = peaks()
    gbs = BYTES_PER_CELL * n * n / (ms * 1e-3) / 1e9
    return {"value": n * n / (ms * 1e-3), "unit": "cell-updates/s", "ms_per_step": ms,
            "kernels_per_step": int((lib.zdev_launch_count() - n0) // reps),
            "algorithmic_bytes_per_cell": BYTES_PER_CELL, "achieved": gbs, "peak": peak, "frac": gbs / peak,
            "what": "current_zero + current_update + emf_advance alone on the %dx%d grid (the grid term of the "
                    "step; with the particles it is %s of a step)" % (n, n, "%.1f %%")}


def slab_parity(lib, A, rank, world):
    """N > 1, after the timed region: a small Weibel box (64 cells per slab along x, 64 along y, 2 x 16 ppc, host
    initialisation = the reference random stream) advanced 10 steps by the N slabs through the C API, gathered on
    rank 0 and compared with the unmodified reference (oracle/_ref) running the whole box on the host: particle
    counts equal, cells bit-identical after step 1, E/B/J within 1e-5 (relative L2)."""
    lib.zpic_b200_set_option(b"device_init", 0)
    lib.zpic_b200_set_option(b"lazy", 0)
    nx, ny, ppc = 64 * world, 64, (4, 4)
    sim, species, _ = build_weibel(lib, A, nx, ny, ppc)
    ref = None
    if rank == 0:
        path = os.path.join(REPO, "oracle", "_ref", "libzpic_ref_em2d.so")
        if os.path.exists(path):
            ref = A.declare(C.CDLL(path))
            rsim, rspecies, _ = build_weibel(ref, A, nx, ny, ppc)
    res = {"deck": "Weibel %dx%d, 2 x 16 ppc, %d slabs, 10 steps vs the reference on the host" % (nx, ny, world)}
    ok = True
    for cp in (1, 10):
        for _ in range(cp - sim.emf.iter):
            lib.sim_iter(C.byref(sim))
        lib.zpic_b200_sync_host(C.byref(sim))                 # collective: gathers the slabs into every mirror
        if ref is None:
            continue
        for _ in range(cp - rsim.emf.iter):
            ref.sim_iter(C.byref(rsim))
        for name, a, b in (("E", sim.emf.E_buf, rsim.emf.E_buf), ("B", sim.emf.B_buf, rsim.emf.B_buf),
                           ("J", sim.current.J_buf, rsim.current.J_buf)):
            x, y = A.grid_view(a, nx, ny).astype(np.float64), A.grid_view(b, nx, ny).astype(np.float64)
            err = float(np.sqrt(((x - y) ** 2).sum() / max((y ** 2).sum(), 1e-300)))
            res["%s_rel_l2_step%d" % (name, cp)] = err
            ok = ok and err < 1e-5
        for k in range(2):
            pa, pb = A.part_view(species[k]), A.part_view(rspecies[k])
            ok = ok and len(pa) == len(pb)
            if cp == 1 and len(pa) == len(pb):
                ca = np.sort(pa["ix"].astype(np.int64) * ny + pa["iy"])
                cb = np.sort(pb["ix"].astype(np.int64) * ny + pb["iy"])
                same = bool(np.array_equal(ca, cb))
                res["cells_identical_step1_species%d" % k] = same
                ok = ok and same
    lib.sim_delete(C.byref(sim))
    if ref is not None:
        ref.sim_delete(C.byref(rsim))
        res["ok"] = bool(ok)
        sys.stderr.write("slab_parity: %s\n" % ("ok" if ok else "FAILED"))
    else:
        res["ok"] = None
        res["note"] = "oracle/_ref not present on this box"
    return res if rank == 0 else None


class Ranks:
    """Barrier and max over the ranks of the job.  Launched by torchrun: torch.distributed over NCCL (the task contract).
    The children the default run starts for the other configurations ($ZPIC_BENCH_CHILD) use the library's own job
    segment instead (zb_par_barrier / zb_par_allgather, csrc/host/common/zb_par.h): a second NCCL job next to the
    parent's would need a rendezvous store of its own."""

    def __init__(self, lib, world, local):
        self.lib, self.world, self.torch, self.dist = lib, world, None, None
        if world > 1 and not os.environ.get("ZPIC_BENCH_CHILD"):
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            os.environ.setdefault("ZPIC_JOB", "bench" + os.environ.get("MASTER_PORT", "0"))
            self.torch, self.dist = torch, dist
        elif world > 1:
            lib.zb_par_allgather.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
            lib.zb_par_allgather.restype = None
            lib.zb_par_init()

    def barrier(self):
        self.lib.zdev_sync()
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize()
        elif self.world > 1:
            self.lib.zb_par_barrier()

    def max(self, x):
        if self.dist is not None:
            t = self.torch.tensor([x], device="cuda", dtype=self.torch.float64)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            return float(t.item())
        if self.world > 1:
            mine, every = C.c_double(x), (C.c_double * self.world)()
            self.lib.zb_par_allgather(C.byref(mine), 8, every)
            return max(every)
        return x

    def close(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()


def other_configs(rank, world, args):
    """The other BASELINE configurations, measured by the default run too (so that the driver's record holds them):
    N = 1: configs[4] (em1d two-stream, 2^31 particles) and configs[2] (LWFA) on the one GPU; N > 1: configs[2] and
    configs[3] (Kelvin-Helmholtz; its 2^31-particle box needs two GPUs) slab-decomposed over the N ranks.  Every rank
    starts `bench.py --workload ...` as a child with its own rank environment - a child that fails or hangs costs
    its timeout, not the line of the main configuration.  Returns {name: summary} on rank 0."""
    if args.no_extras:
        return None
    jobs = [("em1d", ["--workload", "em1d", "--steps", "10", "--warmup", "3"], 240),
            ("lwfa", ["--workload", "lwfa", "--steps", "200", "--warmup", "5"], 240)]
    if world > 1:
        jobs = [jobs[1], ("kh", ["--workload", "kh", "--steps", "20", "--warmup", "3"], 300)]
    out = {}
    for name, extra, limit in jobs:
        env = dict(os.environ, ZPIC_BENCH_CHILD="1", ZPIC_JOB="bx%s%s" % (os.environ.get("MASTER_PORT", "0"), name))
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--gpus", str(world)] + extra, env=env,
                               stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=limit)
            lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            if rank == 0:
                if r.returncode == 0 and lines:
                    d = json.loads(lines[-1])
                    out[name] = {"value": d["value"], "unit": d["unit"], "ms_per_step": d["ms_per_step"], "steps": d["steps"],
                                 "n_gpus": d["n_gpus"], "scaling": d["scaling"], "workload": d["config"]["workload"],
                                 "roofline_frac": d["roofline"]["frac"], "roofline_what": d["roofline"].get("kernel"),
                                 "cell_updates_per_s": (d.get("cells") or {}).get("value"), "wall_s": round(time.time() - t0, 1)}
                else:
                    out[name] = {"error": "exit code %d: %s" % (r.returncode, r.stderr.strip()[-300:])}
        except subprocess.TimeoutExpired:
            if rank == 0:
                out[name] = {"error": "no result within %d s" % limit}
    return out if rank == 0 else None


def run_ours(args):
    from zpic_b200 import abi_em2d as A
    from zpic_b200 import load
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        os.environ.setdefault("ZPIC_JOB", "bench" + os.environ.get("MASTER_PORT", "0"))
    lib = load("em2d")
    if lib.zdev_init(local) != 0:
        raise SystemExit("bench.py: no CUDA device - the CUDA path is the only path")
    K, W = args.steps, max(args.warmup, 0)
    ppc = (args.ppc, args.ppc)
    os.environ.setdefault("ZPIC_TILE_SLACK", "1.25")
    sampler = ClockSampler(local)

    # ---- state resident in HBM (device-side initialisation), fully asynchronous stepping.  N > 1: the SAME calls -
    #      sim_new / sim_iter of the reference API on a box that grows along x with N (weak scaling); the library
    #      reads RANK / WORLD_SIZE, keeps one slab per process / GPU and exchanges guard cells and particles GPU to
    #      GPU inside spec_advance / current_update / emf_advance (csrc/dev/zdev_slab.cuh).  torch.distributed
    #      (NCCL) only carries the barrier and the max-over-ranks of the timing.
    n, free_b = fit_grid(lib, args.n, args.ppc * args.ppc)
    # initial state: at N = 1 the reference's own (its random stream continued on the device, zdev_refrng.cu: the
    # particles are those spec_new of the reference creates for this deck and seed); at N > 1 the box is not square
    # and a warm plasma there takes the reference's cell-mixing means: counter-based generator
    init_mode = args.init if args.init else (2 if world == 1 else 1)
    lib.zpic_b200_set_option(b"device_init", init_mode)
    lib.zpic_b200_set_option(b"lazy", 1)
    lib.zpic_b200_set_option(b"coherent", 0)
    lib.zdev_set_push_timing(1)
    t_init0 = time.perf_counter()
    sim, species, _ = build_weibel(lib, A, n * world, n, ppc)
    lib.sim_iter(C.byref(sim))
    lib.zdev_sync()
    t_init = time.perf_counter() - t_init0
    np_total = 2 * n * n * args.ppc * args.ppc            # per GPU
    for _ in range(max(W, 1)):
        lib.sim_iter(C.byref(sim))
    lib.zdev_sync()
    from zpic_b200._lib import spec_handle
    handles = [spec_handle(lib, C.byref(species[k])) for k in range(2)]
    for h in handles:
        lib.zdev_spec2d_push_timing(h, None, None, 1)
    launches0 = lib.zdev_launch_count()
    if rank == 0:
        sampler.start()
    e0, e1 = lib.zdev_event_create(), lib.zdev_event_create()
    lib.zdev_sync()
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()
    lib.zdev_event_record(e0)
    for _ in range(K):
        lib.sim_iter(C.byref(sim))
    lib.zdev_event_record(e1)
    ms = lib.zdev_event_elapsed_ms(e0, e1)
    lib.zdev_sync()
    if dist is not None:
        dist.barrier()
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.finish() if rank == 0 else None
    launches = lib.zdev_launch_count() - launches0
    push_ms, push_n = 0.0, 0
    for h in handles:
        t, c = C.c_double(), C.c_int64()
        lib.zdev_spec2d_push_timing(h, C.byref(t), C.byref(c), 1)
        push_ms += t.value
        push_n += c.value
    lib.zdev_set_push_timing(0)
    # sanity: the population did not leak (periodic box; over all slabs when decomposed)
    lib.zpic_b200_set_option(b"lazy", 0)
    en, cnt = C.c_double(), C.c_int64()
    lib.zdev_spec2d_fetch(handles[0], C.byref(en), C.byref(cnt))
    total = cnt.value
    if dist is not None:
        tt = torch.tensor([total], device="cuda", dtype=torch.int64)
        dist.all_reduce(tt)
        total = int(tt.item())
    assert total == world * n * n * args.ppc * args.ppc, "particles were lost: %d" % total
    value = world * np_total * K / (ms * 1e-3)
    cells = None
    parity = None
    if world == 1:
        cells = time_grid_term(lib, sim, n)
        cells["what"] = cells["what"] % (100.0 * cells["ms_per_step"] / (ms / K))
    lib.sim_delete(C.byref(sim))
    if world > 1 and not args.no_check:
        parity = slab_parity(lib, A, rank, world)

    out = None
    others = None
    if world > 1:
        others = other_configs(rank, world, args)         # (N = 1: after the host-buffer legs below)
    if rank == 0:
        peak, peak_src = peaks()
        per_launch = np_total / 2.0                       # particles per k_push2d launch (one species)
        avg_ms = push_ms / max(push_n, 1)
        achieved = BYTES_PER_PUSH * per_launch / (avg_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(REPO, "profiles", "push_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get("dram_bytes_per_particle", 0) * per_launch or None
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "em2d Weibel %dx%d cells%s, 2 species x %d ppc, periodic (BASELINE configs[1]%s)"
                                   % (n * world, n, " (%d per GPU along x)" % n if world > 1 else "", args.ppc * args.ppc,
                                      "" if n == 4096 else ", grid reduced to fit memory"),
                       "particles_per_gpu": np_total, "dt": DT, "dx": CELL,
                       "init": ("the reference's initial state: spec_new's particles for this deck and seed, generated on the device "
                                "from the reference's random stream (multiply-with-carry jump-ahead + scan over the Box-Muller "
                                "rejections)" if init_mode == 2 else "device-side counter-based thermal+fluid distribution"),
                       "init_s": round(t_init, 2),
                       "cache": "working set %.1f GB per step >> 126 MB L2, no flush needed" % (np_total * 48 / 1e9),
                       "decomposition": ("%d slabs along x, one process per GPU, through sim_new / sim_iter of the C API; J guard "
                                         "columns (add), E/B halos and migrating particles are written by the sending kernels "
                                         "straight into the neighbour GPU's memory over NVLink (CUDA IPC mailboxes + flags, no "
                                         "host synchronisation); NCCL carries only the bench's barrier / timing reduction" % world)
                       if world > 1 else "single GPU"},
            "roofline": {"bound": "hbm", "kernel": "k_push2d", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "traffic_source": "ncu constant (profiles/push_traffic.json: DRAM bytes per particle of the committed "
                                           "--set full capture of this kernel version) x particles per launch; not measured in this run",
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_particle": BYTES_PER_PUSH, "particles_per_launch": per_launch,
                         "avg_launch_ms": avg_ms, "launches_timed": push_n,
                         "kernel_share_of_step": push_ms / ms},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if world == 1:
            out["cells"] = cells
            out["e2e"] = run_e2e(lib, A, args, n)
            out["cpu_baseline"] = cpu_baseline(seconds=args.cpu_seconds, threads=1)
            others = other_configs(rank, world, args)
        else:
            out["slab_parity"] = parity
            out["e2e"] = {"value": value, "unit": UNIT, "h2d_bytes_per_step": 2 * 32 * world, "d2h_bytes_per_step": 2 * 48 * world,
                          "mode": "the timed loop IS the public C API (sim_iter on every rank); per step only kernel parameters "
                                  "go in and the 48-byte control block of every species comes out (read one step late); the "
                                  "host-buffer legs (host-initialised species, per-step diagnostics, report set) are measured at N=1"}
        if others is not None:
            out["other_configs"] = others
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return out


def _new_species(lib, A, n):
    libc = C.CDLL(None)
    libc.calloc.restype = C.c_void_p
    libc.calloc.argtypes = [C.c_size_t, C.c_size_t]
    return C.cast(libc.calloc(n, C.sizeof(A.Species)), C.POINTER(A.Species))


def build_lwfa(lib, A, nx, ny, a0=5.0):
    """BASELINE configs[2]: the reference's em2d/input/lwfa-large.c geometry (dx = 0.01, 0.05) scaled to nx x ny cells,
    4x4 particles per cell, plasma from x = 0.5 on (STEP: the pulse is inside the plasma from the first step), the gaussian
    laser of that deck (a0 = 5, omega0 10, W0 4, fwhm 2, input/lwfa-large.c:40-50) launched 3 c/wp from the right edge,
    moving window, compensated smoothing level 4 (input/lwfa.c:15-63 settings)"""
    lib.set_rand_seed(12345, 67890)
    cnx, box, dt = (C.c_int * 2)(nx, ny), (C.c_float * 2)(nx * 0.01, ny * 0.05), 0.009
    species = _new_species(lib, A, 1)
    dens = A.Density()
    dens.type, dens.start = A.STEP, 0.5
    lib.spec_new(C.byref(species[0]), b"electrons", -1.0, (C.c_int * 2)(4, 4), None, None, cnx, box, dt, C.byref(dens))
    sim = A.Simulation()
    lib.sim_new(C.byref(sim), cnx, box, dt, 1.0e9, 0, species, 1)
    laser = A.Laser()
    laser.type, laser.start, laser.fwhm, laser.a0, laser.omega0 = A.GAUSSIAN, box[0] - 3.0, 2.0, a0, 10.0
    laser.W0, laser.focus, laser.axis, laser.polarization = 4.0, box[0] - 10.0, box[1] / 2, np.pi / 2
    lib.sim_add_laser(C.byref(sim), C.byref(laser))
    lib.sim_set_moving_window(C.byref(sim))
    sm = A.Smooth(A.COMPENSATED, 0, 4, 0)
    lib.sim_set_smooth(C.byref(sim), C.byref(sm))
    return sim, species, 1, 16, dt, ()


def build_kh(lib, A, n):
    """BASELINE configs[3]: Kelvin-Helmholtz shear, n x n cells, two electron species that each fill one half of the box
    in y through a CUSTOM density (0/1 step along y) and drift along +x / -x at 0.2 c, 8x4 particles per cell,
    binomial smoothing level 1 along x and y, periodic"""
    lib.set_rand_seed(12345, 67890)
    cnx, box = (C.c_int * 2)(n, n), (C.c_float * 2)(n * CELL, n * CELL)
    species = _new_species(lib, A, 2)
    keep = []
    for k, (name, lower, ux) in enumerate(((b"lower", True, 0.2), (b"upper", False, -0.2))):
        fn = A.DENSITY_FN(lambda y, data, lower=lower, half=n * CELL / 2: 1.0 if ((y < half) == lower) else 0.0)
        dens = A.Density()
        dens.type, dens.custom_y = A.CUSTOM, fn
        keep += [fn, dens]
        lib.spec_new(C.byref(species[k]), name, -1.0, (C.c_int * 2)(8, 4), (C.c_float * 3)(ux, 0, 0), (C.c_float * 3)(0.01, 0.01, 0.01),
                     cnx, box, DT, C.byref(dens))
    sim = A.Simulation()
    lib.sim_new(C.byref(sim), cnx, box, DT, 1.0e9, 0, species, 2)
    sm = A.Smooth(A.BINOMIAL, A.BINOMIAL, 1, 1)
    lib.sim_set_smooth(C.byref(sim), C.byref(sm))
    return sim, species, 2, 32, DT, keep


def run_deck(args):
    """--workload lwfa | kh: BASELINE configs[2] / configs[3] through sim_new / sim_iter of the C API, slab-decomposed
    over the ranks of the job (strong scaling: the box is the configuration's, whatever N), device-side
    initialisation; one JSON line with the bench keys plus cell-updates/s.  Secondary lines: the driver's default
    run is the Weibel configuration."""
    from zpic_b200 import abi_em2d as A
    from zpic_b200 import load
    rank, world, local = (int(os.environ.get(k, "0" if k != "WORLD_SIZE" else "1")) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    lib = load("em2d")
    if lib.zdev_init(local) != 0:
        raise SystemExit("bench.py: no CUDA device - the CUDA path is the only path")
    ranks = Ranks(lib, world, local)
    lib.zpic_b200_set_option(b"device_init", 1)
    lib.zpic_b200_set_option(b"lazy", 1)
    K, W = args.steps, max(args.warmup, 3)
    t0 = time.time()
    if args.workload == "lwfa":
        nx, ny = args.lwfa_nx, args.lwfa_ny
        sim, species, nsp, ppc_total, dt, keep = build_lwfa(lib, A, nx, ny, args.lwfa_a0)
        what = ("em2d LWFA %dx%d cells, 16 ppc, laser a0 %g, moving window with host injection of the new column, "
                "compensated smoothing level 4 (BASELINE configs[2])" % (nx, ny, args.lwfa_a0))
    else:
        nx = ny = args.kh_n
        sim, species, nsp, ppc_total, dt, keep = build_kh(lib, A, nx)
        what = ("em2d Kelvin-Helmholtz %dx%d cells, 2 species x 32 ppc (half box each), binomial smoothing x,y level 1, "
                "periodic (BASELINE configs[3])" % (nx, ny))
    for _ in range(W):
        lib.sim_iter(C.byref(sim))
    lib.zdev_sync()
    t_init = time.time() - t0
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.zdev_launch_count()
    e0, e1 = lib.zdev_event_create(), lib.zdev_event_create()
    ranks.barrier()
    lib.zdev_event_record(e0)
    for _ in range(K):
        lib.sim_iter(C.byref(sim))
    lib.zdev_event_record(e1)
    ms = lib.zdev_event_elapsed_ms(e0, e1)
    ranks.barrier()
    ms = ranks.max(ms)
    clocks = sampler.finish() if rank == 0 else None
    launches = lib.zdev_launch_count() - launches0
    # particles of the whole box after the run (collective: the count is summed over the slabs by the library)
    lib.zpic_b200_set_option(b"lazy", 0)
    lib.sim_iter(C.byref(sim))
    npart = sum(int(species[k].np) for k in range(nsp))
    if rank == 0:
        peak, peak_src = peaks()
        value = npart * K / (ms * 1e-3)
        passes = 5 if args.workload == "lwfa" else 2
        bytes_step = BYTES_PER_PUSH * npart + (BYTES_PER_CELL + 24.0 * passes) * nx * ny
        gbs = bytes_step * K / (ms * 1e-3) / 1e9 / world
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": what, "particles": npart, "n_move": int(sim.emf.n_move), "dt": dt,
                       "init": "device-side counter-based distribution; laser launch on the host (libm double precision)",
                       "decomposition": "%d slabs along x, one process per GPU, through sim_new / sim_iter of the C API" % world,
                       "init_and_warmup_s": round(t_init, 1)},
            "cells": {"value": nx * ny * K / (ms * 1e-3), "unit": "cell-updates/s"},
            "roofline": {"bound": "hbm", "kernel": "whole step", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                         "traffic": None, "peak_source": peak_src,
                         "note": "per GPU, algorithmic bytes of the whole step: 56 B per push + (72 + 24 per smoothing pass) B per cell"},
            "gpu_launches": int(launches), "clocks": clocks, "e2e": None, "cpu_baseline": None}))
    lib.sim_delete(C.byref(sim))
    ranks.close()


def run_em1d(args, lib=None):
    """--workload em1d: BASELINE configs[4], the em1d two-stream deck scaled to 2^22 cells x 256 ppc x 2 beams
    (2^31 particles, 118 GB), driven through the device seam like scripts/quick_push_probe1d.py; one JSON line with
    the same keys.  A secondary line: the driver's default run is the em2d configuration above."""
    from zpic_b200._lib import PushParams1D
    if lib is None:
        from zpic_b200 import load
        lib = load("em1d")
    if lib.zdev_init(-1) != 0:
        raise SystemExit("bench.py: no CUDA device - the CUDA path is the only path")
    K, W = args.steps, max(args.warmup, 3)
    n, ppc = 1 << args.log2_cells, args.ppc1d
    dx = np.float32(4 * np.pi / 120)                # em1d/input/twostream.c:15-20
    dt = np.float32(0.1)
    g = lib.zdev_grid1d_create(n)
    specs = []
    # initial state: the reference's own for this deck - its random stream (default seeds, em1d/random.c:16-17)
    # continued on the device through both beams, species after species like spec_new draws them
    z, w, have, spare = C.c_uint32(67890), C.c_uint32(12345), C.c_int(0), C.c_double(0.0)
    klo_a, khi_a = np.zeros(n, dtype=np.int32), np.full(n, ppc, dtype=np.int32)
    klo, khi = klo_a.ctypes.data_as(C.POINTER(C.c_int)), khi_a.ctypes.data_as(C.POINTER(C.c_int))
    t_init0 = time.perf_counter()
    for k, sign in enumerate((1.0, -1.0)):
        sp = lib.zdev_spec1d_create(n, ppc, 0)
        ufl = (C.c_float * 3)(0.2 * sign, 0, 0)
        uth = (C.c_float * 3)(0.001, 0.001, 0.001)
        rc = lib.zdev_spec1d_inject_lattice(sp, ppc, ufl, uth, klo, khi, C.byref(z), C.byref(w), C.byref(have), C.byref(spare))
        assert rc == 0
        q = np.float32(-1.0) / np.float32(ppc)
        prm = PushParams1D(float(np.float32(0.5 * float(dt) / -1.0)), float(dt / dx), float(q * dx / dt), float(q), 0, 0)
        specs.append((sp, prm))
    lib.zdev_sync()
    t_init = time.perf_counter() - t_init0
    npart = 2 * n * ppc

    def step():
        lib.zdev_current1d_zero(g)
        for sp, prm in specs:
            lib.zdev_spec1d_advance(sp, g, g, C.byref(prm))
        lib.zdev_current1d_update(g, 1, 0, 0)
        lib.zdev_emf1d_advance(g, g, float(dt), float(dx), 0, 0)

    for _ in range(W):
        step()
    sampler = ClockSampler(0)
    e0, e1 = lib.zdev_event_create(), lib.zdev_event_create()
    lib.zdev_sync()
    launches0 = lib.zdev_launch_count()
    sampler.start()
    lib.zdev_event_record(e0)
    for _ in range(K):
        step()
    lib.zdev_event_record(e1)
    ms = lib.zdev_event_elapsed_ms(e0, e1)
    clocks = sampler.finish()
    launches = lib.zdev_launch_count() - launches0
    en, cnt = C.c_double(), C.c_int64()
    lib.zdev_spec1d_fetch(specs[0][0], C.byref(en), C.byref(cnt))
    assert cnt.value == n * ppc, "particles were lost: %d" % cnt.value
    peak, peak_src = peaks()
    value = npart * K / (ms * 1e-3)
    gbs = 40.0 * value / 1e9
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": ms / K,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "em1d two-stream 2^%d cells x %d ppc x 2 beams, periodic (BASELINE configs[4])" % (args.log2_cells, ppc),
                      "particles_per_gpu": npart, "dt": float(dt), "dx": float(dx),
                      "init": "the reference's initial state for this deck and seed, generated on the device from the reference's random stream",
                      "init_s": round(t_init, 2),
                      "cache": "working set %.1f GB per step >> 126 MB L2, no flush needed" % (npart * 44 / 1e9)},
           "roofline": {"bound": "hbm", "kernel": "k_push1d", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                        "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_particle": 40.0,
                        "note": "from the whole step (the grid kernels are below 1 % of it), not from per-launch events"},
           "gpu_launches": int(launches), "clocks": clocks, "e2e": None, "cpu_baseline": None}
    for sp, _ in specs:
        lib.zdev_spec1d_destroy(sp)
    lib.zdev_grid1d_destroy(g)
    print(json.dumps(out))
    return out


def _e2e_leg(lib, A, args, n, device_init, steps):
    """sim_iter + per-step diagnostics + the deck's report set every 10 steps, wall clock, host buffers"""
    ppc = (args.ppc, args.ppc)
    lib.zpic_b200_set_option(b"device_init", int(device_init))
    lib.zpic_b200_set_option(b"lazy", 0)
    lib.zpic_b200_set_option(b"coherent", 0)
    t0 = time.perf_counter()
    sim, species, _ = build_weibel(lib, A, n, n, ppc)
    t_build = time.perf_counter() - t0
    np_total = 2 * n * n * args.ppc * args.ppc
    grid_b = (n + 3) * (n + 3) * 12
    rho = [np.zeros((n + 1, n + 1), dtype=np.float32) for _ in range(2)]
    en6 = (C.c_double * 6)()

    def one_step(k):
        lib.sim_iter(C.byref(sim))
        lib.emf_get_energy(C.byref(sim.emf), en6)
        if (k + 1) % 10 == 0:                      # the deck's report cadence (input/weibel.c:44-57, ndump = 10)
            lib.zpic_b200_sync_emf(C.byref(sim.emf))
            lib.zpic_b200_sync_current(C.byref(sim.current))
            for s in range(2):
                rho[s][...] = 0
                lib.spec_deposit_charge(C.byref(species[s]), rho[s].ctypes.data_as(C.POINTER(C.c_float)))

    t0 = time.perf_counter()
    one_step(7)                                    # first step: the one-off upload (host-initialised) / generation
    lib.zdev_sync()
    t_first = time.perf_counter() - t0
    for k in range(8, 10):                         # warm-up, the last one with the report set
        one_step(k)                                # (first use page-locks the E, B, J mirrors)
    lib.zdev_sync()
    t0 = time.perf_counter()
    for k in range(steps):
        one_step(k)
    lib.zdev_sync()
    dt_res = time.perf_counter() - t0
    reports = steps // 10
    leg = {"value": np_total * steps / dt_res, "unit": UNIT,
           "h2d_bytes_per_step": int(2 * 32 + reports * 2 * (n + 1) * (n + 1) * 4 / steps),
           "d2h_bytes_per_step": int(2 * 48 + 48 + reports * (3 * grid_b + 2 * (n + 1) * (n + 1) * 4) / steps),
           "workload": "em2d Weibel %dx%d, 2 x %d ppc" % (n, n, args.ppc ** 2), "steps": steps, "wall_clock": True,
           "ms_per_step": dt_res / steps * 1e3,
           "one_off": {"spec_new_sim_new_s": round(t_build, 3), "first_step_s": round(t_first, 3),
                       "upload_bytes": 0 if device_init else np_total * 28 + 2 * grid_b,
                       "note": "not in `value`: building the deck, and the first sim_iter with the one-off "
                               + ("generation of the particles on the device" if device_init else "upload of the host-initialised particles and fields")}}
    return leg, sim, species


def run_e2e(lib, A, args, n_full):
    """The metric through the public C API as a caller uses it (reference em2d/main.c:53-59 and the Weibel deck's
    sim_report, input/weibel.c:44-57) on the HEADLINE configuration: sim_new, then every timed step is sim_iter
    followed by the energy diagnostics read back to the host (per-species kinetic energy and particle count, the six
    field energies; this synchronises every step), and every `ndump` = 10 steps the deck's report set arrives in
    host buffers (E, B, J mirrors and the two charge densities).  Wall clock around the loop.
    `host_initialised`: the same loop with the species created on the HOST by spec_new (the reference random
    stream) and uploaded once, at the largest size whose host initialisation takes seconds.
    `coherent`: NO state kept on the device - every sim_iter uploads all particles + E, B from the host mirrors and
    downloads particles + E, B, J again."""
    steps = max(args.steps, 10)
    main, sim, species = _e2e_leg(lib, A, args, n_full, 2, steps)
    lib.sim_delete(C.byref(sim))
    main["mode"] = ("public C API on BASELINE configs[1] (the reference's initial state, generated on the device): every step sim_iter + energy "
                    "diagnostics read back (a stream synchronisation per step); every 10 steps the deck's report set "
                    "(E, B, J, 2 charge grids) synchronised to host buffers")
    n = min(n_full, args.e2e_n)
    host, sim, species = _e2e_leg(lib, A, args, n, False, steps)
    host["mode"] = "the same loop, species created on the host by spec_new (reference random stream) and uploaded once"
    main["host_initialised"] = host
    np_total = 2 * n * n * args.ppc * args.ppc
    grid_b = (n + 3) * (n + 3) * 12
    # a caller written for the reference: no zpic_b200_* call anywhere, it just reads a raw buffer of the API after
    # every step (what the reference's Cython module does when a notebook looks at sim.emf.Ez): the mirrors are
    # guarded mappings (csrc/host/common/zb_guard.c), the read faults once per step and pulls E and B over
    lib.zb_guard_fills.restype = C.c_ulong
    E = A.grid_view(sim.emf.E_buf, n, n)
    fills0, acc = lib.zb_guard_fills(), 0.0
    lib.sim_iter(C.byref(sim))
    acc += float(E[1:-2, 1:-2, 2].max())
    t0 = time.perf_counter()
    for _ in range(steps):
        lib.sim_iter(C.byref(sim))
        acc += float(E[1:-2, 1:-2, 2].max())
    dt_raw = time.perf_counter() - t0
    main["unmodified_caller"] = {"value": np_total * steps / dt_raw, "unit": UNIT, "steps": steps, "workload": host["workload"],
                                 "h2d_bytes_per_step": 2 * 32, "d2h_bytes_per_step": 2 * grid_b + 2 * 48,
                                 "mirror_fills": int(lib.zb_guard_fills() - fills0),
                                 "note": "sim_iter + a numpy reduction over the raw E buffer of the API every step, no "
                                         "zpic_b200_* call: the stale mirror faults, the handler downloads E and B"}
    # strict host-buffer round trip
    lib.zpic_b200_set_option(b"coherent", 1)
    lib.sim_iter(C.byref(sim))
    lib.zdev_sync()
    csteps = 3
    t0 = time.perf_counter()
    for _ in range(csteps):
        lib.sim_iter(C.byref(sim))
    lib.zdev_sync()
    dt_coh = time.perf_counter() - t0
    lib.zpic_b200_set_option(b"coherent", 0)
    lib.sim_delete(C.byref(sim))
    main["coherent"] = {"value": np_total * csteps / dt_coh, "unit": UNIT, "steps": csteps,
                        "workload": host["workload"],
                        "h2d_bytes_per_step": np_total * 28 + 2 * grid_b, "d2h_bytes_per_step": np_total * 28 + 3 * grid_b,
                        "note": "ZPIC_COHERENT=1: no state kept on the device between calls - every sim_iter uploads all "
                                "particles + E,B from pageable host buffers and downloads particles + E,B,J"}
    return main


# ------------------------------------------------------------------------------------------
# the reference on the host

def ref_run(n, ppc, steps, fast=True):
    """one process, one core: the unmodified reference on a Weibel box of n x n cells"""
    from zpic_b200 import abi_em2d as A
    path = os.path.join(REPO, "oracle", "_ref", "libzpic_ref_em2d%s.so" % ("_fast" if fast else ""))
    if not os.path.exists(path):
        return None
    lib = A.declare(C.CDLL(path))
    sim, species, _ = build_weibel(lib, A, n, n, (ppc, ppc))
    lib.sim_iter(C.byref(sim))          # warm caches / first-touch
    p0, t0 = lib.spec_npush(), lib.spec_time()
    w0 = time.perf_counter()
    for _ in range(steps):
        lib.sim_iter(C.byref(sim))
    wall = time.perf_counter() - w0
    pushes = lib.spec_npush() - p0
    return {"pushes": int(pushes), "wall_s": wall, "spec_time_s": lib.spec_time() - t0}


def cpu_baseline(seconds=15.0, threads=1):
    """bounded sample of the same workload on `threads` host cores (independent replicas: the
    reference has no threads, em2d/Makefile:1-4)"""
    n, ppc = 256, 8                                   # 2 x 4.2 M particles, ~1 s per step per core
    per_step = 2 * n * n * ppc * ppc
    steps = max(2, int(seconds * 8.5e6 / per_step))
    code = ("import sys, json; sys.path.insert(0, %r); import bench; "
            "print(json.dumps(bench.ref_run(%d, %d, %d)))" % (REPO, n, ppc, steps))
    procs = [subprocess.Popen([sys.executable, "-c", code], stdout=subprocess.PIPE, text=True) for _ in range(threads)]
    results = []
    for p in procs:
        out = p.communicate()[0].strip().splitlines()
        results.append(json.loads(out[-1]) if out else None)
    if not results or any(r is None for r in results):
        return {"value": None, "unit": UNIT, "cores": threads, "kind": "reference",
                "sample": "oracle/_ref not built on this box"}
    wall = max(r["wall_s"] for r in results)
    pushes = sum(r["pushes"] for r in results)
    return {"value": pushes / wall, "unit": UNIT, "cores": threads, "kind": "reference",
            "sample": "unmodified reference (-Ofast, as shipped) em2d Weibel %dx%d, 2 x %d ppc, %d steps%s; "
                      "whole sim_iter wall time" % (n, n, ppc * ppc, steps,
                                                    " x %d independent replicas" % threads if threads > 1 else ""),
            "reference_spec_advance_only": sum(r["pushes"] for r in results) / max(r["spec_time_s"] for r in results)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    K = args.steps
    # each "step" = a bounded sample: every core advances its own 256^2 x 2 x 64 ppc replica once
    n, ppc = 256, 8
    steps_run = K + max(args.warmup - 1, 0)
    code = ("import sys, json; sys.path.insert(0, %r); import bench; "
            "print(json.dumps(bench.ref_run(%d, %d, %d)))" % (REPO, n, ppc, steps_run))
    procs = [subprocess.Popen([sys.executable, "-c", code], stdout=subprocess.PIPE, text=True) for _ in range(threads)]
    results = []
    for p in procs:
        o = p.communicate()[0].strip().splitlines()
        results.append(json.loads(o[-1]) if o else None)
    if any(r is None for r in results):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libzpic_ref_em2d_fast.so not present"}))
        return
    wall = max(r["wall_s"] for r in results)
    pushes = sum(r["pushes"] for r in results)
    value = pushes / wall
    sample = ("unmodified reference (-Ofast) em2d Weibel %dx%d, 2 x %d ppc per replica, %d independent replicas "
              "(one per host core; the reference is single threaded)" % (n, n, ppc * ppc, threads))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": args.warmup, "ms_per_step": wall / steps_run * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "em2d Weibel, 2 species x 64 ppc, periodic (BASELINE configs[1] physics)", "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", dest="n", type=int, default=4096, help="grid cells per side per GPU")
    ap.add_argument("--ppc", type=int, default=8, help="particles per cell per direction (8 -> 64 ppc)")
    ap.add_argument("--e2e-grid", type=int, default=1024, dest="e2e_n")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, dest="cpu_seconds")
    ap.add_argument("--workload", default="em2d", choices=["em2d", "em1d", "lwfa", "kh"],
                    help="em2d = BASELINE configs[1] (the default, what the driver runs); lwfa = configs[2], kh = configs[3] "
                         "(slab-decomposed over the ranks of the job), em1d = configs[4] (one GPU)")
    ap.add_argument("--lwfa-nx", type=int, default=16384, dest="lwfa_nx")
    ap.add_argument("--lwfa-ny", type=int, default=1024, dest="lwfa_ny")
    ap.add_argument("--lwfa-a0", type=float, default=5.0, dest="lwfa_a0",
                    help="lwfa: normalised vector potential of the laser (5 = em2d/input/lwfa-large.c:44; the round-2 runs before this option existed used 3)")
    ap.add_argument("--kh-n", type=int, default=8192, dest="kh_n")
    ap.add_argument("--no-check", action="store_true", dest="no_check", help="N > 1: skip the slab-parity check")
    ap.add_argument("--no-extras", action="store_true", dest="no_extras",
                    help="skip `other_configs` (the other BASELINE configurations, run as children after the main one)")
    ap.add_argument("--init", type=int, default=0, help="device-side initialisation: 2 = the reference random stream "
                    "(default at N = 1), 1 = counter-based generator (default at N > 1)")
    ap.add_argument("--log2-cells", type=int, default=22, dest="log2_cells", help="em1d: log2 of the cell count")
    ap.add_argument("--ppc1d", type=int, default=256, help="em1d: particles per cell per beam")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "em1d":
        run_em1d(args)
    elif args.workload in ("lwfa", "kh"):
        run_deck(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
