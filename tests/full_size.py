"""Size-independent properties of the time step, measured through the reference C API.

    python -m tests.full_size em2d 4096 8 5 [--lib ours|ref] [--device-init]
    python -m tests.full_size em1d 22 256 5 [--lib ours|ref] [--device-init]

At BASELINE.json's full sizes (2^31 particles) nothing can be compared particle by particle with the CPU
reference, so the parity of the big runs rests on what a correct step must preserve whatever the size:

  np            the periodic box keeps every particle of every species
  charge_sum    sum(rho_k) * cell volume = q_k * np_k  (the charge deposit is a partition of unity)
  gauss         div E - rho stays at its initial value: the deposited current is exactly charge conserving
                (em2d/particles.c:773-924 dep_current_zamb, em1d/particles.c:707-779)
  energy        field + kinetic energy (em2d/simulation.c:183-206) drifts by rounding / finite dt only

Prints one JSON object.  Run against the reference build (--lib ref, small sizes) the same script gives the
numbers the thresholds in tests/test_gpu_full_size.py were calibrated on.  Test infrastructure: not imported
by the product.
"""
import argparse
import ctypes as C
import json
import sys
import time

import numpy as np


def _deck(code, lib, n, ppc):
    if code == "em2d":
        from tests import helpers as H
        return H.weibel(lib, n=n, ppc=(ppc, ppc))          # em2d/input/weibel.c:13-40
    from tests import helpers1d as H1
    return H1.twostream(lib, nx=n, ppc=ppc)                 # em1d/input/twostream.c:12-36


def _sync_emf(deck):
    if hasattr(deck.lib, "zpic_b200_sync_emf"):
        deck.lib.zpic_b200_sync_emf(C.byref(deck.sim.emf))   # fields only: the particles stay in HBM


def _div_e(code, deck):
    """div E on the nodes that carry rho (Yee: Ex at i+1/2, Ey at j+1/2), interior nodes only"""
    _sync_emf(deck)
    E = deck.E()
    if code == "em2d":
        ny, nx = E.shape[0] - 3, E.shape[1] - 3
        dx, dy = float(deck.sim.emf.dx[0]), float(deck.sim.emf.dx[1])
        ex = E[1:ny + 1, :, 0].astype(np.float64)
        ey = E[:, 1:nx + 1, 1].astype(np.float64)
        return (ex[:, 1:nx + 1] - ex[:, 0:nx]) / dx + (ey[1:ny + 1, :] - ey[0:ny, :]) / dy
    nx = E.shape[0] - 3
    ex = E[:, 0].astype(np.float64)
    return (ex[1:nx + 1] - ex[0:nx]) / float(deck.sim.emf.dx)


def _rho(code, deck):
    """per-species charge density on the interior nodes + its plain sum over the whole deposit grid"""
    out, sums = [], []
    for k in range(deck.n_species):
        r = deck.charge(k).astype(np.float64)
        sums.append(r[:-1, :-1].sum() if code == "em2d" else r[:-1].sum())
        out.append(r[:-1, :-1] if code == "em2d" else r[:-1])
    return out, sums


def _energy(deck):
    return float(deck.emf_energy().sum()), [float(deck.species[k].energy) for k in range(deck.n_species)]


def measure(code, lib, n, ppc, steps):
    t0 = time.time()
    deck = _deck(code, lib, n, ppc)
    cells = n * n if code == "em2d" else n
    np0 = [int(deck.species[k].np) for k in range(deck.n_species)]
    rho0, _ = _rho(code, deck)
    scale = float(np.abs(rho0[0]).max())
    res0 = _div_e(code, deck) - sum(rho0)                # E(0) = 0: this is -rho(0)
    deck.iter(1)
    en1 = _energy(deck)
    deck.iter(steps - 1)
    enK = _energy(deck)
    rho, sums = _rho(code, deck)
    res = _div_e(code, deck) - sum(rho)
    npK = [int(deck.species[k].np) for k in range(deck.n_species)]
    ppc_total = ppc * ppc if code == "em2d" else ppc
    q = [float(deck.species[k].q) for k in range(deck.n_species)]      # sign(m_q) / ppc (em2d/particles.c:557)
    out = {
        "code": code, "cells": cells, "ppc": ppc_total, "steps": steps, "np0": np0, "npK": npK,
        "charge_sum_rel": [abs(s - qk * m) / abs(qk * m) for s, qk, m in zip(sums, q, npK)],
        "gauss_max": float(np.abs(res - res0).max()) / scale,
        "gauss_rms": float(np.sqrt(((res - res0) ** 2).mean())) / scale,
        "gauss_at": [int(v) for v in np.unravel_index(int(np.abs(res - res0).argmax()), res.shape)],
        "energy_1": en1[0] + sum(en1[1]), "energy_K": enK[0] + sum(enK[1]),
        "field_energy_K": enK[0],
        "seconds": None,
    }
    bad = np.argwhere(np.abs(res - res0) > 1e-5 * scale)
    out["gauss_bad"] = int(len(bad))
    out["gauss_bad_cells"] = [[int(j), int(i), float((res - res0)[j, i]) / scale] for j, i in bad[:24]] if code == "em2d" else []
    out["charge_sum_err"] = [float(s - qk * m) for s, qk, m in zip(sums, q, npK)]
    out["energy_rel_drift"] = abs(out["energy_K"] - out["energy_1"]) / abs(out["energy_1"])
    deck.delete()
    out["seconds"] = round(time.time() - t0, 2)
    return out


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("code", choices=("em2d", "em1d"))
    ap.add_argument("n", type=int, help="em2d: cells per side; em1d: log2 of the cell count")
    ap.add_argument("ppc", type=int, help="em2d: particles per cell per direction; em1d: particles per cell")
    ap.add_argument("steps", type=int)
    ap.add_argument("--lib", choices=("ours", "ref"), default="ours")
    ap.add_argument("--device-init", action="store_true",
                    help="generate the species on the device (populations whose host mirror is impractical)")
    a = ap.parse_args(argv)
    n = a.n if a.code == "em2d" else 1 << a.n
    if a.lib == "ours":
        from zpic_b200 import load
        lib = load(a.code)
        if lib.zdev_init(-1) != 0:
            raise SystemExit("full_size: no CUDA device - the CUDA path is the only path")
        # a population that does not fit the free device memory is reported as such (exit code 77: the test skips)
        # instead of ending in the allocator's fatal error
        free_b, total_b = C.c_size_t(), C.c_size_t()
        lib.zdev_mem_info(C.byref(free_b), C.byref(total_b))
        cells = n * n if a.code == "em2d" else n
        per_particle = (2 * 26 * 1.25 + 4.4) if a.code == "em2d" else (2 * 22 * 1.25 + 3.6)     # A/B records + keys, migrants, overflow list
        need = 2.0 * cells * (a.ppc * a.ppc if a.code == "em2d" else a.ppc) * per_particle + 6 * 12.0 * cells
        if need > 0.97 * free_b.value:
            sys.stderr.write("full_size: %.1f GB needed, %.1f GB free\n" % (need / 1e9, free_b.value / 1e9))
            return 77
        lib.zpic_b200_set_option(b"device_init", int(a.device_init))
        lib.zpic_b200_set_option(b"lazy", 0)
        lib.zpic_b200_set_option(b"coherent", 0)
        lib.zpic_b200_set_option(b"track_ids", 0)
    else:
        from tests import helpers as H
        from tests import helpers1d as H1
        lib = H.load_ref("em2d") if a.code == "em2d" else H1.load_ref()
        if lib is None:
            raise SystemExit("full_size: oracle/_ref is not built")
    print(json.dumps(measure(a.code, lib, n, a.ppc, a.steps)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
