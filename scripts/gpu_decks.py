"""Run the reference PROGRAM both ways on a GPU box and compare what it writes: oracle/_ref/decks/<code>_ref is the
unmodified reference (main.c + shipped deck + its own .c files), <code>_ours the same unmodified main.c + deck
linked against libzpic_b200_<code>.so (built by `make -C oracle decks` where the reference tree exists; the
executables travel with the tree).  Every ZDF file the two runs write up to --upto iterations is compared:
grids by relative L2 per field vector (the north star's 1e-5 is judged on the dumps of iteration --upto),
particle files as canonically sorted sets.

    python scripts/gpu_decks.py em2d [--upto 100]        # em2d/input/weibel.c as shipped (501 iterations)
    python scripts/gpu_decks.py em1d                     # em1d/input/twostream.c as shipped
    python scripts/gpu_decks.py em2d --self-check        # CPU: reference against itself (exercises this script)
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from zpic_b200 import zdf  # noqa: E402


def run(exe, where, nranks=1):
    """nranks > 1: the SAME program started once per slab (ZPIC_RANK / ZPIC_NRANKS / ZPIC_DEVICE), all in one directory;
    the library decomposes the box, rank 0 writes the files"""
    t0 = time.time()
    if nranks <= 1:
        r = subprocess.run([exe], cwd=where, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise SystemExit("%s failed (%d):\n%s" % (exe, r.returncode, r.stdout[-2000:]))
        return time.time() - t0, r.stdout
    procs = []
    for k in range(nranks):
        env = dict(os.environ, ZPIC_RANK=str(k), ZPIC_NRANKS=str(nranks), ZPIC_DEVICE=str(k), ZPIC_JOB="deck%d" % os.getpid())
        procs.append(subprocess.Popen([exe], cwd=where, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate()[0] for p in procs]
    if any(p.returncode != 0 for p in procs):
        raise SystemExit("%s x %d failed:\n%s" % (exe, nranks, "\n".join(o[-1500:] for o in outs)))
    return time.time() - t0, outs[0]


def files(root):
    out = {}
    for base, _, fs in os.walk(root):
        for f in fs:
            if f.endswith(".zdf"):
                out[os.path.relpath(os.path.join(base, f), root)] = os.path.join(base, f)
    return out


def distances(exe_a, exe_b, upto, nranks=1):
    """run two builds of the program and compare every ZDF file of iterations <= upto.
    Returns (at_upto, worst, worst_component, n_files, n_compared, seconds_a, seconds_b, log_a)."""
    with tempfile.TemporaryDirectory() as ta, tempfile.TemporaryDirectory() as tb:
        t_a, log = run(exe_a, ta, nranks)
        t_b, _ = run(exe_b, tb)
        fa, fb = files(ta), files(tb)
        assert set(fa) == set(fb), sorted(set(fa) ^ set(fb))[:10]
        # Grids are compared per FIELD VECTOR and iteration, as the north star states its tolerance (relative L2
        # on E, B, J): the components of one vector (EMF/E0,E1,E2-000010.zdf ...) enter one norm.  The worst
        # single component (relative to its own RMS, floored at 1e-3 of the largest value of its family) is
        # reported next to it; a weak component of a vector carries the rounding noise of the strong ones.
        worst, worst_comp, n_cmp, scale, groups = {}, {}, 0, {}, {}
        for rel in sorted(fb):
            it = int(re.search(r"-(\d{6})\.zdf$", rel).group(1))
            if it > upto:
                continue
            x, info = zdf.read(fa[rel])
            y, _ = zdf.read(fb[rel])
            fam = rel.split(os.sep)[0]
            n_cmp += 1
            if info.type == "particles":
                # permutation-invariant: the sorted values of every quantity (the two runs order their buffers
                # differently, and pairing near-coincident particles by position is not robust)
                assert all(len(x[k]) == len(y[k]) for k in y), rel
                err = max(float(np.abs(np.sort(x[k]) - np.sort(y[k])).max()) / max(float(np.abs(y[k]).max()), 1e-30)
                          for k in y) if len(next(iter(y.values()))) else 0.0
                worst[fam] = max(worst.get(fam, 0.0), err)
                continue
            base = os.path.basename(rel)
            vec = re.sub(r"\d(-\d{6}\.zdf)$", r"\1", base) if re.match(r"^[EBJ]\d-", base) else base
            groups.setdefault((fam, os.path.dirname(rel), vec, it), []).append((x.astype(np.float64), y.astype(np.float64)))
            scale[fam] = max(scale.get(fam, 0.0), float(np.abs(y).max()))
        at_end = {}
        for (fam, _, _, it), comps in groups.items():
            num = sum(((x - y) ** 2).sum() for x, y in comps)
            den = sum((y ** 2).sum() for x, y in comps)
            size = sum(y.size for x, y in comps)
            floor = (1e-3 * scale[fam]) ** 2 * size
            err = float(np.sqrt(num / max(den, floor, 1e-300)))
            worst[fam] = max(worst.get(fam, 0.0), err)
            if it == upto:
                at_end[fam] = max(at_end.get(fam, 0.0), err)
            for x, y in comps:
                e = float(np.sqrt(((x - y) ** 2).mean()) / max(np.sqrt((y ** 2).mean()), 1e-3 * scale[fam], 1e-30))
                worst_comp[fam] = max(worst_comp.get(fam, 0.0), e)
    return at_end, worst, worst_comp, len(fb), n_cmp, t_a, t_b, log


def compare(code, upto=100, tol=1e-5, ours=None, self_check=False, noise=True, noise_factor=4.0, nranks=1):
    """The bar: at iteration `upto` every field family (EMF, CURRENT, CHARGE, PHASESPACE) of the CUDA build is
    within `tol` (relative L2) of the strict reference build - or, where the deck amplifies rounding noise past
    that (the cold two-stream instability of the shipped em1d deck: the reference's OWN -Ofast and strict builds
    are 1.7e-3 apart in E at iteration 100), within `noise_factor` x the distance between the reference's own two
    builds, which is measured in the same call and reported beside ours."""
    d = os.path.join(REPO, "oracle", "_ref", "decks")
    ref_exe = os.path.join(d, code + "_ref")
    our_exe = ref_exe if self_check else (ours or os.path.join(d, code + "_ours"))
    fast_exe = os.path.join(d, code + "_ref_fast")
    for e in (ref_exe, our_exe):
        if not os.path.exists(e):
            raise SystemExit("%s missing: run `make -C oracle decks` where the reference tree exists" % e)
    at_end, worst, worst_comp, n_files, n_cmp, t_ours, t_ref, log = distances(our_exe, ref_exe, upto, nranks)
    floor = {}
    if noise and os.path.exists(fast_exe):
        floor = distances(fast_exe, ref_exe, upto)[0]
    # the bar is stated at a given number of steps (fields that have grown out of the noise): judged on the dumps
    # of iteration `upto`; the worst over all earlier dumps (tiny fields, relative noise) is reported beside it
    ok = bool(at_end) and all(v <= max(tol, noise_factor * floor.get(k, 0.0)) for k, v in at_end.items())
    return {"code": code, "slabs": nranks, "files": n_files, "compared": n_cmp, "upto": upto, "tol": tol, "rel_err_at_upto": at_end,
            "reference_fast_vs_strict_at_upto": floor, "noise_factor": noise_factor,
            "worst_rel_err_upto": worst, "worst_single_component": worst_comp,
            "seconds_ours": round(t_ours, 2), "seconds_reference": round(t_ref, 2), "ok": ok,
            "last_lines_ours": log.strip().splitlines()[-3:]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("code", choices=("em2d", "em1d"))
    ap.add_argument("--upto", type=int, default=100, help="compare the files of iterations <= this")
    ap.add_argument("--tol", type=float, default=1e-5)
    ap.add_argument("--self-check", action="store_true")
    ap.add_argument("--ours", default=None, help="another executable to put in place of <code>_ours")
    ap.add_argument("--gpus", type=int, default=1, help="em2d: run the program as this many slabs (one process each)")
    a = ap.parse_args()
    res = compare(a.code, a.upto, a.tol, a.ours, a.self_check, nranks=a.gpus)
    print(json.dumps(res))
    return 0 if res["ok"] else 1


if __name__ == "__main__":
    sys.exit(main())
