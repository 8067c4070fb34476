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


def run(exe, where):
    t0 = time.time()
    r = subprocess.run([exe], cwd=where, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise SystemExit("%s failed (%d):\n%s" % (exe, r.returncode, r.stdout[-2000:]))
    return time.time() - t0, r.stdout


def files(root):
    out = {}
    for base, _, fs in os.walk(root):
        for f in fs:
            if f.endswith(".zdf"):
                out[os.path.relpath(os.path.join(base, f), root)] = os.path.join(base, f)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("code", choices=("em2d", "em1d"))
    ap.add_argument("--upto", type=int, default=100, help="compare the files of iterations <= this")
    ap.add_argument("--tol", type=float, default=1e-5)
    ap.add_argument("--self-check", action="store_true")
    ap.add_argument("--ours", default=None, help="another executable to put in place of <code>_ours")
    a = ap.parse_args()
    d = os.path.join(REPO, "oracle", "_ref", "decks")
    ref_exe = os.path.join(d, a.code + "_ref")
    our_exe = ref_exe if a.self_check else (a.ours or os.path.join(d, a.code + "_ours"))
    for e in (ref_exe, our_exe):
        if not os.path.exists(e):
            raise SystemExit("%s missing: run `make -C oracle decks` where the reference tree exists" % e)
    with tempfile.TemporaryDirectory() as ta, tempfile.TemporaryDirectory() as tb:
        t_ours, log = run(our_exe, ta)
        t_ref, _ = run(ref_exe, tb)
        fa, fb = files(ta), files(tb)
        assert set(fa) == set(fb), sorted(set(fa) ^ set(fb))[:10]
        # Grids are compared per FIELD VECTOR and iteration, as the north star states its tolerance (relative L2
        # on E, B, J): the components of one vector (EMF/E0,E1,E2-000010.zdf ...) enter one norm.  The worst
        # single component (relative to its own RMS, floored at 1e-3 of the largest value of its family) is
        # reported next to it; a weak component of a vector carries the rounding noise of the strong ones.
        worst, worst_comp, n_cmp, scale, groups = {}, {}, 0, {}, {}
        for rel in sorted(fb):
            it = int(re.search(r"-(\d{6})\.zdf$", rel).group(1))
            if it > a.upto:
                continue
            x, info = zdf.read(fa[rel])
            y, _ = zdf.read(fb[rel])
            fam = rel.split(os.sep)[0]
            n_cmp += 1
            if info.type == "particles":
                keys = sorted(y)
                ox = np.lexsort([x[k] for k in keys])
                oy = np.lexsort([y[k] for k in keys])
                assert len(ox) == len(oy), rel
                err = max(float(np.abs(x[k][ox] - y[k][oy]).max()) / max(float(np.abs(y[k]).max()), 1e-30) for k in keys) if len(oy) else 0.0
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
            if it == a.upto:
                at_end[fam] = max(at_end.get(fam, 0.0), err)
            for x, y in comps:
                e = float(np.sqrt(((x - y) ** 2).mean()) / max(np.sqrt((y ** 2).mean()), 1e-3 * scale[fam], 1e-30))
                worst_comp[fam] = max(worst_comp.get(fam, 0.0), e)
    # the bar is stated at a given number of steps (fields that have grown out of the noise): judged on the dumps
    # of iteration --upto; the worst over all earlier dumps (tiny fields, relative noise) is reported beside it
    ok = bool(at_end) and all(v <= a.tol for v in at_end.values())
    print(json.dumps({"code": a.code, "files": len(fb), "compared": n_cmp, "upto": a.upto, "rel_err_at_upto": at_end, "worst_rel_err_upto": worst, "worst_single_component": worst_comp,
                      "seconds_ours": round(t_ours, 2), "seconds_reference": round(t_ref, 2), "ok": ok,
                      "last_lines_ours": log.strip().splitlines()[-3:]}))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
