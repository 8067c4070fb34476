"""Basic-block view of a kernel capture: consecutive SASS instructions with the same execution count are one
block; prints address range, size, executions per instruction, share of executed warp-instructions and of the
stall samples, and the block's op mix.  usage: python scripts/ncu_blocks.py x.ncu-rep [min_share_pct]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
sass = [dict(zip(hdr, r)) for r in rows[2:] if len(r) == len(hdr)]
ti = sum(float(d["Instructions Executed"] or 0) for d in sass)
ts = sum(float(d["# Samples"] or 0) for d in sass)
print("%d SASS instructions, %.0f executed warp-instructions, %.0f samples" % (len(sass), ti, ts))
blocks, cur = [], None
for k, d in enumerate(sass):
    n = float(d["Instructions Executed"] or 0)
    if cur is None or n != cur[2]:
        cur = [k, k, n, 0.0, collections.Counter()]
        blocks.append(cur)
    cur[1] = k
    cur[3] += float(d["# Samples"] or 0)
    s = d["Source"].split()
    cur[4][s[1 if s[0].startswith("@") else 0].split(".")[0]] += 1
for b in blocks:
    size = b[1] - b[0] + 1
    share = 100.0 * size * b[2] / ti
    if share < min_share and 100.0 * b[3] / ts < min_share:
        continue
    print("%5d..%5d %4d instr x %10.0f  %5.1f %% instr %5.1f %% samples  %s" % (
        b[0], b[1], size, b[2], share, 100.0 * b[3] / ts, " ".join("%s%d" % kv for kv in b[4].most_common(8))))
