"""Join the per-instruction samples of an .ncu-rep (SASS page) with the line table of the current build
(nvdisasm -g of the kernel's cubin) and list the source lines that hold the stall samples / the executed
instructions.  The SASS of the capture and of the build must be the same kernel binary (checked by length and
opcode).  usage: python scripts/ncu_source_lines.py x.ncu-rep obj.o 'mangled kernel name' [top]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, obj, fun = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
sass = [dict(zip(hdr, r)) for r in rows[2:] if len(r) == len(hdr)]

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
lines, cur, inside = [], None, False
for l in dis:
    if l.startswith("\t.section\t.text."):
        inside = l.startswith("\t.section\t.text." + fun + ",")
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "(.*?)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m:
        lines.append((cur, m.group(2).strip()))
assert len(lines) == len(sass), (len(lines), len(sass))
for (loc, txt), d in zip(lines, sass):
    a = txt.split()[1 if txt.startswith("@") else 0].split(".")[0]
    s = d["Source"].split()
    b = s[1 if s[0].startswith("@") else 0].split(".")[0]
    assert a == b, (txt, d["Source"])

samp, inst = collections.Counter(), collections.Counter()
for (loc, _), d in zip(lines, sass):
    samp[loc] += float(d["# Samples"] or 0)
    inst[loc] += float(d["Instructions Executed"] or 0)
ts, ti = sum(samp.values()), sum(inst.values())
src = {}
for loc in samp:
    if loc and loc[0] not in src:
        for base, _, fs in os.walk(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "zpic_b200", "csrc")):
            if loc[0] in fs:
                src[loc[0]] = open(os.path.join(base, loc[0])).read().splitlines()
print("%d SASS instructions, %d stall samples, %.0f executed warp-instructions" % (len(sass), ts, ti))
print("  samples   instr   line")
for loc, s in samp.most_common(top):
    text = src.get(loc[0], [""] * (loc[1] + 1))[loc[1] - 1].strip()[:100] if loc else ""
    print("  %5.1f %%  %5.1f %%  %s:%d  %s" % (100 * s / ts, 100 * inst[loc] / ti, loc[0], loc[1], text))
