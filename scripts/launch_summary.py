"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): launches, total time and share per kernel.
usage: python scripts/launch_summary.py gpurun_out/bench_launches.csv"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
iK, iM, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    if r[iM] != "gpu__time_duration.sum":
        continue
    v = float(r[iV].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iU], 1.0)
    name = re.sub(r"\(.*", "", r[iK]).replace("void ", "").strip()
    tot[name] += v
    cnt[name] += 1
allus = sum(tot.values())
print("%-42s %8s %12s %7s" % ("kernel", "launches", "total_us", "share"))
for k, v in tot.most_common():
    print("%-42s %8d %12.1f %6.1f%%" % (k, cnt[k], v, 100 * v / allus))
