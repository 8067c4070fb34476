"""List the loops (backward branches) of a kernel's SASS with their static size and op mix.
usage: cuobjdump -sass -fun NAME lib.o | python scripts/sass_loops.py [min_size]"""
import collections
import re
import sys

min_size = int(sys.argv[1]) if len(sys.argv) > 1 else 40
ins = []
for line in sys.stdin:
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr2idx = {a: i for i, (a, _) in enumerate(ins)}
print("%d instructions" % len(ins))
for i, (a, txt) in enumerate(ins):
    m = re.search(r"BRA(?:\.\w+)*\s+(?:\w+,\s*)?(0x[0-9a-f]+)", txt)
    if m:
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr2idx and i - addr2idx[tgt] >= min_size:
            j = addr2idx[tgt]
            ops = collections.Counter()
            for _, t in ins[j:i + 1]:
                toks = t.split()
                op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
                ops[op] += 1
            print("loop 0x%x..0x%x: %d instructions: %s" % (tgt, a, i - j + 1, ", ".join("%s %d" % kv for kv in ops.most_common(40))))
