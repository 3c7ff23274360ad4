"""Summarise an `ncu --page source --csv` dump: top sampled SASS instructions and per-opcode shares."""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ia, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
data = []
for r in rows[hi + 1:]:
    if len(r) <= isamp or not r[isamp].isdigit():
        continue
    data.append((int(r[isamp]), int(r[iex] or 0), r[ia].strip()))
tot = sum(d[0] for d in data) or 1
totex = sum(d[1] for d in data) or 1
print("total samples", tot, "warp-instr executed", totex, "SASS instrs", len(data))
for s, e, src in sorted(data, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{s:7d} {100*s/tot:5.1f}% ex={e:9d}  {src[:100]}")
c, ce = Counter(), Counter()
for s, e, src in data:
    t = src.split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    c[op] += s; ce[op] += e
print("--- by opcode (samples%, exec%)")
for op, s in c.most_common(24):
    print(f"{op:12s} {100*s/tot:5.1f}%  {100*ce[op]/totex:5.1f}%")
