"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and share per kernel name."""
import csv, re, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if r]
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= iv:
        continue
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    unit = r[iu]
    ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v if unit in ("ms", "msecond") else v / 1e6
    name = re.sub(r"\(.*", "", r[ik])
    name = re.sub(r"b200swin::\(anonymous namespace\)::|b200swin::|void |<unnamed>::", "", name)[:100]
    agg[name][0] += 1
    agg[name][1] += ms
tot = sum(v[1] for v in agg.values())
print(f"captured launches: {sum(v[0] for v in agg.values())}, total {tot:.1f} ms")
for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{ms:9.3f} ms {100 * ms / tot:5.1f}%  n={n:5d}  {name}")
