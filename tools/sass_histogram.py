"""SASS opcode histogram of every kernel in the built objects (multi-modal-monodepth-estimation_b200/build/*.o):
which kernels carry tcgen05 (UTC*MMA / LDTM / STTM), TMA (UTMALDG / UTMASTG / UBLKCP), cp.async (LDGSTS) or legacy
tensor instructions (HMMA).  Runs on the CPU box (cuobjdump only).

    python tools/sass_histogram.py > profiles/r02_sass_histogram.txt
"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "LDGSTS", "HMMA", "MUFU", "SYNCS",
        "UTCBAR", "UTCCP", "BAR", "ATOMS", "ATOMG", "RED", "STL", "LDL"]
print("SASS opcode counts per kernel (cuobjdump -sass of the shipped objects; sm_100a)\n")
print(f"{'kernel':70s} " + " ".join(f"{k:>8s}" for k in KEYS) + "   total")
for obj in sorted(glob.glob(os.path.join(ROOT, "multi-modal-monodepth-estimation_b200", "build", "*.o"))):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    fn, counts = None, None
    rows = []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if fn:
                rows.append((fn, counts))
            fn, counts = m.group(1), collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and fn:
            op = m.group(1)
            counts["__total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    counts[k] += 1
    if fn:
        rows.append((fn, counts))
    print(f"--- {os.path.basename(obj)}")
    for fn, c in rows:
        name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip() or fn
        name = re.sub(r"\(anonymous namespace\)::|b200swin::|unnamed>::", "", name)
        name = re.sub(r"\(.*", "", name)[:69]
        print(f"{name:70s} " + " ".join(f"{c[k]:8d}" for k in KEYS) + f" {c['__total']:7d}")
