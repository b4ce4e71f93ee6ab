"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): time share per kernel."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
hdr = rows[0]
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
tot = defaultdict(float); cnt = defaultdict(int)
for r in rows[1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r"b200ms::", "", name)
    g = r[gi]
    key = name if not name.startswith("stencil_march") else f"{name} grid={g}"
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    tot[key] += v; cnt[key] += 1
T = sum(tot.values())
print(f"total {T/1e6:.2f} ms over {sum(cnt.values())} launches")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:45]:
    print(f"{100*v/T:6.2f}%  {v/1e6:9.3f} ms  n={cnt[k]:5d}  avg {v/cnt[k]/1e3:8.1f} us  {k[:130]}")
