"""Per-kernel totals of the LAST pass in an ncu launch list (gpu__time_duration.sum CSV): python scripts/summarize_launches.py file.csv [first-kernel-substring]"""
import csv, sys
from collections import OrderedDict
rows = list(csv.reader(open(sys.argv[1])))
first = sys.argv[2] if len(sys.argv) > 2 else "ybar"
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
ki, vi, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
data = [(r[ki][:60], r[gi], float(r[vi].replace(",", ""))) for r in rows[hdr + 1:] if len(r) > vi]
idx = [i for i, d in enumerate(data) if first in d[0]]
last = data[idx[-1]:]
agg, tot = OrderedDict(), 0.0
for name, g, v in last:
    k = name.split("(")[0]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
for k, (c, v) in agg.items():
    print("%-46s x%-3d %9.1f us" % (k, c, v / 1e3))
print("total %.1f us over %d launches" % (tot / 1e3, len(last)))
