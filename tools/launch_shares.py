"""Aggregates an ncu --csv launch list (gpu__time_duration.sum) by kernel name."""
import csv
import collections
import re
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
name_i, val_i, unit_i = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[val_i].replace(",", ""))
    u = r[unit_i]
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(u, 1e-3)
    k = re.sub(r"\(.*", "", r[name_i])
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"total {tot / 1e3:.3f} ms over {sum(a[0] for a in agg.values())} launches")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{a[1] / 1e3:9.3f} ms {100 * a[1] / tot:5.1f} %  {a[0]:5d} x  {k[:110]}")
