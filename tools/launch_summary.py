#!/usr/bin/env python
"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count, average, share."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = collections.OrderedDict()
for r in rows[1 + skip:]:
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[iu], 1e-3)
    name = r[ik].split("(")[0]
    a = agg.setdefault(name, [0, 0.0, 1e30, 0.0])
    a[0] += 1; a[1] += v; a[2] = min(a[2], v); a[3] = max(a[3], v)
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':28s} {'launches':>8s} {'avg us':>10s} {'min us':>10s} {'max us':>10s} {'share':>7s}")
for k, a in agg.items():
    print(f"{k:28s} {a[0]:8d} {a[1]/a[0]:10.1f} {a[2]:10.1f} {a[3]:10.1f} {100*a[1]/tot:6.1f}%")
print(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches")
