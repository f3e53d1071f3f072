#!/usr/bin/env python
"""Per-kernel time + DRAM traffic from an ncu csv with gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ik, im, iv, iu, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6572.2
per = collections.OrderedDict()
for r in rows[1:]:
    key = (r[iid], r[ik].split("(")[0])
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    if r[im].startswith("gpu__time"):
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(u, 1e-3)
        per.setdefault(key, {})["us"] = v
    else:
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "bytes": 1.0}.get(u, 1.0)
        per.setdefault(key, {})["rd" if "read" in r[im] else "wr"] = v
agg = collections.OrderedDict()
for (_, name), d in per.items():
    a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += d.get("us", 0); a[2] += d.get("rd", 0); a[3] += d.get("wr", 0)
print(f"{'kernel':22s} {'launches':>8s} {'avg us':>9s} {'DRAM rd MB':>11s} {'DRAM wr MB':>11s} {'GB/s':>8s} {'% of peak':>9s}")
for k, a in agg.items():
    n = a[0]; us = a[1] / n; rd = a[2] / n; wr = a[3] / n
    gbs = (rd + wr) / (us * 1e-6) / 1e9 if us > 0 else 0
    print(f"{k:22s} {n:8d} {us:9.1f} {rd/1e6:11.3f} {wr/1e6:11.3f} {gbs:8.1f} {100*gbs/peak:8.2f}%")
