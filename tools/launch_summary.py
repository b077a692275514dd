#!/usr/bin/env python
"""Per-kernel summary (launches, total us, share of all device time) of an `ncu --metrics gpu__time_duration.sum --csv`
launch list. usage: python tools/launch_summary.py <launches.csv> > <launches_summary.csv>"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
for r in rows[1:]:
    a = agg.setdefault(r[ik], [0, 0.0])
    a[0] += 1
    a[1] += float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
tot = sum(v[1] for v in agg.values())
w = csv.writer(sys.stdout)
w.writerow(["kernel", "launches", "total_us", "share"])
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    w.writerow([k[:90], n, round(us, 1), round(us / tot, 4)])
