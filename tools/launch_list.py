#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: launch_list.py <launches.csv> ["header line" ...]"""
import csv, sys, collections, re
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("mp2::", "")
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
    tot[name] += v; cnt[name] += 1
for h in sys.argv[2:]:
    print(h)
s = sum(tot.values())
for k, v in tot.most_common():
    print(f"{k:60s} launches {cnt[k]:4d}  total {v:10.3f} ms  avg {v / cnt[k] * 1e3:10.1f} us  share {v / s * 100:5.1f}%")
