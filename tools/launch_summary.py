#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import csv, re, sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if r]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
ix = {h: i for i, h in enumerate(rows[hdr])}
agg, cnt = defaultdict(float), defaultdict(int)
for r in rows[hdr + 1:]:
    if len(r) <= ix["Metric Value"] or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]])[:110]
    agg[name] += v
    cnt[name] += 1
tot = sum(agg.values())
print("ncu --metrics gpu__time_duration.sum launch list (%d launches), aggregated by kernel" % sum(cnt.values()))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    print("%9.3f ms %5.1f%%  n=%4d  %s" % (v, 100 * v / tot, cnt[k], k))
print("total %.3f ms" % tot)
