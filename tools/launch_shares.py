#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file F):
   python tools/launch_shares.py F"""
import csv, sys, collections
rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
rd = list(csv.reader(rows)); hdr = rd[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.defaultdict(float); cnt = collections.Counter()
for r in rd[1:]:
    v = float(r[iv].replace(",", "")); u = r[iu]
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(u, 1.0)
    k = r[ik].split("(")[0][:62]
    tot[k] += v; cnt[k] += 1
s = sum(tot.values())
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print("%-62s n=%4d %11.1f us %6.1f%%" % (k, cnt[k], v, 100 * v / s))
