#!/usr/bin/env python
"""Per-instruction view of an ncu --set full report (source page): memory instructions with
tag requests / sectors, and the top stall samples.   tools/ncu_source.py rep.ncu-rep [kernel#]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = txt.splitlines()
# find header line
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = [r for r in csv.reader(io.StringIO("\n".join(lines[start:]))) if not (r and r[0] == "Address" and r is not None and False)]
hdr = rows[0]; idx = {k: i for i, k in enumerate(hdr)}
tot = sum(int(r[idx["# Samples"]] or 0) for r in rows[1:] if len(r) == len(hdr))
print("total samples", tot)
print("%-6s %-58s %8s %9s %9s %9s %9s %s" % ("off", "sass", "samples", "inst_exec", "tagreq", "l2sect", "l2ideal", "top stalls"))
base = int(rows[1][0], 16)
stall_cols = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
for r in rows[1:]:
    if len(r) != len(hdr): continue
    s = int(r[idx["# Samples"]] or 0)
    mem = r[idx["Address Space"]] != "-"
    if s * 200 < tot and not mem: continue
    st = sorted(((int(r[idx[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print("%-6x %-58s %8d %9s %9s %9s %9s %s" % (int(r[0], 16) - base, r[1].strip()[:58], s, r[idx["Instructions Executed"]],
          r[idx["L1 Tag Requests Global"]], r[idx["L2 Theoretical Sectors Global"]], r[idx["L2 Theoretical Sectors Global Ideal"]],
          " ".join("%s:%d" % (n, v) for v, n in st if v)))
