#!/usr/bin/env python
"""Instructions executed and stall samples of an ncu --set full report, aggregated by SOURCE LINE
(needs --import-source on and -lineinfo).   tools/ncu_lines.py rep.ncu-rep [top]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = txt.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
hdr = rows[0]; idx = {k: i for i, k in enumerate(hdr)}
print([k for k in hdr][:12])
