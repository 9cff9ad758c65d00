#!/usr/bin/env python
"""Executed instructions and stall samples per CUDA source line from an ncu --set full --import-source on report.
   tools/debug/prof_lines.py rep.ncu-rep kernel_substring [top]"""
import csv, collections, subprocess, sys
rep, want = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file = cur_fn = hdr = None
agg = collections.defaultdict(lambda: [0, 0, ""])
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": cur_fn = r[1]; continue
    if r[0] == "Line No": hdr = r; ie = hdr.index("Instructions Executed"); si = hdr.index("# Samples"); continue
    if hdr and len(r) == len(hdr) and r[0].isdigit() and want in (cur_fn or ""):
        a = agg[(cur_file, int(r[0]))]; a[0] += int(r[ie] or 0); a[1] += int(r[si] or 0); a[2] = r[1].strip()[:100]
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print("total warp instructions", tot, "samples", ts)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("  %-16s:%4d %10d %5.1f%% samp %5.1f%%  %s" % (k[0], k[1], v[0], 100 * v[0] / tot, 100 * v[1] / max(ts, 1), v[2]))
