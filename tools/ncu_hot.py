#!/usr/bin/env python
"""Hot instructions (stall samples) of the first kernel in an ncu --set full report.
  python tools/ncu_hot.py report.ncu-rep"""
import csv, subprocess, io, sys
rep=sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = txt.splitlines()
st=[i for i,l in enumerate(lines) if l.startswith('"Address"')][0]
rows=list(csv.reader(io.StringIO("\n".join(lines[st:]))))
hdr=rows[0]; idx={k:i for i,k in enumerate(hdr)}
data=[]
for r in rows[1:]:
    if r and r[0]=="Address": break  # next kernel of the report
    if len(r)==len(hdr): data.append(r)
tot=sum(int(r[idx["# Samples"]]) for r in data)
base=int(data[0][0],16)
stall_cols=[k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
agg={c:sum(int(r[idx[c]] or 0) for r in data) for c in stall_cols}
print("total samples",tot, {k[6:]:v for k,v in sorted(agg.items(), key=lambda kv:-kv[1])[:7]})
for r in data:
    s=int(r[idx["# Samples"]])
    if s*100>=tot*1.5:
        stl=sorted(((int(r[idx[c]] or 0),c[6:]) for c in stall_cols),reverse=True)[:3]
        print("%5x %-58s %6d %5.1f%% exec=%-9s %s"%(int(r[0],16)-base, r[1].strip()[:58], s, 100*s/tot, r[idx["Instructions Executed"]], " ".join("%s:%d"%(n,v) for v,n in stl if v)))
