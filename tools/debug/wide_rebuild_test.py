#!/usr/bin/env python
"""Rebuild timings around a switch of tile width / precision (the sequence of bench.py's mixed-precision block)."""
import numpy as np, torch, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lj_gpu_b200 import LJContext, init_fcc
ctx = LJContext(0)
q = init_fcc(1.0, 100.1); q4 = np.zeros((len(q), 4)); q4[:, :3] = q
qd = torch.from_numpy(q4).cuda(); pd = torch.zeros_like(qd)
pl = ctx.makepair(qd, tiles=True)
def t(label, fn):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0 = time.perf_counter(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    print("%-28s gpu %.3f ms  host %.3f ms" % (label, e0.elapsed_time(e1), 1e3 * (time.perf_counter() - h0)), flush=True)
for k in range(2): t("rebuild 40-row", lambda: ctx.rebuild(qd, pl, tiles=True))
ctx.force_loop(qd, pd, pl, loop=20)
pm, pf = torch.zeros_like(qd), torch.zeros_like(qd)
ctx.force_loop(qd, pf, pl, loop=20)
t("first rebuild wide", lambda: ctx.rebuild(qd, pl, tiles="wide"))
t("first mixed loop", lambda: ctx.force_loop(qd, pm, pl, loop=20, precision="mixed"))
t("mixed loop", lambda: ctx.force_loop(qd, pm, pl, loop=20, precision="mixed"))
for k in range(5): t("rebuild wide after mixed", lambda: ctx.rebuild(qd, pl, tiles="wide"))
t("mixed loop", lambda: ctx.force_loop(qd, pm, pl, loop=20, precision="mixed"))
for k in range(3): t("rebuild wide after mixed", lambda: ctx.rebuild(qd, pl, tiles="wide"))
for k in range(3): t("rebuild 40-row again", lambda: ctx.rebuild(qd, pl, tiles=True))
