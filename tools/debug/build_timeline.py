#!/usr/bin/env python
"""GPU timeline of one list rebuild (config C) from torch.profiler (CUPTI sees the library's kernels too):
kernel name, start, duration and the idle gap before it -- shows what the host read-backs cost.
  python tools/debug/build_timeline.py [--steps 2]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser(); ap.add_argument("--steps", type=int, default=2); ap.add_argument("--L", type=float, default=100.1)
a = ap.parse_args()
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from lj_gpu_b200 import LJContext, init_fcc
ctx = LJContext(0)
q = init_fcc(1.0, a.L); pn = len(q)
qh = np.zeros((pn, 4)); qh[:, :3] = q
qd = torch.from_numpy(qh).cuda(); pd = torch.zeros_like(qd)
pl = ctx.makepair(qd, tiles=True)
for _ in range(3):
    ctx.rebuild(qd, pl, tiles=True); ctx.force_loop(qd, pd, pl, loop=2)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    ctx.force_loop(qd, pd, pl, loop=a.steps)
    ctx.rebuild(qd, pl, tiles=True)
    ctx.force_loop(qd, pd, pl, loop=a.steps)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start; prev_end = t0; busy = 0.0; gaps = 0.0
for e in ev:
    s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
    gap = e.time_range.start - prev_end
    print("%9.1f us  +%7.1f  gap %6.1f  %s" % (s, d, gap, e.name[:70]))
    busy += d; gaps += max(gap, 0.0); prev_end = max(prev_end, e.time_range.end)
print("busy %.1f us, idle gaps %.1f us, span %.1f us" % (busy, gaps, prev_end - t0))
