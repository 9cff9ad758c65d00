#!/usr/bin/env python
"""List-build timing at config C (N = 1,000,188): ms per lj_build_list with LJ_LIST_TILES, CUDA events.
   python tools/debug/build_time.py [--reps 10] [--wide]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from lj_gpu_b200 import LJContext, init_fcc

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--wide", action="store_true")
ap.add_argument("--L", type=float, default=100.1)
ap.add_argument("--density", type=float, default=1.0)
a = ap.parse_args()
ctx = LJContext(0)
q = init_fcc(a.density, a.L)
q4 = np.zeros((len(q), 4)); q4[:, :3] = q
qd = torch.from_numpy(q4).cuda()
tiles = "wide" if a.wide else True
pl = ctx.makepair(qd, tiles=tiles)
for _ in range(3):
    ctx.rebuild(qd, pl, tiles=tiles)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    ctx.rebuild(qd, pl, tiles=tiles)
e1.record()
torch.cuda.synchronize()
print("N=%d pairs=%d  list build %.4f ms (%s tiles)" % (len(q), pl.number_of_pairs, e0.elapsed_time(e1) / a.reps, "wide" if a.wide else "standard"))
