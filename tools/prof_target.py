#!/usr/bin/env python
"""Small profiling target for ncu: config C (N=1M, rho=1.0), a few launches of chosen kernels.
  python tools/prof_target.py --variant tile --group 8 --steps 3 [--rebuild 1] [--prec fp64]
"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--variant", default="auto"); ap.add_argument("--group", type=int, default=0)
ap.add_argument("--steps", type=int, default=3); ap.add_argument("--rebuild", type=int, default=0)
ap.add_argument("--prec", default="fp64"); ap.add_argument("--L", type=float, default=100.1)
ap.add_argument("--density", type=float, default=1.0); ap.add_argument("--sort-rows", action="store_true")
ap.add_argument("--tb", type=int, default=0); ap.add_argument("--layout", default="aos4")
ap.add_argument("--wide", action="store_true", help="mirror with the wide tiles of the mixed kernel")
a = ap.parse_args()
import numpy as np, torch
from lj_gpu_b200 import LJContext, init_fcc
ctx = LJContext(0)
q = init_fcc(a.density, a.L); pn = len(q)
if a.layout == "aos4":
    qh = np.zeros((pn, 4)); qh[:, :3] = q
elif a.layout == "aos3":
    qh = q.copy()
else:
    qh = np.ascontiguousarray(q.T)
npn = pn if a.layout == "soa" else None
qd = torch.from_numpy(qh).cuda(); pd = torch.zeros_like(qd)
tiles = a.variant in ("celltile", "auto")
if tiles and a.wide:
    tiles = "wide"
pl = ctx.makepair(qd, layout=a.layout, pn=npn, sort_rows=a.sort_rows, clusters=(a.variant == "cluster"), tiles=tiles)
for _ in range(a.rebuild):
    ctx.rebuild(qd, pl, layout=a.layout, pn=npn, sort_rows=a.sort_rows, clusters=(a.variant == "cluster"), tiles=tiles)
for _ in range(a.steps):
    ctx.force_step(qd, pd, pl, layout=a.layout, pn=npn, variant=a.variant, group=a.group, precision=a.prec, threads_per_block=a.tb)
torch.cuda.synchronize()
print("done pn=%d pairs=%d" % (pn, pl.number_of_pairs))
