#!/usr/bin/env python
"""BASELINE config 5 (N = 131,072,000) on ONE GPU: first build, force steps, one rebuild, memory high-water mark.
   python tools/debug/config5_one_gpu.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from lj_gpu_b200 import LJContext, init_fcc
from lj_gpu_b200.decomp import lattice_spacing
ctx = LJContext(0)
t0 = time.perf_counter()
q5 = init_fcc(0.8, (320 + 0.05) * lattice_spacing(0.8)); n5 = len(q5)
q54 = np.zeros((n5, 4)); q54[:, :3] = q5; del q5
qd = torch.from_numpy(q54).cuda(); del q54
pd = torch.zeros_like(qd)
pl = ctx.makepair(qd, pointer64=True, tiles=True)
torch.cuda.synchronize()
print("N=%d pairs=%d setup %.1f s, free %.1f GB" % (n5, pl.number_of_pairs, time.perf_counter() - t0, torch.cuda.mem_get_info()[0] / 2**30), flush=True)
ctx.force_loop(qd, pd, pl, loop=2)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
e[0].record(); ctx.force_loop(qd, pd, pl, loop=10); e[1].record(); ctx.rebuild(qd, pl); e[2].record(); ctx.rebuild(qd, pl); e[3].record()
torch.cuda.synchronize()
print("force %.2f ms/step, rebuild %.1f ms, again %.1f ms, free %.1f GB" % (e[0].elapsed_time(e[1]) / 10, e[1].elapsed_time(e[2]), e[2].elapsed_time(e[3]), torch.cuda.mem_get_info()[0] / 2**30))
# sampled rows of the list against a brute-force search on the host
rng = np.random.RandomState(5); rows = rng.randint(0, n5, 16)
qh = qd[:, :3].cpu().numpy()
nop = pl.number_of_partners[torch.from_numpy(rows).cuda()].cpu().numpy()
for r, c in zip(rows, nop):
    d2 = ((qh - qh[r]) ** 2).sum(1); exp = int((d2 < 3.3 * 3.3).sum()) - 1
    assert exp == c, (r, exp, c)
print("16 sampled row lengths match brute force")
