import os, sys
sys.path.insert(0, ".")
import numpy as np, torch
from lj_gpu_b200 import LJContext, init_fcc
ctx = LJContext(0)
q = init_fcc(1.0, 100.1); pn = len(q)
qh = np.zeros((pn, 4)); qh[:, :3] = q
qd = torch.from_numpy(qh).cuda()
pl = ctx.makepair(qd, tiles=True)
ref = torch.zeros_like(qd)
ctx.force_loop(qd, ref, pl, loop=100, variant="subwarp", group=8)
for wpg in ("16", "8", "4"):
    os.environ["LJ_TILE_WPG"] = wpg
    a = torch.zeros_like(qd); b = torch.zeros_like(qd); c = torch.zeros_like(qd)
    ctx.force_loop(qd, a, pl, loop=100, variant="celltile")
    ctx.force_loop(qd, b, pl, loop=100, variant="celltile", use_graph=True)
    ctx.force_loop(qd, c, pl, loop=100, variant="celltile", use_graph=True)
    torch.cuda.synchronize()
    print("WPG", wpg, "direct==ref", torch.equal(a, ref), "graph==ref", torch.equal(b, ref), "graph2==ref", torch.equal(c, ref),
          "maxdiff", (b - ref).abs().max().item() / ref.abs().max().item(), (c - ref).abs().max().item() / ref.abs().max().item(), flush=True)
    ctx.list_invalidate() if False else None
