"""debug: where does the mixed cell-tile kernel produce NaN?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from lj_gpu_b200 import LJContext, init_fcc
L = float(sys.argv[1]) if len(sys.argv) > 1 else 48.0
ctx = LJContext(0)
q = init_fcc(1.0, L); pn = len(q)
q4 = np.zeros((pn, 4)); q4[:, :3] = q
qd = torch.from_numpy(q4).cuda()
pl = ctx.makepair(qd, tiles=True)
pf = torch.zeros_like(qd); ctx.force_step(qd, pf, pl, variant="celltile")
for trial in range(2):
    pm = torch.zeros_like(qd); ctx.force_step(qd, pm, pl, variant="celltile", precision="mixed")
    torch.cuda.synchronize()
    bad = torch.isnan(pm[:, :3]).any(dim=1) | torch.isinf(pm[:, :3]).any(dim=1)
    nb = int(bad.sum())
    ok = ~bad
    err = ((pm - pf)[ok][:, :3].abs().max() / pf[:, :3].abs().max()).item()
    print("trial %d: pn=%d bad rows=%d, rel err on the others %.3e" % (trial, pn, nb, err))
    if nb:
        idx = torch.nonzero(bad).flatten()[:12].cpu().numpy()
        nop = pl.number_of_partners.cpu().numpy()
        for i in idx:
            print("  row %d q=%s nop=%d p=%s" % (i, q[i], nop[i], pm[i, :3].cpu().numpy()))
        qa = q[torch.nonzero(bad).flatten().cpu().numpy()]
        edge = 1.65 * (1 + 1e-9); lo = q.min(axis=0)
        cells = np.floor((qa - lo) / edge).astype(int)
        import collections
        cnt = collections.Counter((c[1], c[2], c[0] // 8) for c in cells)
        print("  bad tiles (cy, cz, tx): rows", sorted(cnt.items())[:40])
        print("  bad rows: x range %.2f..%.2f y %.2f..%.2f z %.2f..%.2f" % (qa[:,0].min(), qa[:,0].max(), qa[:,1].min(), qa[:,1].max(), qa[:,2].min(), qa[:,2].max()))
    big = (pm - pf)[:, :3].abs().max(dim=1).values
    big[bad] = 0
    w = torch.argsort(big, descending=True)[:5].cpu().numpy()
    for i in w:
        print("  worst finite row %d q=%s dp=%.3e" % (i, q[i], big[i].item()))
