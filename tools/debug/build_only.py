import sys
sys.path.insert(0, ".")
import numpy as np, torch
from lj_gpu_b200 import LJContext, init_fcc
ctx = LJContext(0)
q = init_fcc(1.0, 100.1); pn = len(q)
qh = np.zeros((pn, 4)); qh[:, :3] = q
qd = torch.from_numpy(qh).cuda()
pl = ctx.makepair(qd, tiles=True)
ctx.rebuild(qd, pl)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ctx.rebuild(qd, pl)
e1.record(); torch.cuda.synchronize()
print("rebuild ms", e0.elapsed_time(e1) / 5)
