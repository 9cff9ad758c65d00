import sys; sys.path.insert(0,'.')
import numpy as np, torch
from lj_gpu_b200 import LJContext, init_fcc
ctx=LJContext(0); import os
q=init_fcc(1.0,float(os.environ.get('LJ_DIAG_L','100.1'))); pn=len(q)
q4=np.zeros((pn,4)); q4[:,:3]=q; qd=torch.from_numpy(q4).cuda(); pd=torch.zeros_like(qd)
pl=ctx.makepair(qd)
def t(fn,reps=10):
    fn();fn();torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/reps
for name,v in (("normal g8 scalar",1),("mem only",100),("math only (L1-resident gather)",101),("list loads only",102),("gather only (no list, no math)",103),("list via TEX pipe (experiment)",104),("positions via TEX pipe (experiment)",105)):
    print("%-34s %.4f ms"%(name,t(lambda: ctx.force_step(qd,pd,pl,variant=v,group=8,list_scalar=True))))
