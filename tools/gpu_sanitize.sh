#!/bin/bash
# compute-sanitizer on the small smoke configuration (every kernel family) + one mid-size parity test
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
cat > /tmp/san.py <<'PY'
import sys; sys.path.insert(0, ".")
import numpy as np, torch
from lj_gpu_b200 import LJContext, init_fcc
ctx = LJContext(0)
q = init_fcc(0.8, 14.0); pn = len(q)
q4 = np.zeros((pn, 4)); q4[:, :3] = q
qd = torch.from_numpy(q4).cuda()
for kw in (dict(), dict(clusters=True), dict(per_particle=True), dict(half=True), dict(sort_rows=True, pointer64=True)):
    pl = ctx.makepair(qd, **kw)
    pd = torch.zeros_like(qd)
    if kw.get("half"):
        ctx.force_loop(qd, pd, pl, loop=2, variant="n3", group=8)
        continue
    for variant, group, prec in (("subwarp", 8, "fp64"), ("subwarp", 32, "fp64"), ("subwarp", 1, "fp64"), ("tile", 8, "fp64"),
                                 ("tile", 32, "fp64"), ("subwarp", 4, "mixed")):
        ctx.force_loop(qd, pd, pl, loop=2, variant=variant, group=group, precision=prec)
    ctx.force_loop(qd, pd, pl, loop=2, variant="subwarp", group=8, list_scalar=2)
    if kw.get("clusters"):
        for g in (0, 16, 32):
            ctx.force_loop(qd, pd, pl, loop=2, variant="cluster", group=g)
        ctx.force_loop(qd, pd, pl, loop=2, variant="cluster", precision="mixed")
    tl = ctx.make_transposed_pairlist(pl)
    ctx.force_loop(qd, pd, pl, loop=2, ell=True)
    ctx.random_shfl(pl); ctx.check_loadedpair(pl)
torch.cuda.synchronize()
m = ctx.measure(q4.copy(), np.zeros_like(q4), layout="aos4", loop=5, rebuild_every=2)
print("sanitizer target done", pn, m.number_of_pairs)
PY
timeout -s KILL 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python /tmp/san.py > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck.log
timeout -s KILL 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python /tmp/san.py > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/racecheck.log
