"""compute-sanitizer target: every kernel family once on a small system (tools/gpu.sh sanitize)."""
import sys
sys.path.insert(0, ".")
import numpy as np
import torch
from lj_gpu_b200 import LJContext, PairList, init_fcc
SKIP_CELLTILE = "--skip-celltile" in sys.argv   # racecheck: see profiles/README.md (mbarrier pipelines)
ctx = LJContext(0)
q = init_fcc(0.8, 14.0); pn = len(q)
q4 = np.zeros((pn, 4)); q4[:, :3] = q
qd = torch.from_numpy(q4).cuda()
for kw in (dict(), dict(clusters=True), dict(per_particle=True), dict(half=True), dict(sort_rows=True, pointer64=True),
           dict(tiles=True), dict(tiles="wide", rows=(0, pn // 2))):
    pl = ctx.makepair(qd, **kw)
    pd = torch.zeros_like(qd)
    if kw.get("half"):
        ctx.force_loop(qd, pd, pl, loop=2, variant="n3", group=8)
        ctx.make_transposed_pairlist(pl)
        ctx.force_loop(qd, pd, pl, loop=2, ell=True)                # Newton-3 on the half ELL table
        continue
    rows = kw.get("rows")
    if kw.get("tiles") and SKIP_CELLTILE:
        continue
    if kw.get("tiles"):
        for prec in ("fp64", "mixed"):
            ctx.force_loop(qd, pd, pl, loop=2, variant="celltile", precision=prec, rows=rows)
            ctx.force_step(qd, pd, pl, variant="celltile", precision=prec, rows=rows, part="interior")
            ctx.force_step(qd, pd, pl, variant="celltile", precision=prec, rows=rows, part="boundary")
        ctx.rebuild(qd, pl)
        continue
    for variant, group, prec in (("subwarp", 8, "fp64"), ("subwarp", 32, "fp64"), ("subwarp", 1, "fp64"), ("tile", 8, "fp64"),
                                 ("tile", 32, "fp64"), ("subwarp", 4, "mixed")):
        ctx.force_loop(qd, pd, pl, loop=2, variant=variant, group=group, precision=prec)
    ctx.force_loop(qd, pd, pl, loop=2, variant="subwarp", group=8, list_scalar=2)
    if kw.get("clusters"):
        for g in (0, 16, 32):
            ctx.force_loop(qd, pd, pl, loop=2, variant="cluster", group=g)
        ctx.force_loop(qd, pd, pl, loop=2, variant="cluster", precision="mixed")
    ctx.make_transposed_pairlist(pl)
    ctx.force_loop(qd, pd, pl, loop=2, ell=True)
    ctx.make_sorted_list2d(pl)
    ctx.force_loop(qd, pd, pl, loop=2, ell_rows=True, group=8)
    ctx.random_shfl(pl); ctx.check_loadedpair(pl)
    ctx.list_mirror(qd, pl)                                          # mirror of a caller-supplied (shuffled) list
    if not SKIP_CELLTILE:
        ctx.force_loop(qd, pd, pl, loop=2, variant="celltile")
qf = torch.from_numpy(q.astype(np.float32)).cuda()                  # float3 layout
plf = ctx.makepair(qf)
ctx.force_loop(qf, torch.zeros_like(qf), plf, loop=2, precision="mixed")
torch.cuda.synchronize()
m = ctx.measure(q4.copy(), np.zeros_like(q4), layout="aos4", loop=5, rebuild_every=2, variant="subwarp", group=8)
print("sanitizer target done", pn, m.number_of_pairs)
