#!/usr/bin/env python
"""Cell-tile mirror: parity against the per-row kernel and timing (one GPU).

  python tools/celltile_check.py [--L 100.1] [--density 1.0] [--reps 20]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--L", type=float, default=100.1)
    ap.add_argument("--density", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--layouts", default="aos4")
    args = ap.parse_args()
    import numpy as np
    import torch

    from lj_gpu_b200 import LJContext, init_fcc

    ctx = LJContext(0)
    q = init_fcc(args.density, args.L)
    pn = q.shape[0]

    def timeit(fn, reps):
        fn(); fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    for layout in args.layouts.split(","):
        if layout == "aos3":
            qh = q.copy()
        elif layout == "aos4":
            qh = np.zeros((pn, 4)); qh[:, :3] = q
        else:
            qh = np.ascontiguousarray(q.T)
        qd = torch.from_numpy(qh).cuda()
        npn = pn if layout == "soa" else None
        pl = ctx.makepair(qd, layout=layout, pn=npn, tiles=True)
        P = pl.number_of_pairs
        plain = ctx.makepair(qd, layout=layout, pn=npn)  # drops the mirror: rebuilt right below
        rows = torch.repeat_interleave(torch.arange(pn, device="cuda"), plain.number_of_partners[:pn].long())
        ka = torch.sort(rows * pn + plain.sorted_list[:plain.number_of_pairs].long()).values
        kb = torch.sort(rows * pn + pl.sorted_list[:P].long()).values
        same = P == plain.number_of_pairs and bool((plain.number_of_partners == pl.number_of_partners).all()) \
            and bool((ka == kb).all())
        print("layout=%s list written by the tile fill pass == plain build: %s" % (layout, same), flush=True)
        del plain, rows, ka, kb
        ctx.rebuild(qd, pl, layout=layout, pn=npn, tiles=True)
        p_ref = torch.zeros_like(qd)
        ctx.force_step(qd, p_ref, pl, layout=layout, pn=npn, variant="subwarp", group=8)
        p_new = torch.zeros_like(qd)
        ctx.force_step(qd, p_new, pl, layout=layout, pn=npn, variant="celltile")
        torch.cuda.synchronize()
        err = (p_new - p_ref).abs().max().item()
        scale = p_ref.abs().max().item()
        p_ref_one = p_ref.clone()
        print("layout=%s N=%d P=%d  max|dp| = %.3e (scale %.3e, rel %.3e)" % (layout, pn, P, err, scale, err / scale), flush=True)
        ms_sub = timeit(lambda: ctx.force_step(qd, p_ref, pl, layout=layout, pn=npn, variant="subwarp", group=8), args.reps)
        ms_ct = timeit(lambda: ctx.force_step(qd, p_new, pl, layout=layout, pn=npn, variant="celltile"), args.reps)
        ms_auto = timeit(lambda: ctx.force_step(qd, p_new, pl, layout=layout, pn=npn), args.reps)
        ms_b0 = timeit(lambda: ctx.rebuild(qd, pl, layout=layout, pn=npn), 5)
        ms_b1 = timeit(lambda: ctx.rebuild(qd, pl, layout=layout, pn=npn, tiles=True), 5)
        print("  force: subwarp g8 %.4f ms, celltile %.4f ms (auto %.4f ms); build %.3f ms, +tiles %.3f ms"
              % (ms_sub, ms_ct, ms_auto, ms_b0, ms_b1), flush=True)
        # mixed precision on the same mirror (fixed-point records, FP32 pair arithmetic)
        p_mx = torch.zeros_like(qd)
        ctx.force_step(qd, p_mx, pl, layout=layout, pn=npn, variant="celltile", precision="mixed")
        torch.cuda.synchronize()
        err = (p_mx - p_ref_one).abs().max().item()
        print("  mixed celltile vs FP64: max|dp| = %.3e (rel %.3e of %.3e)" % (err, err / scale, scale), flush=True)
        ms_mx = timeit(lambda: ctx.force_step(qd, p_mx, pl, layout=layout, pn=npn, variant="celltile", precision="mixed"), args.reps)
        ms_mr = timeit(lambda: ctx.force_step(qd, p_mx, pl, layout=layout, pn=npn, variant="subwarp", group=4, precision="mixed"), args.reps)
        print("  mixed: celltile %.4f ms (consumers=%s), per-row g4 %.4f ms" % (ms_mx, os.environ.get("LJ_TILE_CONSUMERS", "default"), ms_mr), flush=True)


if __name__ == "__main__":
    main()
