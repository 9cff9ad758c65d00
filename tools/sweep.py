#!/usr/bin/env python
"""Time every force-kernel variant and the list build on one GPU (CUDA events, list >> L2).

  python tools/sweep.py [--L 100.1] [--density 1.0] [--reps 20] [--out gpurun_out/sweep.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--L", type=float, default=100.1)
    ap.add_argument("--density", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.json"))
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    import numpy as np
    import torch

    from bench import algorithmic_bytes, measured_peak_gbs
    from lj_gpu_b200 import LJContext, init_fcc

    ctx = LJContext(0)
    q = init_fcc(args.density, args.L)
    pn = q.shape[0]
    peak, _ = measured_peak_gbs()
    results = []

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

    def arrays(layout):
        if layout == "aos3":
            qh = q.copy()
        elif layout == "aos4":
            qh = np.zeros((pn, 4)); qh[:, :3] = q
        else:
            qh = np.ascontiguousarray(q.T)
        qd = torch.from_numpy(qh).cuda()
        return qd, torch.zeros_like(qd)

    for layout in (["aos4"] if args.quick else ["aos4", "aos3", "soa"]):
        qd, pd = arrays(layout)
        npn = pn if layout == "soa" else None
        for sort_rows in ([False] if (args.quick or layout != "aos4") else [False, True]):
            pl = ctx.makepair(qd, layout=layout, pn=npn, sort_rows=sort_rows, clusters=True)
            P = pl.number_of_pairs
            svec = 32 if layout == "aos4" else 24
            B = algorithmic_bytes(pn, P, svec)
            ms_build = timeit(lambda: ctx.rebuild(qd, pl, layout=layout, pn=npn, sort_rows=sort_rows, clusters=True), 5)
            plain = ctx.makepair(qd, layout=layout, pn=npn, sort_rows=sort_rows)
            ms_plain = timeit(lambda: ctx.rebuild(qd, plain, layout=layout, pn=npn, sort_rows=sort_rows), 5)
            ms_pp = timeit(lambda: ctx.rebuild(qd, plain, layout=layout, pn=npn, sort_rows=sort_rows, per_particle=True), 5)
            del plain
            print("layout=%s sort_rows=%d N=%d P=%d max_np=%d  list build %.3f ms (+cluster list: %.3f ms, "
                  "per-particle search: %.3f ms)" % (layout, sort_rows, pn, P, pl.max_partners, ms_plain, ms_build, ms_pp), flush=True)
            results.append(dict(kind="build", layout=layout, sort_rows=sort_rows, ms=ms_build, pn=pn, pairs=P))
            cases = [("cluster", 0, 0, "fp64"), ("cluster", 0, 256, "fp64"), ("cluster", 32, 0, "fp64"), ("cluster", 16, 0, "fp64")]
            for g in (1, 2, 4, 8, 16, 32):
                for tb in ((128, 256) if layout == "aos4" and not sort_rows else (128,)):
                    cases.append(("subwarp", g, tb, "fp64"))
            for g in (4, 8, 16, 32):
                cases.append(("tile", g, 0, "fp64"))
            for g in (4, 8, 16):
                cases.append(("subwarp", g, 128, "fp64-int4-list"))
            for g in (4, 8, 16, 32):
                cases.append(("subwarp", g, 128, "mixed"))
            cases.append(("cluster", 0, 0, "mixed"))
            for variant, g, tb, prec in cases:
                kw = dict(layout=layout, pn=npn, variant=variant, group=g, threads_per_block=tb,
                          precision=prec.split("-")[0], list_scalar=2 if prec.endswith("int4-list") else 0)
                try:
                    ms = timeit(lambda: ctx.force_step(qd, pd, pl, **kw), args.reps)
                except Exception as e:  # noqa: BLE001
                    print("  %-8s g=%-2d tb=%-4d %-5s FAILED %s" % (variant, g, tb, prec, e), flush=True)
                    continue
                gbs = B / ms / 1e6
                print("  %-8s g=%-2d tb=%-4d %-5s %8.4f ms  %8.3e pairs/s  %7.1f GB/s alg  %5.1f%% of HBM peak" % (
                    variant, g, tb, prec, ms, P / ms * 1e3, gbs, 100 * gbs / peak), flush=True)
                results.append(dict(kind="force", layout=layout, sort_rows=sort_rows, variant=variant, group=g,
                                    tb=tb, prec=prec, ms=ms, pairs_per_s=P / ms * 1e3, gbs=gbs, frac=gbs / peak))
            if layout == "aos4" and not sort_rows and not args.quick:
                tl = ctx.make_transposed_pairlist(pl)
                ms = timeit(lambda: ctx.force_step(qd, pd, pl, ell=True), args.reps)
                print("  ell thread-per-i           %8.4f ms  %8.3e pairs/s" % (ms, P / ms * 1e3), flush=True)
                results.append(dict(kind="force", layout=layout, variant="ell", group=1, ms=ms, pairs_per_s=P / ms * 1e3))
                del tl
                pl.transposed_list = None
                half = ctx.makepair(qd, half=True)
                for g in (8, 32):
                    ms = timeit(lambda: ctx.force_step(qd, pd, half, variant="n3", group=g), 5)
                    print("  newton3 g=%-2d half list     %8.4f ms  %8.3e directed-pair-equivalents/s" % (
                        g, ms, 2 * half.number_of_pairs / ms * 1e3), flush=True)
                    results.append(dict(kind="force", layout=layout, variant="n3", group=g, ms=ms,
                                        pairs_per_s=2 * half.number_of_pairs / ms * 1e3))
                del half
            del pl
        del qd, pd
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(results, f, indent=1)
    ctx.close()


if __name__ == "__main__":
    main()
