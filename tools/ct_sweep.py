#!/usr/bin/env python
"""Cell-tile force kernel: timing sweep over the diagnostic knobs of a `make DIAG=1` build (one GPU).

  python tools/ct_sweep.py [--reps 20] [--configs "LJ_TILE_WPG=16;LJ_TILE_CONSUMERS=24;..."]

Each configuration is a comma-separated list of NAME=VALUE environment settings (read per launch by
the DIAG build).  Every configuration is first checked bit for bit against the per-row kernel
(FP64) / within 1e-5 of it (mixed), then timed with CUDA events.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

KNOBS = ("LJ_TILE_NG", "LJ_TILE_WPG", "LJ_TILE_CONSUMERS", "LJ_TILE_RY", "LJ_TILE_RL", "LJ_TILE_SEG", "LJ_TILE_MODE",
         "LJ_TILE_ROWS")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--L", type=float, default=100.1)
    ap.add_argument("--density", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--configs", default=";LJ_TILE_WPG=16")
    ap.add_argument("--prec", default="fp64,mixed")
    ap.add_argument("--check-steps", type=int, default=30, help="steps of the bit-for-bit check against the per-row kernel")
    ap.add_argument("--rows", default="", help="comma list of LJ_TILE_ROWS values to rebuild the mirror with")
    args = ap.parse_args()
    import numpy as np
    import torch

    from lj_gpu_b200 import LJContext, init_fcc

    ctx = LJContext(0)
    q = init_fcc(args.density, args.L)
    pn = q.shape[0]
    qh = np.zeros((pn, 4)); qh[:, :3] = q
    qd = torch.from_numpy(qh).cuda()

    def timeit(fn, reps):
        fn(); fn(); fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def setenv(cfg):
        for k in KNOBS:
            os.environ.pop(k, None)
        for kv in filter(None, cfg.split(",")):
            k, v = kv.split("=")
            os.environ[k] = v

    for rows in (args.rows.split(",") if args.rows else [""]):
        for prec in args.prec.split(","):
            wide = rows == "wide"            # LJ_LIST_TILES_WIDE (product builds too); numbers need DIAG=1
            if wide:
                rows = ""
            setenv("LJ_TILE_ROWS=%s" % rows if rows else "")
            pl = ctx.makepair(qd, tiles="wide" if wide or (prec == "mixed" and not rows) else True)
            P = pl.number_of_pairs
            p_ref = torch.zeros_like(qd)
            ctx.force_loop(qd, p_ref, pl, loop=args.check_steps, variant="subwarp", group=8)
            scale = p_ref.abs().max().item()
            algo = 4.0 * P + 104.0 * pn
            for cfg in args.configs.split(";"):
                setenv(cfg + (",LJ_TILE_ROWS=%s" % rows if rows else ""))
                p_new = torch.zeros_like(qd)
                ctx.force_loop(qd, p_new, pl, loop=args.check_steps, variant="celltile", precision=prec)
                torch.cuda.synchronize()
                err = (p_new - p_ref).abs().max().item() / scale
                ok = (err == 0.0) if prec == "fp64" else (0 < err < 1e-5)
                ms = timeit(lambda: ctx.force_step(qd, p_new, pl, variant="celltile", precision=prec), args.reps)
                print("rows=%-4s %-5s %-44s %.4f ms  %5.1f %% of 6535 GB/s  err %.2e %s" % (
                    "wide" if wide else (rows or "def"), prec, cfg or "(default)", ms, 100 * algo / (ms * 1e-3) / 6535.1e9, err,
                    "ok" if ok else "MISMATCH"), flush=True)
            del pl
            if wide:
                rows = "wide"


if __name__ == "__main__":
    main()
