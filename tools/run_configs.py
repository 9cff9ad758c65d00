#!/usr/bin/env python
"""BASELINE.json configs 4 and 5 on ONE GPU (evidence runs, results as JSON lines):

  --config D    synthetic FCC N=16,078,716 rho=1.0 (159 cells/side): FP64 gather vs FP32-mixed vs Newton-3
                scatter, 64-bit pointer[] (2.2e9 list entries > 2^31), sampled parity against numpy FP64
  --config E1   synthetic FCC N=131,072,000 rho=0.8 (320 cells/side) on one GPU: the strong-scaling
                denominator for the 8-GPU decomposed run (bench.py under torchrun with LJ_BENCH_CELLS=320)
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from bench import algorithmic_bytes, measured_peak_gbs
from lj_gpu_b200 import LJContext, init_fcc
from lj_gpu_b200.decomp import lattice_spacing


def timeit(fn, reps):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def sampled_parity(q, p_dev, pl, steps, n_sample=64, seed=3):
    """FP64 numpy recomputation of `steps` force steps for a few rows, using the GPU list rows, plus a
    brute-force membership check of those rows against all N particles."""
    pn = len(q)
    rng = np.random.RandomState(seed)
    rows = np.sort(rng.choice(pn, n_sample, replace=False))
    nop = pl.number_of_partners[torch.from_numpy(rows).cuda()].cpu().numpy()
    ptr = pl.pointer[torch.from_numpy(rows).cuda()].cpu().numpy().astype(np.int64)
    worst, list_ok = 0.0, True
    pmax = float(p_dev[:, :3].abs().max().item())
    for r, n, o in zip(rows, nop, ptr):
        js = pl.sorted_list[o:o + n].cpu().numpy().astype(np.int64)
        d = q[js] - q[r]
        r2 = d[:, 2] * d[:, 2] + (d[:, 1] * d[:, 1] + d[:, 0] * d[:, 0])
        r6 = r2 ** 3
        df = np.where(r2 > 9.0, 0.0, (24.0 * r6 - 48.0) / (r6 * r6 * r2) * 0.001)
        want = steps * (df[:, None] * d).sum(0)
        got = p_dev[r, :3].cpu().numpy()
        worst = max(worst, np.abs(got - want).max() / pmax)
        # membership: every particle within 3.3 (strict) must be listed (full list)
        if not pl.half:
            dd = q - q[r]
            near = np.flatnonzero((dd * dd).sum(1) < 3.3 * 3.3 - 1e-9)
            near = near[near != r]
            list_ok &= set(near.tolist()) <= set(js.tolist()) and len(js) <= len(near) + 2
    return worst, bool(list_ok)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, choices=["D", "E1", "C"])
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    rho, cells = {"C": (1.0, 63), "D": (1.0, 159), "E1": (0.8, 320)}[a.config]
    L = (cells + 0.05) * lattice_spacing(rho)
    ctx = LJContext(0)
    peak, _ = measured_peak_gbs()
    t0 = time.time()
    q = init_fcc(rho, L)
    pn = len(q)
    print(json.dumps({"config": a.config, "N": pn, "rho": rho, "cells": cells, "generator_s": time.time() - t0}), flush=True)
    q4 = np.zeros((pn, 4)); q4[:, :3] = q
    qd = torch.from_numpy(q4).cuda(); del q4
    pd = torch.zeros_like(qd)
    pl = ctx.makepair(qd, pointer64=True)
    P = pl.number_of_pairs
    ms_build = timeit(lambda: ctx.rebuild(qd, pl), 3)
    out = {"config": a.config, "N": pn, "pairs_full": P, "max_partners": pl.max_partners, "pointer": "int64",
           "list_build_ms": ms_build, "list_GB": 4 * P / 1e9}
    for name, kw in (("fp64_gather_g8", dict(variant="subwarp", group=8)),
                     ("mixed_g4", dict(variant="subwarp", group=4, precision="mixed"))):
        if a.config == "E1" and name != "fp64_gather_g8":
            continue
        ms = timeit(lambda: ctx.force_step(qd, pd, pl, **kw), a.reps)
        B = algorithmic_bytes(pn, P, 32, 8)
        out[name] = {"ms_per_step": ms, "pairs_per_s": P / ms * 1e3, "roofline_frac": B / ms / 1e6 / peak,
                     "amortised_ms_rebuild_every_20": ms + ms_build / 20}
    # parity on sampled rows: 3 fresh steps
    pd.zero_()
    ctx.force_loop(qd, pd, pl, loop=3, variant="subwarp", group=8)
    err, ok = sampled_parity(q, pd, pl, 3)
    out["fp64_sampled_rel_err"], out["list_sample_ok"] = err, ok
    if a.config == "D":
        pd.zero_()
        ctx.force_loop(qd, pd, pl, loop=3, variant="subwarp", group=4, precision="mixed")
        out["mixed_sampled_rel_err"], _ = sampled_parity(q, pd, pl, 3)
        del pl
        torch.cuda.empty_cache()
        half = ctx.makepair(qd, half=True, pointer64=True)
        ms = timeit(lambda: ctx.force_step(qd, pd, half, variant="n3", group=8), 3)
        out["newton3_g8_half_list"] = {"ms_per_step": ms, "pairs_half": half.number_of_pairs,
                                       "directed_pair_equivalents_per_s": 2 * half.number_of_pairs / ms * 1e3}
        pd.zero_()
        ctx.force_loop(qd, pd, half, loop=3, variant="n3", group=8)
        # Newton-3 result must equal the gather result: compare via total momentum ~ 0 and sampled rows
        out["newton3_total_momentum_over_pmax"] = float(pd[:, :3].sum(0).abs().max().item() / pd[:, :3].abs().max().item())
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
