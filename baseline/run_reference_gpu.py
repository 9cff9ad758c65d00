#!/usr/bin/env python
"""Run the REFERENCE's own CUDA program, recompiled for sm_100 by baseline/Makefile, on this box's GPU.

  python baseline/run_reference_gpu.py [--density 0.5 --density 1.0] [--json out.json]

For each density: (1) `gpu_cuda_test_d<rho>.out` (-DEN_TEST_GPU: warp_unroll2 only) -- its stdout
must be byte-identical to ref_data/density<rho>.dat (committed as tests/golden/), and it leaves the
reference's pair-list cache `.cache_pair_all.dat` in the scratch directory so that the O(N^2)
host makepair() runs once; (2) `gpu_cuda_d<rho>.out`: the 14 full-list kernels x {double3, double4},
LOOP = 100 steps each; the "without Host<->Device" seconds the reference prints
(cuda/force_cuda.cu:341) are parsed into {kernel: seconds per 100 steps}.

Nothing here is the product: this is the on-box GPU baseline bench.py quotes (`reference_gpu`).
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")
GOLDEN = {0.5: "density0.5.dat", 1.0: "density1.dat"}
PAIRS_FULL = {0.5: 4536276, 1.0: 15679772}      # directed pairs of configs A and B
LINE = re.compile(r"N=(\d+), (\S+) ([0-9.eE+-]+) \[sec\] \(without Host<->Device\)")


def exe(kind: str, density: float) -> str:
    return os.path.join(REF, "gpu_cuda%s_d%.1f.out" % (kind, density))


def available(density: float) -> bool:
    return os.path.exists(exe("", density)) and os.path.exists(exe("_test", density))


def run_density(density: float, timeout: float = 600.0, thread_block: int = 128) -> dict:
    out = {"density": density, "thread_block": thread_block}
    with tempfile.TemporaryDirectory(prefix="ljref_gpu_") as tmp:
        r = subprocess.run([exe("_test", density), str(thread_block)], cwd=tmp, capture_output=True, text=True,
                           timeout=timeout)
        if r.returncode != 0:
            return {"density": density, "error": "test build rc=%d: %s" % (r.returncode, r.stderr[-300:])}
        with open(os.path.join(ROOT, "tests", "golden", GOLDEN[density])) as f:
            out["goldens_reproduced"] = r.stdout == f.read()
        r = subprocess.run([exe("", density), str(thread_block)], cwd=tmp, capture_output=True, text=True,
                           timeout=timeout)
        if r.returncode != 0:
            return {"density": density, "error": "rc=%d: %s" % (r.returncode, r.stderr[-300:])}
        rows = {}
        for m in LINE.finditer(r.stderr):
            out["N"] = int(m.group(1))
            rows[m.group(2).replace("force_kernel_", "")] = float(m.group(3))
        out["seconds_per_100_steps"] = rows
        if rows:
            best = min(rows, key=rows.get)
            out["best_kernel"] = best
            out["best_seconds_per_100_steps"] = rows[best]
            out["best_pairs_per_s"] = PAIRS_FULL[density] * 100 / rows[best]
        out["list_cache_loaded"] = "is successfully loaded" in r.stderr
    return out


def reference_gpu(densities=(0.5, 1.0), timeout: float = 600.0) -> dict:
    res = {"what": "reference cuda/force_cuda.cu + kernel.cuh recompiled for sm_100 (baseline/Makefile: two shims, "
                   "nvcc -O3 -arch=sm_100), run on this GPU: seconds per LOOP=100 steps, 'without Host<->Device' "
                   "as the reference prints it", "rows": []}
    for d in densities:
        if not available(d):
            res["rows"].append({"density": d, "unavailable": "baseline/_ref binaries not built (run make -C baseline "
                                                              "where /root/reference exists)"})
            continue
        try:
            res["rows"].append(run_density(d, timeout))
        except subprocess.TimeoutExpired:
            res["rows"].append({"density": d, "error": "timeout"})
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--density", type=float, action="append")
    ap.add_argument("--json")
    a = ap.parse_args()
    r = reference_gpu(tuple(a.density) if a.density else (0.5, 1.0))
    s = json.dumps(r, indent=1)
    print(s)
    if a.json:
        with open(a.json, "w") as f:
            f.write(s)
    sys.exit(0)
