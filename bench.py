#!/usr/bin/env python
"""bench.py -- headline benchmark of the LJ force + neighbour-list hot path on B200.

Contract (one JSON line on stdout, rank 0):
  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path

Workload (BASELINE.json configs[2]): synthetic jittered FCC lattice, rho = 1.0, 63 cells per
side -> N = 1,000,188 atoms, cutoff 3.0 sigma, search 3.3 sigma, FP64, full (directed) Verlet
list rebuilt ON THE GPU every 20 steps.  A "step" is one force/momentum-update pass over the
whole list; every 20th step also pays one list rebuild.  metric = directed pair interactions
per second = (list entries consumed per step) * steps / time.

  value      inputs resident in HBM, CUDA-event timed on the launching stream.
  e2e        the same metric through the reference-facing call lj_measure() (the reference's
             measure(): upload q,p from pinned HOST buffers -> list build -> K steps ->
             download p), wall clock of the call, copies inside the timed region.
  roofline   force kernel alone: algorithmic bytes B = 4*P + N*(S_q + 2*S_p + 8|12) per launch
              / mean launch duration measured live with CUDA events; peak from
             MEASURED_PEAKS.json.
  cpu_baseline  the REAL reference (cpu_ref/force_soa.cpp via oracle/_ref) on this box's host,
             1 core (the reference is serial), on a bounded sample (the workload's density, L=50).
  config     the workload as a function of the command line alone (bench_config): `--impl reference`
             prints the same `config` and states the bounded sample it timed in `sample`.

With N > 1 (torchrun) the lattice is split into z-slabs, one per rank; ghost positions are
exchanged every step over NVLink (see lj_gpu_b200/decomp.py) and value aggregates all ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

REBUILD_EVERY = 20
METRIC = "pair_interactions_per_s"
UNIT = "pairs/s"
CONFIG5_DENSITY = 0.8   # BASELINE config 5 (the N > 1 workload): FCC rho = 0.8, 320 cells per side


def fcc_cells(density, L):
    """init() of the reference driver (cuda/force_cuda.cu:47-94, lj_host.cpp): n = int(L / s) cells per
    side with s = (rho / 4)^(-1/3), four atoms per cell."""
    s = 1.0 / (density * 0.25) ** (1.0 / 3.0)
    n = int(L / s)
    return n, 4 * n ** 3


def bench_config(args, n_gpus):
    """`config` of the JSON line: the WORKLOAD, a function of the command line alone, so that both arms
    (`--impl cuda` and `--impl reference`) name the same one.  What a run measured on it (pairs per step,
    slab sizes, the kernel AUTO picked, the bounded sample the CPU arm ran) is stated outside `config`."""
    if n_gpus <= 1:
        n, pn = fcc_cells(args.density, args.L)
        return {"workload": "synthetic FCC lattice N=%d (%d cells/side) rho=%.1f cutoff=3.0 search=3.3, full "
                            "(directed) neighbour list, one force step per step, list rebuild every %d steps" % (
                                pn, n, args.density, REBUILD_EVERY),
                "l2": "inputs larger than L2 (the neighbour list alone, 4 B per directed pair, see l2_detail)",
                "rebuild_every": REBUILD_EVERY}
    cells5 = int(os.environ.get("LJ_BENCH_CELLS", "320"))     # 320 -> N = 131,072,000
    return {"workload": "BASELINE config 5: synthetic FCC lattice N=%d (%d cells/side) rho=%.1f cutoff=3.0 "
                        "search=3.3, full (directed) neighbour list, %d z-slabs, ghost positions exchanged every "
                        "step, list rebuild every %d steps" % (4 * cells5 ** 3, cells5, CONFIG5_DENSITY, n_gpus,
                                                               REBUILD_EVERY),
            "parallelism": "z-slab x%d" % n_gpus, "l2": "inputs larger than L2",
            "rebuild_every": REBUILD_EVERY}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag.is_set():
                    break
                self.samples.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag.set()
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx = max(mx, float(s[1]))
                for n, v in zip(names, s[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes(pn, pairs, s_vec=32, ptr_bytes=4):
    """list read once, q read once, p read+written once, nop+pointer read once."""
    return 4 * pairs + pn * (s_vec + 2 * s_vec + 4 + ptr_bytes)


def kernel_source_sha():
    """sha256 over the sources of the dominant kernel: ties an ncu capture to the code it was taken from"""
    import hashlib
    h = hashlib.sha256()
    for f in ("lj_force_celltile.cu", "lj_celltile.cuh", "lj_tile.cuh", "lj_common.cuh"):
        with open(os.path.join(ROOT, "lj_gpu_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def roofline_traffic(prec):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed
    `ncu --set full` capture -- only if that capture was taken from the kernel sources of this tree
    (profiles/roofline_traffic.json records their hash and the commit); a stale capture is dropped."""
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        tj = json.load(open(tpath))
        entry = tj["mixed_kernel"] if prec == "mixed" else tj
        if tj.get("kernel_source_sha") != kernel_source_sha():
            return None, "dropped: profiles/roofline_traffic.json was captured at %s from other kernel sources (%s != %s)" % (
                tj.get("git_head", "?"), tj.get("kernel_source_sha", "none"), kernel_source_sha())
        return entry.get("dram_bytes_per_launch"), "%s, captured at commit %s (%s)" % (
            entry.get("kernel"), tj.get("git_head"), tj.get("source"))
    except Exception as e:  # noqa: BLE001
        return None, "no capture: %s" % e


# ----------------------------------------------------------------------------- reference arm
def cpu_reference_sample(steps_force=20, L=50.0, density=1.0):
    """The real reference on the host: one makepair()+sortpair() and `steps_force` x
    force_sorted() at the workload's density in the reference's own box, L=50 (N=119,164 at rho=1.0,
    97,556 at rho=0.8; its static arrays end at N=400,000): one rebuild period of the bench cadence."""
    from oracle import ljoracle as lo
    if lo.have_ref(density):
        ref = lo.Ref(density, L)
        t0 = time.perf_counter()
        nop, ptr, lst = ref.makepair()
        t_list = time.perf_counter() - t0
        ref.zero_p()
        t0 = time.perf_counter()
        ref.force("sorted", steps_force)
        t_force = time.perf_counter() - t0
        return dict(kind="reference", pn=ref.pn, pairs_half=len(lst), t_list=t_list, t_force=t_force,
                    steps=steps_force, cores=1,
                    what="cpu_ref/force_soa.cpp makepair()+sortpair() once + %d x force_sorted(), "
                         "rho=%.1f L=50 N=%d, 1 thread (the reference is serial)" % (steps_force, density, ref.pn))
    # prebuilt reference missing: fall back to the C restatement of the same loops
    import numpy as np
    o = lo.Oracle()
    q = o.init_fcc(density, L)
    o.set_num_threads(1)
    t0 = time.perf_counter()
    nop, ptr, lst = o.makepair(q, full=False, brute=True)
    t_list = time.perf_counter() - t0
    p = np.zeros_like(q)
    t0 = time.perf_counter()
    o.force_sorted(q, p, nop, ptr, lst, steps=steps_force)
    t_force = time.perf_counter() - t0
    return dict(kind="port", pn=len(q), pairs_half=len(lst), t_list=t_list, t_force=t_force,
                steps=steps_force, cores=1,
                what="oracle/lj_oracle.c brute-force makepair once + %d x force_sorted, rho=%.1f L=50 "
                     "N=%d, 1 thread" % (steps_force, density, len(q)))


def cpu_restatement_config_c(q, nop, ptr, lst):
    """SURVEY 8(d) / BASELINE.md 4.4: the CPU RESTATEMENT (oracle/lj_oracle.c -- NOT the reference, which
    cannot run at this size: static N = 400000, O(N^2) makepair) on the bench's own config: one full-list
    gather step on 1 core and on all host cores (OpenMP), and its O(N) cell-list build on 1 core."""
    import numpy as np
    from oracle import ljoracle as lo
    o = lo.Oracle()
    nthreads = os.cpu_count() or 1
    out = {"kind": "port", "what": "oracle/lj_oracle.c full-list gather (cuda/kernel.cuh:36-65 arithmetic) + O(N) cell-list "
                                   "makepair on the bench config N=%d, labelled NOT the reference" % len(q), "nproc": nthreads}
    P = len(lst)
    p = np.zeros_like(q)
    for cores in (1, nthreads):
        o.set_num_threads(cores)
        o.force_gather(q, p, nop, ptr, lst, steps=1)          # warm
        t0 = time.perf_counter()
        o.force_gather(q, p, nop, ptr, lst, steps=2)
        t = (time.perf_counter() - t0) / 2
        out["force_step_s_%s" % ("1core" if cores == 1 else "allcores")] = t
        out["force_pairs_per_s_%s" % ("1core" if cores == 1 else "allcores")] = P / t
    o.set_num_threads(1)
    t0 = time.perf_counter()
    o.makepair(q, full=True, cap=P + 1024)
    out["list_build_s_1core"] = time.perf_counter() - t0
    amort = out["force_step_s_allcores"] + out["list_build_s_1core"] / REBUILD_EVERY
    out["amortised_pairs_per_s_allcores_force_1core_build"] = P / amort
    o.set_num_threads(nthreads)
    return out


def repo_on_reference_configs(ctx, torch, np, init_fcc, stream):
    """This repository's kernels on the reference's own two configurations (A: rho=0.5, B: rho=1.0,
    L=50), on lists whose rows were shuffled like the reference's random_shfl(): seconds per LOOP=100
    steps, kernel only -- the number cuda/force_cuda.cu:341 prints."""
    rows = []
    for rho in (0.5, 1.0):
        q = init_fcc(rho, 50.0)
        q4 = np.zeros((len(q), 4)); q4[:, :3] = q
        qd = torch.from_numpy(q4).cuda(); pd = torch.zeros_like(qd)
        best = None
        for name, build, fkw in (("per-row gather, 8 lanes per row (shuffled rows)", dict(), dict(variant="subwarp", group=8)),
                                 ("warp per row = warp_unroll mapping (shuffled rows)", dict(), dict(variant="warp", group=32)),
                                 ("cell-tile mirror (library-built list)", dict(tiles=True), dict(variant="celltile"))):
            pl = ctx.makepair(qd, **build)
            if not build:
                ctx.random_shfl(pl, seed=10)
            for graph in (False, True):
                ctx.force_loop(qd, pd, pl, loop=100, use_graph=graph, **fkw)
                torch.cuda.synchronize()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record(stream)
                for _ in range(5):
                    ctx.force_loop(qd, pd, pl, loop=100, use_graph=graph, **fkw)
                a1.record(stream)
                torch.cuda.synchronize()
                sec = a0.elapsed_time(a1) / 5 * 1e-3
                if best is None or sec < best["seconds_per_100_steps"]:
                    best = {"kernel": name, "cuda_graph": graph, "seconds_per_100_steps": sec,
                            "pairs_per_s": pl.number_of_pairs * 100 / sec}
            del pl
        rows.append({"density": rho, "N": len(q), **best})
        del qd, pd
    return rows


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # K steps of the bench cadence on the sample: K force steps + ceil(K/20) list builds.  The
    # reference's O(N^2) makepair takes ~15 s on the sample, so it is timed ONCE and charged per
    # scheduled rebuild; the force loop is timed on min(K, 40) steps and scaled.
    k_force = max(1, min(args.steps, 40))
    density = args.density if args.gpus <= 1 else CONFIG5_DENSITY
    r = cpu_reference_sample(k_force, density=density)
    per_step = r["t_force"] / r["steps"]
    builds = (args.steps + REBUILD_EVERY - 1) // REBUILD_EVERY
    total = per_step * args.steps + builds * r["t_list"]
    pairs_directed = 2 * r["pairs_half"]
    value = pairs_directed * args.steps / total
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak" if args.gpus <= 1 else "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "gpu_launches": 0,
        # the other arm's workload; what was timed here is a bounded SAMPLE of it (same lattice, density, cutoff,
        # search length and cadence in the reference's own L=50 box), stated in `sample` / cpu_baseline.sample
        "config": bench_config(args, args.gpus),
        "sample": "L=50 box of the same lattice: N=%d, %d directed pairs per step (full-list equivalent of the "
                  "reference's half list); pairs/s of the sample stands for the workload (the reference's O(N^2) "
                  "makepair would cost more per particle at the full size, so this favours the reference)" % (
                      r["pn"], pairs_directed),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                         "sample": r["what"] + "; force loop timed on %d steps (%.3f s), makepair timed "
                                   "once (%.2f s) and charged %d times" % (r["steps"], r["t_force"], r["t_list"], builds),
                         "force_only_pairs_per_s": pairs_directed / per_step,
                         "ms_per_force_step": 1e3 * per_step, "s_per_list_build": r["t_list"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------- CUDA arm
def run_cuda(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from lj_gpu_b200 import LJContext, init_fcc

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        from lj_gpu_b200 import decomp
        return decomp.bench_decomposed(args, METRIC, UNIT, REBUILD_EVERY, ClockSampler, measured_peak_gbs,
                                       algorithmic_bytes, cpu_reference_sample, bench_config)
    torch.cuda.set_device(local)
    ctx = LJContext(local)
    stream = torch.cuda.current_stream()
    K, W = args.steps, max(args.warmup, 3)

    q = init_fcc(args.density, args.L)
    pn = q.shape[0]
    q4 = np.zeros((pn, 4)); q4[:, :3] = q
    qd = torch.from_numpy(q4).cuda()
    pd = torch.zeros_like(qd)
    use_cl = args.variant == "cluster" and args.prec == "fp64"
    # "auto" on a list the library builds itself: the build also emits the cell-tile mirror and the
    # FP64 step runs on it (k_tile_permute + lj_celltile_force, two launches per step)
    use_tiles = args.variant in ("auto", "celltile")
    if use_tiles and args.prec == "mixed":
        use_tiles = "wide"   # LJ_LIST_TILES_WIDE: the tile size the mixed kernel prefers
    pl = ctx.makepair(qd, pointer64=False, clusters=use_cl, tiles=use_tiles)
    P = pl.number_of_pairs
    list_host = None
    if not args.no_cpu:   # the CPU restatement leg runs on the SAME list (row order is unspecified by contract)
        list_host = (pl.number_of_partners.cpu().numpy(), pl.pointer.cpu().numpy().astype(np.int64),
                     pl.sorted_list[:P].cpu().numpy())
    fkw = dict(variant=args.variant, group=args.group, precision=args.prec,
               threads_per_block=args.threads_per_block)

    def step_block(n0, n):
        """steps n0 .. n0+n-1 of the cadence: rebuild at multiples of REBUILD_EVERY"""
        s = n0
        while s < n0 + n:
            if s % REBUILD_EVERY == 0:
                ctx.rebuild(qd, pl, clusters=use_cl, tiles=use_tiles)
            m = min(REBUILD_EVERY - s % REBUILD_EVERY, n0 + n - s)
            ctx.force_loop(qd, pd, pl, loop=m, **fkw)
            s += m

    # ---- warm-up, then the timed region: EXACTLY K steps between two events
    step_block(0, W)
    torch.cuda.synchronize()
    sampler = ClockSampler(local); sampler.start()
    time.sleep(0.3)
    l0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    step_block(0, K)
    e1.record(stream)
    torch.cuda.synchronize()
    launches = ctx.launches - l0
    ms = e0.elapsed_time(e1)

    # ---- roofline leg: the force kernel alone, live CUDA events, list (564 MB) >> L2
    nf = max(20, min(K, 200))
    ctx.force_loop(qd, pd, pl, loop=3, **fkw)
    torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(stream)
    ctx.force_loop(qd, pd, pl, loop=nf, **fkw)
    f1.record(stream)
    torch.cuda.synchronize()
    ms_force = f0.elapsed_time(f1) / nf                    # the whole force step (position permute + force kernel)
    # the dominant kernel ALONE: the same loop again with the library's event pairs around each of its launches
    ctx.kernel_timing(True)
    ctx.force_loop(qd, pd, pl, loop=nf, **fkw)
    torch.cuda.synchronize()
    kt_ms, kt_n = ctx.kernel_timing_read()
    ctx.kernel_timing(False)
    ms_kernel = kt_ms / kt_n if kt_n == nf else None       # lj_celltile_force alone (None: AUTO took a per-row kernel)
    # ---- side measurement: the same force step in the mixed-precision mode north_star allows
    #      (FP32 pair arithmetic on fixed-point positions, FP64 momenta; parity bound 1e-5), on the
    #      same list, with its deviation from the FP64 kernel after 20 steps
    mixed = None
    if args.prec == "fp64":
        try:
            mkw = dict(fkw); mkw["precision"] = "mixed"
            pm, pf = torch.zeros_like(qd), torch.zeros_like(qd)
            ctx.force_loop(qd, pf, pl, loop=20, **fkw)
            if use_tiles:   # the same list with the mirror's tiles sized for the mixed kernel
                ctx.rebuild(qd, pl, clusters=use_cl, tiles="wide")
            ctx.force_loop(qd, pm, pl, loop=20, **mkw)
            torch.cuda.synchronize()
            dev = ((pm - pf)[:, :3].abs().max() / pf[:, :3].abs().max()).item()
            m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            m0.record(stream)
            ctx.force_loop(qd, pm, pl, loop=nf, **mkw)
            m1.record(stream)
            torch.cuda.synchronize()
            ms_mixed = m0.elapsed_time(m1) / nf
            # (one untimed rebuild first: the first rebuild after the first mixed-precision step grows the
            # stream-ordered pool -- a one-off of 10-100 ms that is not the cost of a rebuild)
            ctx.rebuild(qd, pl, clusters=use_cl, tiles="wide" if use_tiles else False)
            torch.cuda.synchronize()
            mb0, mb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            mb0.record(stream)
            for _ in range(5):
                ctx.rebuild(qd, pl, clusters=use_cl, tiles="wide" if use_tiles else False)
            mb1.record(stream)
            torch.cuda.synchronize()
            ms_build_mixed = mb0.elapsed_time(mb1) / 5
            if use_tiles:
                ctx.rebuild(qd, pl, clusters=use_cl, tiles=use_tiles)   # back to the FP64 mirror
            mixed = {"ms_per_launch": ms_mixed, "pairs_per_s_force_only": P / (ms_mixed * 1e-3),
                     "list_build_ms": ms_build_mixed,
                     "max_rel_deviation_from_fp64_after_20_steps": dev, "bound_stated": 1e-5,
                     "kernel": "k_tile_permute_fx + lj_celltile_force<mixed>" if use_tiles else "lj_gather_mixed"}
            del pm, pf
        except Exception as e:  # noqa: BLE001 -- the side measurement must not take the headline down
            mixed = {"error": str(e)}
    # ---- one list rebuild alone
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0.record(stream)
    for _ in range(5):
        ctx.rebuild(qd, pl, clusters=use_cl, tiles=use_tiles)
    b1.record(stream)
    torch.cuda.synchronize()
    ms_build = b0.elapsed_time(b1) / 5
    clocks = sampler.finish()

    # ---- e2e: the plugin call on pinned HOST buffers (the reference's measure())
    qh = torch.from_numpy(q4).pin_memory()
    ph = torch.zeros_like(qh).pin_memory()
    ctx.measure(qh.numpy(), ph.numpy(), layout="aos4", loop=min(K, REBUILD_EVERY), rebuild_every=REBUILD_EVERY, **fkw)
    ph.zero_()
    m = ctx.measure(qh.numpy(), ph.numpy(), layout="aos4", loop=K, rebuild_every=REBUILD_EVERY, **fkw)
    e2e_value = P * K / m.seconds_total
    checksum = float(ph.numpy()[:, :3].sum())

    # ---- the reference's own published workload (rho=0.5, L=50, N=62,500, LOOP=100, FP64, full list
    #      4,536,276 pairs): "without Host<->Device" seconds per 100 steps, as cuda/force_cuda.cu:341
    #      prints it; published best = 0.017518 s on P100 (profile/p100/cuda_p100.log:56)
    ref_cfg = None
    try:
        qa = init_fcc(0.5, 50.0)
        qa4 = np.zeros((len(qa), 4)); qa4[:, :3] = qa
        qad = torch.from_numpy(qa4).cuda(); pad_ = torch.zeros_like(qad)
        pla = ctx.makepair(qad)
        for graph in (False, True):
            ctx.force_loop(qad, pad_, pla, loop=100, use_graph=graph, **fkw)   # warm (and capture)
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            for _ in range(5):
                ctx.force_loop(qad, pad_, pla, loop=100, use_graph=graph, **fkw)
            a1.record(stream)
            torch.cuda.synchronize()
            sec100 = a0.elapsed_time(a1) / 5 * 1e-3
            if ref_cfg is None or sec100 < ref_cfg["seconds_per_100_steps"]:
                ref_cfg = {"workload": "reference config: rho=0.5 L=50 N=%d, %d directed pairs, LOOP=100, FP64, "
                                       "kernel only (no H<->D)" % (len(qa), pla.number_of_pairs),
                           "seconds_per_100_steps": sec100, "ms_per_step": sec100 * 10,
                           "pairs_per_s": pla.number_of_pairs * 100 / sec100, "cuda_graph": graph,
                           "published_best_p100_seconds_per_100_steps": 0.017518,
                           "speedup_vs_published_p100": 0.017518 / sec100}
        del qad, pad_, pla
    except Exception as e:  # noqa: BLE001 -- the side measurement must not take the headline down
        ref_cfg = {"error": str(e)}

    # ---- the reference's OWN CUDA kernels recompiled for sm_100 (baseline/_ref), on this GPU, next to this
    #      repository's kernels on the same two configurations
    ref_gpu = None
    if not args.no_ref_gpu:
        try:
            sys.path.insert(0, os.path.join(ROOT, "baseline"))
            import run_reference_gpu as rrg
            ref_gpu = rrg.reference_gpu((0.5, 1.0), timeout=300.0)
            mine = repo_on_reference_configs(ctx, torch, np, init_fcc, stream)
            for row in ref_gpu["rows"]:
                for m_ in mine:
                    if m_["density"] == row.get("density") and "best_seconds_per_100_steps" in row:
                        row["this_repo"] = m_
                        row["this_repo_faster_than_best_reference_kernel"] = \
                            m_["seconds_per_100_steps"] < row["best_seconds_per_100_steps"]
                        row["speedup_vs_best_reference_kernel"] = row["best_seconds_per_100_steps"] / m_["seconds_per_100_steps"]
        except Exception as e:  # noqa: BLE001 -- the side measurement must not take the headline down
            ref_gpu = {"error": str(e)}

    # ---- BASELINE config 5 on ONE GPU (N = 131,072,000): the denominator of the strong-scaling runs
    cfg5 = None
    if not args.no_config5:
        try:
            del qd, pd, pl
            torch.cuda.empty_cache()
            from lj_gpu_b200.decomp import lattice_spacing
            t0 = time.perf_counter()
            q5 = init_fcc(0.8, (320 + 0.05) * lattice_spacing(0.8))
            n5 = len(q5)
            q54 = np.zeros((n5, 4)); q54[:, :3] = q5
            del q5
            q5d = torch.from_numpy(q54).cuda(); del q54
            p5d = torch.zeros_like(q5d)
            pl5 = ctx.makepair(q5d, pointer64=True, tiles=True)
            setup = time.perf_counter() - t0
            ctx.force_loop(q5d, p5d, pl5, loop=2, **fkw)
            torch.cuda.synchronize()
            c0, c1, c2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            c0.record(stream)
            ctx.force_loop(q5d, p5d, pl5, loop=10, **fkw)
            c1.record(stream)
            ctx.rebuild(q5d, pl5)
            c2.record(stream)
            torch.cuda.synchronize()
            f5, b5 = c0.elapsed_time(c1) / 10, c1.elapsed_time(c2)
            cfg5 = {"workload": "BASELINE config 5 on one GPU: FCC rho=0.8, 320 cells/side, N=%d, %d directed pairs, int64 pointer[]" % (n5, pl5.number_of_pairs),
                    "ms_per_force_step": f5, "list_build_ms": b5, "amortised_step_ms": f5 + b5 / REBUILD_EVERY,
                    "pairs_per_s_amortised": pl5.number_of_pairs / ((f5 + b5 / REBUILD_EVERY) * 1e-3),
                    "roofline_frac_force": algorithmic_bytes(n5, pl5.number_of_pairs, 32, 8) / (f5 * 1e-3) / 1e9 / measured_peak_gbs()[0],
                    "setup_seconds_generator_and_first_build": setup,
                    "note": "strong-scaling efficiency of `bench.py --gpus N` = this amortised_step_ms / (N * its ms_per_step)"}
            del q5d, p5d, pl5
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            cfg5 = {"error": str(e)}

    peak, peak_src = measured_peak_gbs()
    bytes_force = algorithmic_bytes(pn, P)
    # roofline of the DOMINANT KERNEL: its own average launch duration, live CUDA events on its stream
    ms_roof = ms_kernel if ms_kernel else ms_force
    achieved = bytes_force / (ms_roof * 1e-3) / 1e9
    traffic, traffic_src = roofline_traffic(args.prec)

    cfg = bench_config(args, 1)
    if fcc_cells(args.density, args.L)[1] != pn:   # cannot happen: lj_init_fcc is the same arithmetic
        cfg["workload"] += " [generator returned N=%d]" % pn
    if 4 * P <= 126e6:   # a small --L: the statement in `config` would be false
        cfg["l2"] = "NOT larger than L2 (list %.0f MB) and no flush: not a valid bench size" % (4 * P / 1e6)
    out = {
        "metric": METRIC, "value": P * K / (ms * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": K,
        "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64" if args.prec == "fp64" else "f32-mixed",
        "data": "synthetic",
        "config": cfg,
        "pairs_per_step": int(P),
        "l2_detail": "neighbour list %.0f MB per step vs 126 MB L2" % (4 * P / 1e6),
        "kernel_selection": {"layout": "aos_double4", "variant": args.variant, "group": args.group,
                             "list_build": "on the GPU (lj_build_list)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": m.h2d_bytes / K,
                "d2h_bytes_per_step": m.d2h_bytes / K, "seconds_total": m.seconds_total,
                "seconds_kernel": m.seconds_kernel, "list_builds": m.list_builds,
                "call": "lj_measure(): pinned host q,p -> H2D -> GPU list build -> K steps (rebuild "
                        "every 20) -> D2H p", "p_checksum": checksum},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peak_src,
                     "kernel": ("lj_celltile_force (%s), timed alone by the library's own event pairs "
                                "(lj_kernel_timing)" % args.prec if ms_kernel
                                else "force step (k_tile_permute + lj_celltile_force, %s)" % args.prec if use_tiles
                                else "force step (lj_gather_*)"),
                     "algorithmic_bytes_per_launch": bytes_force,
                     "ms_per_launch": ms_roof,
                     "force_step_ms": ms_force, "frac_of_whole_force_step": bytes_force / (ms_force * 1e-3) / 1e9 / peak,
                     "force_step_is": "k_tile_permute (positions into cell order) + lj_celltile_force" if use_tiles else "one kernel",
                     "pairs_per_s_force_only": P / (ms_force * 1e-3),
                     "list_build_ms": ms_build,
                     "amortised_step_ms": ms_force + ms_build / REBUILD_EVERY},
    }
    if mixed is not None and "ms_per_launch" in mixed:
        # same algorithmic bytes: the caller's arrays are the same FP64 q/p and int32 list
        mixed["roofline_frac"] = bytes_force / (mixed["ms_per_launch"] * 1e-3) / 1e9 / peak
        mixed["amortised_step_ms"] = mixed["ms_per_launch"] + mixed["list_build_ms"] / REBUILD_EVERY
        mixed["pairs_per_s_amortised"] = P / (mixed["amortised_step_ms"] * 1e-3)
    out["mixed_precision"] = mixed
    out["reference_config"] = ref_cfg
    out["reference_gpu"] = ref_gpu
    out["config5_single_gpu"] = cfg5
    if not args.no_cpu:
        r = cpu_reference_sample(20, density=args.density if args.density in (0.5, 0.8, 1.0) else 1.0)
        t = r["t_force"] + r["t_list"]
        out["cpu_baseline"] = {
            "value": 2 * r["pairs_half"] * r["steps"] / t, "unit": UNIT, "cores": r["cores"],
            "kind": r["kind"], "sample": r["what"],
            "force_only_pairs_per_s": 2 * r["pairs_half"] * r["steps"] / r["t_force"],
            "ms_per_force_step": 1e3 * r["t_force"] / r["steps"], "s_per_list_build": r["t_list"]}
        try:   # the restatement on the bench's own configuration (1 core and all cores)
            out["cpu_baseline"]["restatement_config_c"] = cpu_restatement_config_c(q, *list_host)
        except Exception as e:  # noqa: BLE001
            out["cpu_baseline"]["restatement_config_c"] = {"error": str(e)}
    print(json.dumps(out), flush=True)
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--density", type=float, default=1.0)
    ap.add_argument("--L", type=float, default=100.1, help="box edge; 100.1 -> 63 cells/side -> N=1,000,188")
    ap.add_argument("--variant", default="auto")
    ap.add_argument("--group", type=int, default=0)
    ap.add_argument("--prec", default="fp64", choices=["fp64", "mixed"])
    ap.add_argument("--threads-per-block", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the reference_gpu block (baseline/_ref binaries)")
    ap.add_argument("--no-config5", action="store_true", help="skip the config5_single_gpu side block (N=131M on one GPU)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
