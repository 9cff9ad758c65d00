"""z-slab spatial decomposition of the LJ force + neighbour-list path over the GPUs of one box.

No counterpart in the reference (single GPU, no NCCL/MPI: SURVEY 0.9).  The path shards with ONE
exchange step because the gather kernels only write p[i] of owned particles: per step each rank
needs the current POSITIONS of the ghost particles within the search length of its slab faces,
nothing flows back (no reverse force communication).

Layout.  init() emits particles with iz outermost (cuda/force_cuda.cu:68-77), so a slab of
lattice layers is a CONTIGUOUS range of global indices and so are the ghost layers a neighbour
needs: no packing kernels.  A rank's local array is

    q_local = [ owned (n_own) | ghosts from the rank below (n_lo) | ghosts from above (n_hi) ]

with local int32 indices; the list is built for rows [0, n_own) with all n_local particles as
candidates; boundary rows (within `halo` layers of a face) wait for the ghosts, interior rows do
not, which is what lets the halo transfer overlap the interior force kernel.

Transports: "nccl"/"gloo" = torch.distributed batch_isend_irecv (grouped ncclSend/ncclRecv over
NVLink); "p2p" = lj_halo_pull, a copy kernel doing 16-byte loads straight from the neighbour's
memory mapped with CUDA IPC (NVLink P2P), launched on a side stream.
"""
from __future__ import annotations

import json
import os
import time
from dataclasses import dataclass

import numpy as np


@dataclass
class Slab:
    """Index bookkeeping of one rank (pure host logic, no device, no communication)."""
    rank: int
    world: int
    cells: int            # lattice cells per side (n)
    z0: int               # first owned lattice layer
    z1: int               # one past the last owned layer
    halo: int             # ghost layers on each interior face
    layer: int            # particles per lattice layer = 4 n^2

    @property
    def lo(self) -> int:          # first owned global index
        return self.z0 * self.layer

    @property
    def hi(self) -> int:
        return self.z1 * self.layer

    @property
    def n_own(self) -> int:
        return self.hi - self.lo

    @property
    def n_lo(self) -> int:        # ghosts received from rank-1 (its top layers)
        return min(self.halo, self.z0) * self.layer if self.rank > 0 else 0

    @property
    def n_hi(self) -> int:        # ghosts received from rank+1 (its bottom layers)
        return min(self.halo, self.cells - self.z1) * self.layer if self.rank < self.world - 1 else 0

    @property
    def n_local(self) -> int:
        return self.n_own + self.n_lo + self.n_hi

    def global_ranges(self):
        """[(global_begin, global_end)] of the three local segments, in local order."""
        return [(self.lo, self.hi), (self.lo - self.n_lo, self.lo), (self.hi, self.hi + self.n_hi)]

    def interior_rows(self):
        """Local rows whose whole neighbourhood is owned: no ghost needed."""
        b = min(self.halo * self.layer, self.n_own) if self.n_lo else 0
        e = max(self.n_own - (self.halo * self.layer if self.n_hi else 0), b)
        return b, e

    def boundary_rows(self):
        b, e = self.interior_rows()
        return [(0, b), (e, self.n_own)]


def lattice_spacing(density: float) -> float:
    return 1.0 / (density * 0.25) ** (1.0 / 3.0)


def make_slab(rank: int, world: int, density: float, L: float, search_len: float = 3.3,
              jitter: float = 0.1) -> Slab:
    """Split the n lattice layers of init() as evenly as possible into `world` z-slabs.
    halo = number of lattice layers that can hold a particle within search_len of a slab face:
    layer iz spans z in [iz*s, iz*s + s/2 + jitter)."""
    s = lattice_spacing(density)
    n = int(L / s)
    if world > n:
        raise ValueError("more ranks (%d) than lattice layers (%d)" % (world, n))
    bounds = [(n * r) // world for r in range(world + 1)]
    halo = int(np.ceil((search_len + 0.5 * s + jitter) / s))
    thinnest = min(bounds[r + 1] - bounds[r] for r in range(world))
    if world > 1 and thinnest < halo:
        # the ghosts of a slab thinner than the halo live on rank +-2: the one-hop exchange below
        # would send rows it does not own
        raise ValueError("slab of %d lattice layers is thinner than the halo (%d layers): use fewer ranks "
                         "(<= %d) for this system" % (thinnest, halo, max(1, n // halo)))
    return Slab(rank, world, n, bounds[rank], bounds[rank + 1], halo, 4 * n * n)


def local_positions(slab: Slab, q_global_xyz: np.ndarray) -> np.ndarray:
    """Assemble q_local (float64 [n_local,3]) from the global lattice (setup only)."""
    return np.concatenate([q_global_xyz[b:e] for b, e in slab.global_ranges()], axis=0)


def halo_plan(slab: Slab, neighbours: dict):
    """Send/recv description of one halo exchange in LOCAL index ranges:
    [(peer_rank, 'send'|'recv', local_begin, local_end)].  `neighbours` maps rank -> Slab."""
    ops = []
    if slab.rank > 0:
        below = neighbours[slab.rank - 1]
        ops.append((below.rank, "send", 0, below.n_hi))                     # my bottom layers
        ops.append((below.rank, "recv", slab.n_own, slab.n_own + slab.n_lo))
    if slab.rank < slab.world - 1:
        above = neighbours[slab.rank + 1]
        ops.append((above.rank, "send", slab.n_own - above.n_lo, slab.n_own))  # my top layers
        ops.append((above.rank, "recv", slab.n_own + slab.n_lo, slab.n_local))
    return ops


def exchange_halo(q_local, plan, dist, async_op: bool = False):
    """One halo exchange with torch.distributed (gloo on CPU tensors, NCCL on CUDA tensors).
    q_local: tensor [n_local, w]; rows are contiguous so every op is one contiguous view."""
    ops = []
    for peer, kind, b, e in plan:
        if e <= b:
            continue
        view = q_local[b:e]
        ops.append(dist.P2POp(dist.isend if kind == "send" else dist.irecv, view, peer))
    if not ops:
        return []
    reqs = dist.batch_isend_irecv(ops)
    if not async_op:
        for r in reqs:
            r.wait()
    return reqs


# ------------------------------------------------------------------------------------------
# GPU execution (one process per GPU, torchrun)
# ------------------------------------------------------------------------------------------
class DecomposedSystem:
    """One rank (= one process, one GPU) of the z-slab decomposition.

    Per step:   publish()  "my q of step k is final"           (compute stream, lj_flag_set)
                halo()     pull the neighbours' boundary layers (comm stream; each pull WAITS on the
                           device for the owner's flag of step k and then tells the owner it is done)
                force      interior tiles / rows while the halo flies, boundary tiles / rows after
                           its event (lj_force_step_part on the cell-tile mirror, row ranges on the
                           per-row kernels)
    and before q is overwritten (drift):  wait_pulled()  both neighbours have read step k.
    """

    READY, PULLED_BY_BELOW, PULLED_BY_ABOVE = 0, 1, 2   # int32 slots of a rank's flag block

    def __init__(self, density: float, L: float, halo_mode: str = "nccl", search_len: float = 3.3,
                 tiles: bool = False):
        import torch
        import torch.distributed as dist

        from .api import LJContext, init_fcc
        self.torch, self.dist = torch, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.local = int(os.environ.get("LOCAL_RANK", self.rank))
        torch.cuda.set_device(self.local)
        self.ctx = LJContext(self.local)
        self.halo_mode = halo_mode
        self.slabs = {r: make_slab(r, self.world, density, L, search_len) for r in range(self.world)}
        self.slab = self.slabs[self.rank]
        self.plan = halo_plan(self.slab, self.slabs)
        # every rank runs the (sequential, single-stream) generator and keeps its own segments
        q_all = init_fcc(density, L)
        self.pn_global = q_all.shape[0]
        ql = local_positions(self.slab, q_all)
        del q_all
        q4 = np.zeros((ql.shape[0], 4)); q4[:, :3] = ql
        del ql
        if halo_mode == "p2p":  # the neighbours map this array and the flag block through CUDA IPC
            self.q = self.ctx.ipc_tensor(q4.shape, torch.float64)
            self.q.copy_(torch.from_numpy(q4))
            self.flags = self.ctx.ipc_tensor((4,), torch.int32)
            self.flags.zero_()
        else:
            self.q = torch.from_numpy(q4).cuda()
            self.flags = None
        del q4
        self.p = torch.zeros_like(self.q)
        self.compute = torch.cuda.current_stream()
        self.comm = torch.cuda.Stream()
        self.stepno = 0
        # tiles: also build the cell-tile mirror for the owned rows: both schedules then run the
        # shared-memory kernel (the overlapped one as lj_force_step_part INTERIOR / BOUNDARY)
        self.tiles = tiles
        # int32 pointer[] holds 2^32 - 1 list entries: 64-bit offsets for slabs beyond ~14 M particles
        # (the reference's int32 offsets overflow at BASELINE config 4 already)
        self.pl = self.ctx.makepair(self.q, rows=(0, self.slab.n_own), search_len=search_len, tiles=tiles,
                                    pointer64=self.slab.n_own * 160 > 2 ** 31)
        self.search_len = search_len
        self.pairs_local = self.pl.number_of_pairs
        self.peer_ptr, self.peer_flags = {}, {}
        if halo_mode == "p2p":
            ok = 1
            try:
                self._open_peers()
            except Exception:  # noqa: BLE001 -- no peer access on this box: every rank falls back
                ok = 0
            flag = torch.tensor([ok], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                self.halo_mode = "nccl"

    # -- CUDA IPC: map the neighbours' q arrays and flag blocks ------------------------------
    def _open_peers(self):
        dist = self.dist
        mine = (self.ctx.ipc_export(self.q), self.ctx.ipc_export(self.flags))  # each starts its own cudaMalloc block
        handles = [None] * self.world
        dist.all_gather_object(handles, mine)
        for peer in (self.rank - 1, self.rank + 1):
            if 0 <= peer < self.world:
                self.peer_ptr[peer] = self.ctx.ipc_open(handles[peer][0])
                self.peer_flags[peer] = self.ctx.ipc_open(handles[peer][1])
        dist.barrier()

    def publish(self):
        """q is final for the step that starts now (compute stream order)."""
        self.stepno += 1
        if self.halo_mode == "p2p":
            self.ctx.flag_set(self.flags.data_ptr() + 4 * self.READY, self.stepno, stream=self.compute)
        # my own exchange of this step must not start before my previous step is through with the ghost
        # rows it is about to overwrite (and, NCCL, before my q is final for the sends)
        self.q_final = self.torch.cuda.Event()
        self.q_final.record(self.compute)

    def wait_pulled(self):
        """Hold the compute stream until both neighbours have read this step's q (call before q changes)."""
        if self.halo_mode != "p2p":
            return   # NCCL send/recv pairs are ordered by the collective itself
        if self.rank > 0:
            self.ctx.flag_wait(self.flags.data_ptr() + 4 * self.PULLED_BY_BELOW, self.stepno, stream=self.compute)
        if self.rank < self.world - 1:
            self.ctx.flag_wait(self.flags.data_ptr() + 4 * self.PULLED_BY_ABOVE, self.stepno, stream=self.compute)

    def halo(self, wait_event=None):
        """Start one ghost-position exchange on the comm stream; returns an event."""
        torch = self.torch
        with torch.cuda.stream(self.comm):
            if wait_event is not None:
                self.comm.wait_event(wait_event)
            if getattr(self, "q_final", None) is not None:
                self.comm.wait_event(self.q_final)
            if self.halo_mode == "p2p":
                row = self.q.shape[1] * 8
                segs = []
                for peer, kind, b, e in self.plan:
                    if kind != "recv" or e <= b:
                        continue
                    ps = self.slabs[peer]
                    # my ghosts from below are the peer's top owned rows, from above its bottom rows;
                    # seen from the peer I am its neighbour above / below
                    src_b = ps.n_own - (e - b) if peer < self.rank else 0
                    done = self.PULLED_BY_ABOVE if peer < self.rank else self.PULLED_BY_BELOW
                    segs.append((self.q.data_ptr() + b * row, self.peer_ptr[peer] + src_b * row, (e - b) * row,
                                 self.peer_flags[peer] + 4 * self.READY, self.stepno,
                                 self.peer_flags[peer] + 4 * done, self.stepno))
                if segs:
                    self.ctx.halo_pull_sync(segs, stream=self.comm)
            else:
                exchange_halo(self.q, self.plan, self.dist)
            ev = torch.cuda.Event()
            ev.record(self.comm)
        return ev

    def force(self, ev, overlap: bool = True, **fkw):
        """The force step of this rank around the halo event `ev`."""
        ctx, s = self.ctx, self.slab
        rows = (0, s.n_own)
        if overlap and self.tiles and fkw.get("variant", "auto") in ("auto", "celltile"):
            # cell-tile kernel in two parts: tiles that read no ghost run while the halo is in flight
            ctx.force_step(self.q, self.p, self.pl, rows=rows, part="interior", **fkw)
            self.compute.wait_event(ev)
            ctx.force_step(self.q, self.p, self.pl, rows=rows, part="boundary", **fkw)
            return
        ib, ie = s.interior_rows()
        if overlap and ie > ib:
            ctx.force_step(self.q, self.p, self.pl, rows=(ib, ie), **fkw)
            self.compute.wait_event(ev)
            for b, e in s.boundary_rows():
                if e > b:
                    ctx.force_step(self.q, self.p, self.pl, rows=(b, e), **fkw)
        else:
            self.compute.wait_event(ev)
            ctx.force_step(self.q, self.p, self.pl, rows=rows, **fkw)

    def step(self, overlap: bool = True, rebuild: bool = False, **fkw):
        """One force step: halo exchange overlapped with the work that needs no ghost.  With
        rebuild the list is rebuilt from this step's positions first (needs the ghosts: no overlap)."""
        self.publish()
        ev = self.halo()
        if rebuild:
            self.compute.wait_event(ev)
            self.rebuild()
            self.force(ev, overlap=False, **fkw)
        else:
            self.force(ev, overlap=overlap, **fkw)

    def drift(self, dt: float):
        """q += p dt for the OWNED particles, once both neighbours have read this step's q."""
        self.wait_pulled()
        self.ctx.drift(self.q, self.p, dt=dt, pn=self.slab.n_own, stream=self.compute)

    def rebuild(self):
        self.ctx.rebuild(self.q, self.pl, search_len=self.search_len, rows=(0, self.slab.n_own),
                         tiles=self.tiles)

    def run(self, steps: int, rebuild_every: int, first_step: int = 0, overlap: bool = True, **fkw):
        """Static positions (the reference benchmark): `steps` force steps, list rebuilt every
        `rebuild_every` steps (the rebuilt list is the same list: the cadence measures its cost)."""
        for k in range(first_step, first_step + steps):
            self.step(overlap=overlap, rebuild=bool(rebuild_every and k % rebuild_every == 0), **fkw)

    def run_md(self, steps: int, dt: float, rebuild_every: int, overlap: bool = True, **fkw):
        """Moving particles: kick (force step) + drift of the owned particles, fixed rebuild cadence."""
        for k in range(steps):
            self.step(overlap=overlap, rebuild=bool(k > 0 and rebuild_every and k % rebuild_every == 0), dt=dt, **fkw)
            self.drift(dt)

    def gather(self, what: str = "p") -> np.ndarray | None:
        """Owned momenta (or positions) of every rank concatenated on rank 0 (tests / checks only)."""
        dist = self.dist
        src = self.p if what == "p" else self.q
        mine = src[:self.slab.n_own, :3].contiguous().cpu()
        out = [None] * self.world if self.rank == 0 else None
        dist.gather_object(mine.numpy(), out, dst=0)
        return np.concatenate(out, axis=0) if self.rank == 0 else None

    def gather_p(self) -> np.ndarray | None:
        return self.gather("p")


def _timed(torch, dist, fn, reps):
    """mean ms per call over `reps` calls, device timed, MAX over ranks"""
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _measure_system(system, K, W, rebuild_every, fkw, ClockSampler=None):
    """Warm up, pick the schedule (both are timed, untimed region), then time EXACTLY K steps of the
    cadence between two events; max over ranks.  Returns a dict of raw numbers (every rank)."""
    torch, dist = system.torch, system.dist
    system.run(W, rebuild_every, **fkw)
    torch.cuda.synchronize(); dist.barrier()
    for ov in (False, True):   # first launches load kernels
        system.step(overlap=ov, **fkw)
        system.step(overlap=ov, **fkw)
    reps = 10
    ms_serial = _timed(torch, dist, lambda: system.step(overlap=False, **fkw), reps)
    ms_overlap = _timed(torch, dist, lambda: system.step(overlap=True, **fkw), reps)
    use_overlap = ms_overlap <= ms_serial
    sampler = ClockSampler(system.local) if (ClockSampler and system.rank == 0) else None
    if sampler:
        sampler.start(); time.sleep(0.3)
    l0 = system.ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); dist.barrier()
    e0.record()
    system.run(K, rebuild_every, overlap=use_overlap, **fkw)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)          # max over ranks, device timed
    dist.barrier()
    launches = system.ctx.launches - l0
    pairs = torch.tensor([system.pl.number_of_pairs], dtype=torch.int64, device="cuda")
    dist.all_reduce(pairs)
    # the pieces: the force step alone (no halo wait beyond the schedule's), one list rebuild, the halo alone
    ms_force = _timed(torch, dist, lambda: system.step(overlap=use_overlap, **fkw), reps)
    ms_build = _timed(torch, dist, system.rebuild, 3)
    ms_halo = _timed(torch, dist, lambda: torch.cuda.current_stream().wait_event(system.halo()), reps)
    clocks = sampler.finish() if sampler else None
    return dict(ms_total=float(ms.item()), pairs=int(pairs.item()), launches=int(launches), ms_serial=ms_serial,
                ms_overlap=ms_overlap, use_overlap=use_overlap, ms_force=ms_force, ms_build=ms_build, ms_halo=ms_halo,
                clocks=clocks)


def bench_decomposed(args, metric, unit, rebuild_every, ClockSampler, measured_peak_gbs, algorithmic_bytes,
                     cpu_reference_sample=None, bench_config=None):
    """bench.py body for N > 1 (torchrun).  Main line: BASELINE config 5 -- FCC rho = 0.8, 320 cells per
    side, N = 131,072,000 -- split into N z-slabs: STRONG scaling (the total work is fixed).  Side
    block: the weak-scaling run of round 1 (~1.0e6 particles per GPU at the bench density)."""
    import torch
    import torch.distributed as dist

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    halo_mode = os.environ.get("LJ_HALO", "p2p")
    tiles = args.variant in ("auto", "celltile")
    if tiles and args.prec == "mixed":
        tiles = "wide"
    fkw = dict(variant=args.variant, group=args.group, precision=args.prec,
               threads_per_block=args.threads_per_block)
    K, W = args.steps, max(args.warmup, 3)
    peak, peak_src = measured_peak_gbs()

    # ---------------- main: config 5, strong scaling
    density5 = 0.8
    cells5 = int(os.environ.get("LJ_BENCH_CELLS", "320"))     # 320 -> N = 131,072,000
    L5 = (cells5 + 0.05) * lattice_spacing(density5)
    t0 = time.perf_counter()
    system = DecomposedSystem(density5, L5, halo_mode=halo_mode, tiles=tiles)
    setup_s = time.perf_counter() - t0
    # e2e: the resident arrays come from pinned HOST buffers and the momenta go back to the host; both
    # copies are timed on the device and charged to the K steps (the reference's measure(): upload,
    # LOOP launches, download)
    qh = system.q.cpu().pin_memory()
    ph = torch.zeros_like(qh).pin_memory()
    ms_h2d = _timed(torch, dist, lambda: (system.q.copy_(qh, non_blocking=True), system.p.copy_(ph, non_blocking=True)), 1)
    r = _measure_system(system, K, W, rebuild_every, fkw, ClockSampler)
    ms_d2h = _timed(torch, dist, lambda: ph.copy_(system.p, non_blocking=True), 1)
    checksum = float(ph[:system.slab.n_own, :3].sum())
    n_local, n_own = system.slab.n_local, system.slab.n_own
    pairs_local = system.pl.number_of_pairs
    # per-rank roofline of the force step: this rank's algorithmic bytes / its step time
    bytes_rank = algorithmic_bytes(n_own, pairs_local) + (n_local - n_own) * 32   # + the ghosts it reads
    ghost_bytes = int((system.slab.n_lo + system.slab.n_hi) * 32)
    sched = "overlap" if r["use_overlap"] else "halo-then-force"
    halo_used = system.halo_mode
    slab_layers = system.slab.z1 - system.slab.z0
    pn5 = system.pn_global
    del system, qh, ph
    torch.cuda.empty_cache()

    # ---------------- side: weak scaling, ~1.0e6 particles per GPU (round 1's SCALE workload)
    weak = None
    if not os.environ.get("LJ_BENCH_NO_WEAK"):
        s = lattice_spacing(args.density)
        n = int(round((250047.0 * world) ** (1.0 / 3.0)))
        n = max(world, (n + world - 1) // world * world)
        try:
            wsys = DecomposedSystem(args.density, (n + 0.05) * s, halo_mode=halo_mode, tiles=tiles)
            wr = _measure_system(wsys, min(K, 100), 5, rebuild_every, fkw)
            weak = {"workload": "FCC rho=%.1f, %d cells/side, N=%d (~1.0e6 per GPU), %d directed pairs" % (
                        args.density, n, wsys.pn_global, wr["pairs"]),
                    "value": wr["pairs"] * min(K, 100) / (wr["ms_total"] * 1e-3), "unit": unit,
                    "ms_per_step": wr["ms_total"] / min(K, 100), "scaling": "weak",
                    "schedule": "overlap" if wr["use_overlap"] else "halo-then-force",
                    "ms_step_overlap": wr["ms_overlap"], "ms_step_serial": wr["ms_serial"],
                    "ms_halo_alone": wr["ms_halo"], "list_build_ms": wr["ms_build"]}
            del wsys
        except Exception as e:  # noqa: BLE001 -- the side block must not take the headline down
            weak = {"error": str(e)}
    if rank == 0:
        P, ms_total = r["pairs"], r["ms_total"]
        out = {
            "metric": metric, "value": P * K / (ms_total * 1e-3), "unit": unit, "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,
            "dtype": "f64" if args.prec == "fp64" else "f32-mixed", "data": "synthetic",
            # the workload as a function of the command line (bench.py: bench_config, shared with the reference
            # arm); what this run measured on it is stated under "decomposition"
            "config": bench_config(args, world) if bench_config else {
                "workload": "BASELINE config 5: synthetic FCC lattice N=%d (%d cells/side) rho=%.1f cutoff=3.0 "
                            "search=3.3" % (pn5, cells5, density5),
                "parallelism": "z-slab x%d" % world, "l2": "inputs larger than L2", "rebuild_every": rebuild_every},
            "pairs_per_step": int(P),
            "decomposition": {"particles": int(pn5), "slabs": world, "lattice_layers_per_slab": int(slab_layers),
                              "halo": halo_used, "list_build": "on the GPU (lj_build_list), per slab",
                              "setup_seconds_generator_and_first_build": setup_s},
            "gpu_launches": r["launches"], "clocks": r["clocks"],
            "halo": {"mode": halo_used, "schedule": sched, "ms_step_overlap": r["ms_overlap"],
                     "ms_step_serial": r["ms_serial"], "ms_halo_alone": r["ms_halo"],
                     "ghost_bytes_per_step_per_rank": ghost_bytes,
                     "ordering": "device-side step counters (lj_flag_set / lj_halo_pull_sync / lj_flag_wait)"
                                 if halo_used == "p2p" else "NCCL send/recv"},
            "roofline": {"bound": "hbm", "achieved": bytes_rank / (r["ms_force"] * 1e-3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": bytes_rank / (r["ms_force"] * 1e-3) / 1e9 / peak, "traffic": None,
                         "peak_source": peak_src, "per": "rank 0 (its slab: %d owned + %d ghost particles, %d pairs)" % (
                             n_own, n_local - n_own, pairs_local),
                         "kernel": "force step of one rank incl. its halo wait (k_tile_permute + lj_celltile_force x2 parts)",
                         "algorithmic_bytes_per_launch": bytes_rank, "ms_per_launch": r["ms_force"],
                         "list_build_ms": r["ms_build"],
                         "amortised_step_ms": r["ms_force"] + r["ms_build"] / rebuild_every},
            "e2e": {"value": P * K / ((ms_total + ms_h2d + ms_d2h) * 1e-3), "unit": unit,
                    "h2d_bytes_per_step": 2 * n_local * 32 / K, "d2h_bytes_per_step": n_local * 32 / K,
                    "ms_h2d_once": ms_h2d, "ms_d2h_once": ms_d2h, "p_checksum_rank0": checksum,
                    "call": "per rank: pinned host q,p -> H2D once, K decomposed steps (rebuild every %d), D2H p once; "
                            "bytes per rank, charged to the K steps" % rebuild_every},
            "weak_scaling_1M_per_gpu": weak,
        }
        if cpu_reference_sample is not None and not getattr(args, "no_cpu", False):
            try:
                c = cpu_reference_sample(20, density=density5)
                t = c["t_force"] + c["t_list"]
                out["cpu_baseline"] = {"value": 2 * c["pairs_half"] * c["steps"] / t, "unit": unit, "cores": c["cores"],
                                       "kind": c["kind"], "sample": c["what"],
                                       "force_only_pairs_per_s": 2 * c["pairs_half"] * c["steps"] / c["t_force"]}
            except Exception as e:  # noqa: BLE001
                out["cpu_baseline"] = {"error": str(e)}
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()
