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
        if halo_mode == "p2p":  # the neighbours map this array through CUDA IPC
            self.q = self.ctx.ipc_tensor(q4.shape, torch.float64)
            self.q.copy_(torch.from_numpy(q4))
        else:
            self.q = torch.from_numpy(q4).cuda()
        self.p = torch.zeros_like(self.q)
        self.compute = torch.cuda.current_stream()
        self.comm = torch.cuda.Stream()
        # tiles: also build the cell-tile mirror for the owned rows; the halo-then-force schedule
        # (one launch over rows [0, n_own)) then runs the shared-memory kernel, the overlapped
        # schedule (row sub-ranges) the per-row kernel
        self.tiles = tiles
        self.pl = self.ctx.makepair(self.q, rows=(0, self.slab.n_own), search_len=search_len, tiles=tiles)
        self.search_len = search_len
        self.pairs_local = self.pl.number_of_pairs
        self.peer_ptr = {}
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

    # -- CUDA IPC: map the neighbours' q arrays ------------------------------------------
    def _open_peers(self):
        dist, torch = self.dist, self.torch
        handle = self.ctx.ipc_export(self.q)  # q starts its own cudaMalloc block: offset 0
        handles = [None] * self.world
        dist.all_gather_object(handles, handle)
        for peer in (self.rank - 1, self.rank + 1):
            if 0 <= peer < self.world:
                self.peer_ptr[peer] = self.ctx.ipc_open(handles[peer])
        dist.barrier()

    def halo(self, wait_event=None):
        """Start one ghost-position exchange on the comm stream; returns an event."""
        torch = self.torch
        with torch.cuda.stream(self.comm):
            if wait_event is not None:
                self.comm.wait_event(wait_event)
            if self.halo_mode == "p2p":
                row = self.q.shape[1] * 8
                for peer, kind, b, e in self.plan:
                    if kind != "recv" or e <= b:
                        continue
                    ps = self.slabs[peer]
                    # my ghosts from below are the peer's top owned rows, from above its bottom rows
                    src_b = ps.n_own - (e - b) if peer < self.rank else 0
                    self.ctx.halo_pull(self.q.data_ptr() + b * row, self.peer_ptr[peer] + src_b * row,
                                       (e - b) * row, stream=self.comm)
            else:
                exchange_halo(self.q, self.plan, self.dist)
            ev = torch.cuda.Event()
            ev.record(self.comm)
        return ev

    def step(self, overlap: bool = True, **fkw):
        """One force step: halo exchange overlapped with the interior rows."""
        ctx, s = self.ctx, self.slab
        ev = self.halo()
        ib, ie = s.interior_rows()
        if overlap and ie > ib:
            ctx.force_step(self.q, self.p, self.pl, rows=(ib, ie), **fkw)
            self.compute.wait_event(ev)
            for b, e in s.boundary_rows():
                if e > b:
                    ctx.force_step(self.q, self.p, self.pl, rows=(b, e), **fkw)
        else:
            self.compute.wait_event(ev)
            ctx.force_step(self.q, self.p, self.pl, rows=(0, s.n_own), **fkw)

    def rebuild(self):
        self.ctx.rebuild(self.q, self.pl, search_len=self.search_len, rows=(0, self.slab.n_own),
                         tiles=self.tiles)

    def run(self, steps: int, rebuild_every: int, first_step: int = 0, overlap: bool = True, **fkw):
        for k in range(first_step, first_step + steps):
            if rebuild_every and k % rebuild_every == 0:
                self.rebuild()
            self.step(overlap=overlap, **fkw)

    def gather_p(self) -> np.ndarray | None:
        """Owned momenta of every rank concatenated on rank 0 (tests / checks only)."""
        dist, torch = self.dist, self.torch
        mine = self.p[:self.slab.n_own, :3].contiguous().cpu()
        out = [None] * self.world if self.rank == 0 else None
        dist.gather_object(mine.numpy(), out, dst=0)
        return np.concatenate(out, axis=0) if self.rank == 0 else None


def bench_decomposed(args, metric, unit, rebuild_every, ClockSampler, measured_peak_gbs, algorithmic_bytes):
    """bench.py body for N > 1 (torchrun): weak scaling, ~1M particles per GPU."""
    import torch
    import torch.distributed as dist

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    density = args.density
    s = lattice_spacing(density)
    # cubic lattice with ~1.0e6 * world particles and a layer count divisible by world
    n = int(round((250047.0 * world) ** (1.0 / 3.0)))
    n = max(world, (n + world - 1) // world * world)
    if os.environ.get("LJ_BENCH_CELLS"):  # e.g. 320 at rho=0.8: BASELINE config 5 (N=131,072,000)
        n = int(os.environ["LJ_BENCH_CELLS"])
    L = (n + 0.05) * s
    halo_mode = os.environ.get("LJ_HALO", "p2p")
    tiles = args.variant in ("auto", "celltile") and args.prec == "fp64"
    system = DecomposedSystem(density, L, halo_mode=halo_mode, tiles=tiles)
    halo_mode = system.halo_mode
    fkw = dict(variant=args.variant, group=args.group, precision=args.prec,
               threads_per_block=args.threads_per_block)
    K, W = args.steps, max(args.warmup, 3)
    system.run(W, rebuild_every, **fkw)
    torch.cuda.synchronize(); dist.barrier()

    def timed(fn, reps):
        torch.cuda.synchronize(); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / reps], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    # schedule choice (untimed): interior/boundary split with the halo in flight, or halo first
    # and one launch; thin slabs and a 25-60 us halo can favour the latter
    for ov in (False, True):   # first launches load kernels (the serial schedule runs the cell-tile kernel)
        system.step(overlap=ov, **fkw)
        system.step(overlap=ov, **fkw)
    ms_serial = timed(lambda: system.step(overlap=False, **fkw), 20)
    ms_overlap = timed(lambda: system.step(overlap=True, **fkw), 20)
    use_overlap = ms_overlap < ms_serial
    if use_overlap and system.tiles:
        # the overlapped schedule works on row sub-ranges, which the per-row kernels serve: do not pay
        # for a mirror nobody reads
        system.tiles = False
        system.rebuild()
    sampler = ClockSampler(local) if rank == 0 else None
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
    ms_halo = timed(lambda: torch.cuda.current_stream().wait_event(system.halo()), 20)
    clocks = sampler.finish() if sampler else None
    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        P = int(pairs.item())
        ms_total = float(ms.item())
        out = {
            "metric": metric, "value": P * K / (ms_total * 1e-3), "unit": unit, "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if args.prec == "fp64" else "f32-mixed", "data": "synthetic",
            "config": {"workload": "synthetic FCC lattice N=%d (%d cells/side) rho=%.1f cutoff=3.0 search=3.3, "
                                   "%d z-slabs, %d directed pairs, on-GPU list rebuild every %d steps, ghost "
                                   "positions exchanged every step (%s)" % (
                                       system.pn_global, n, density, world, P, rebuild_every, halo_mode),
                       "parallelism": "z-slab x%d" % world, "halo_layers": system.slab.halo,
                       "l2": "inputs larger than L2", "rebuild_every": rebuild_every},
            "gpu_launches": int(launches), "clocks": clocks,
            "halo": {"mode": halo_mode, "schedule": "overlap" if use_overlap else "halo-then-force",
                     "ms_step_overlap": ms_overlap, "ms_step_serial": ms_serial,
                     "ms_halo_alone": ms_halo,
                     "ghost_bytes_per_step_per_rank": int((system.slab.n_lo + system.slab.n_hi) * 32)},
            "e2e": {"value": P * K / (ms_total * 1e-3), "unit": unit, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0,
                    "note": "decomposed runs keep q,p resident; the host-buffer plugin call is the N=1 line"},
        }
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()
