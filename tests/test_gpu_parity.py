"""GPU parity tests: the CUDA path, called through the C ABI (lj_gpu_b200._capi), against the
CPU oracle on the same seeded inputs, and against the committed goldens of the reference.

Bars (BASELINE.json north_star): neighbour lists bit-exact after per-row sorting; momenta
within 1e-12 (FP64) / 1e-5 (mixed), norm-wise relative |dp|_max / |p|_max, after the
reference's LOOP = 100 steps.
"""
import hashlib
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden_rows

pytestmark = pytest.mark.gpu

TOL_FP64 = 1e-12
TOL_MIXED = 1e-5


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def torch():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


@pytest.fixture(scope="module")
def ctx(torch):
    from lj_gpu_b200 import LJContext
    c = LJContext(0)
    yield c
    c.close()


class System:
    """One reference configuration with its oracle results (computed once per module)."""

    def __init__(self, oracle, rho, L=50.0, steps=100):
        from oracle import ljoracle as lo
        self.rho, self.L, self.steps = rho, L, steps
        self.q = oracle.init_fcc(rho, L)
        self.pn = self.q.shape[0]
        self.full = oracle.makepair(self.q, full=True)
        self.half = oracle.makepair(self.q, full=False)
        self.p = np.zeros_like(self.q)
        oracle.force_gather(self.q, self.p, *self.full, steps=steps)
        self.scale = np.abs(self.p).max()
        self.lo = lo

    def device_arrays(self, torch, layout):
        q, pn = self.q, self.pn
        if layout == "aos3":
            qh = q.copy()
        elif layout == "aos4":
            qh = np.zeros((pn, 4)); qh[:, :3] = q; qh[:, 3] = -123.25  # .w is ignored
        else:
            stride = pn + 24
            qh = np.zeros((3, stride)); qh[:, :pn] = q.T
        qd = torch.from_numpy(qh).cuda()
        pd = torch.zeros_like(qd)
        if layout == "aos4":
            pd[:, 3] = 77.5  # .w of p must survive
        return qd, pd

    def p_host(self, pd, layout):
        a = pd.cpu().numpy()
        return a[:, :3] if layout != "soa" else a[:, :self.pn].T

    def err(self, pd, layout):
        return np.abs(self.p_host(pd, layout) - self.p).max() / self.scale


@pytest.fixture(scope="module")
def sysA(oracle):
    return System(oracle, 0.5)


@pytest.fixture(scope="module")
def sysB(oracle):
    return System(oracle, 1.0)


@pytest.fixture(scope="module")
def sysS(oracle):
    return System(oracle, 0.7, L=14.0, steps=7)


def list_to_host(pl):
    nop = pl.number_of_partners.cpu().numpy()
    ptr = pl.pointer.cpu().numpy().astype(np.int64)
    lst = pl.sorted_list.cpu().numpy()[:pl.number_of_pairs]
    return nop, ptr, lst


# ------------------------------------------------------------------------ generator
def test_init_fcc_bit_exact(oracle, golden):
    from lj_gpu_b200 import init_fcc
    for rho in (0.5, 1.0):
        q = init_fcc(rho, 50.0)
        assert sha(q) == str(golden(rho)["q_sha256"])
    assert np.array_equal(init_fcc(1.0, 20.3), oracle.init_fcc(1.0, 20.3))


# ------------------------------------------------------------------------ neighbour list
@pytest.mark.parametrize("which", ["A", "B"])
@pytest.mark.parametrize("half", [False, True])
@pytest.mark.parametrize("engine", ["cluster-search", "cluster-search+mirror", "per-particle-search"])
def test_list_bit_exact(ctx, torch, sysA, sysB, golden, which, half, engine):
    s = sysA if which == "A" else sysB
    qd, _ = s.device_arrays(torch, "aos4")
    pl = ctx.makepair(qd, half=half, clusters=engine.endswith("mirror"),
                      per_particle=engine.startswith("per"))  # every engine emits the same CSR list
    nop, ptr, lst = list_to_host(pl)
    nop_o, ptr_o, lst_o = s.half if half else s.full
    assert pl.number_of_pairs == len(lst_o)
    assert np.array_equal(nop, nop_o)
    assert np.array_equal(ptr, ptr_o)
    rows = s.lo.sort_rows(nop, ptr, lst)
    assert np.array_equal(rows, lst_o)
    assert pl.max_partners == nop_o.max()
    if half:  # straight against the real reference's half list (hash in the golden file)
        g = golden(s.rho)
        assert sha(nop) == str(g["nop_half_sha256"]) and sha(rows) == str(g["list_half_sha256"])
    ctx.check_loadedpair(pl)


@pytest.mark.parametrize("layout", ["aos3", "soa"])
def test_list_layouts_pointer64_sorted_rows(ctx, torch, sysS, layout):
    s = sysS
    qd, _ = s.device_arrays(torch, layout)
    pl = ctx.makepair(qd, layout=layout, pointer64=True, sort_rows=True,
                      pn=s.pn if layout == "soa" else None)
    nop, ptr, lst = list_to_host(pl)
    assert pl.pointer.dtype == torch.int64
    assert np.array_equal(nop, s.full[0]) and np.array_equal(ptr, s.full[1])
    assert np.array_equal(lst, s.full[2])  # LJ_LIST_SORT_ROWS: already ascending, like makepair()


def test_list_deterministic_and_rebuild(ctx, torch, sysA):
    qd, _ = sysA.device_arrays(torch, "aos4")
    a = ctx.makepair(qd)
    la = list_to_host(a)
    b = ctx.makepair(qd)
    lb = list_to_host(b)
    for x, y in zip(la, lb):
        assert np.array_equal(x, y)  # same row ORDER too: the build has no timing dependence
    b.sorted_list.zero_(); b.number_of_partners.zero_()
    ctx.rebuild(qd, b)  # asynchronous in-place rebuild
    assert ctx.list_result() == (a.number_of_pairs, a.max_partners)
    for x, y in zip(la, list_to_host(b)):
        assert np.array_equal(x, y)


def test_list_capacity_overflow_is_reported_not_written(ctx, torch, sysS):
    from lj_gpu_b200 import LJError, PairList, _capi
    qd, _ = sysS.device_arrays(torch, "aos4")
    need = len(sysS.full[2])
    cap, guard = need // 2, 4096
    buf = torch.full((cap + guard,), -7, dtype=torch.int32, device="cuda")
    small = PairList(torch.empty(sysS.pn, dtype=torch.int32, device="cuda"),
                     torch.empty(sysS.pn, dtype=torch.int32, device="cuda"), buf[:cap], 0, 0)
    with pytest.raises(LJError) as e:
        ctx.makepair(qd, out=small)
    assert e.value.status == _capi.LJ_ERR_CAPACITY
    total = ctx.lib.lj_launch_count(ctx.h)  # context still usable after the error
    assert total > 0
    assert torch.all(buf[cap:] == -7)  # nothing past the capacity was touched
    # the needed size is reported so the caller can re-allocate (makepair() does exactly that)
    assert ctx.makepair(qd).number_of_pairs == need


def test_list_threshold_semantics(ctx, torch, oracle):
    # listed iff r2 < SL2 (strict), decided in FP64 even where the FP32 pre-filter cannot tell
    sl = 3.3
    pts = np.array([[0, 0, 0, 0], [sl, 0, 0, 0],                          # r2 == SL2: not listed
                    [0, 50, 0, 0], [np.nextafter(sl, 0), 50, 0, 0],       # one ulp inside
                    [0, 0, 50, 0], [np.nextafter(sl, 10), 0, 50, 0],      # one ulp outside
                    [100, 100, 100, 0], [100 + 1e-9, 100, 100, 0],        # nearly coincident
                    [30, 30, 30, 0], [30 + 1.9052558883257651, 30 + 1.9052558883257651, 30 + 1.9052558883257651, 0]],
                   np.float64)
    qd = torch.from_numpy(pts).cuda()
    pl = ctx.makepair(qd)
    nop, ptr, lst = list_to_host(pl)
    nop_o, ptr_o, lst_o = oracle.makepair(pts[:, :3].copy(), full=True, brute=True)
    assert nop_o[0] == 0 and nop_o[2] == 1 and nop_o[4] == 0 and nop_o[6] == 1
    assert np.array_equal(nop, nop_o)
    from oracle import ljoracle as lo
    assert np.array_equal(lo.sort_rows(nop, ptr, lst), lst_o)


def test_list_edge_sizes(ctx, torch, oracle):
    for pts in (np.zeros((1, 4)), np.array([[0.0, 0, 0, 0], [1.2, 0, 0, 0]]),
                np.array([[0.0, 0, 0, 0], [10.0, 0, 0, 0]])):
        qd = torch.from_numpy(pts).cuda()
        pl = ctx.makepair(qd)
        nop_o, ptr_o, lst_o = oracle.makepair(pts[:, :3], full=True)
        assert pl.number_of_pairs == len(lst_o)
        assert np.array_equal(pl.number_of_partners.cpu().numpy()[:len(pts)], nop_o)
    # sparse, non-cubic cloud with a different search length
    rng = np.random.RandomState(3)
    q = oracle.init_fcc(1.0, 12.0)
    q = q[rng.rand(len(q)) < 0.5] * np.array([1.0, 0.4, 2.5]) + np.array([-7.0, 3.0, 0.125])
    q4 = np.zeros((len(q), 4)); q4[:, :3] = q
    pl = ctx.makepair(torch.from_numpy(q4).cuda(), search_len=2.1)
    nop_o, ptr_o, lst_o = oracle.makepair(q, search_len=2.1, full=True)
    nop, ptr, lst = list_to_host(pl)
    from oracle import ljoracle as lo
    assert np.array_equal(nop, nop_o) and np.array_equal(lo.sort_rows(nop, ptr, lst), lst_o)


def test_particle_order_without_spatial_coherence(ctx, torch, sysA, oracle):
    """The reference's commented-out std::shuffle(q) idea (cuda/force_cuda.cu:92-93): same lattice,
    particle order permuted.  Clusters of 4 consecutive particles are then spatially loose; list and
    forces must stay exact (and the build must not degenerate into an all-pairs scan)."""
    import time
    s = sysA
    perm = np.random.RandomState(123).permutation(s.pn)
    q = np.ascontiguousarray(s.q[perm])
    q4 = np.zeros((s.pn, 4)); q4[:, :3] = q
    qd = torch.from_numpy(q4).cuda()
    nop_o, ptr_o, lst_o = oracle.makepair(q, full=True)
    for kw in (dict(), dict(clusters=True), dict(per_particle=True)):
        torch.cuda.synchronize(); t0 = time.time()
        pl = ctx.makepair(qd, **kw)
        torch.cuda.synchronize()
        assert time.time() - t0 < 5.0
        nop, ptr, lst = list_to_host(pl)
        assert np.array_equal(nop, nop_o) and np.array_equal(s.lo.sort_rows(nop, ptr, lst), lst_o)
    p_o = np.zeros_like(q)
    oracle.force_gather(q, p_o, nop_o, ptr_o, lst_o, steps=5)
    pl = ctx.makepair(qd, clusters=True)
    for variant, group in (("subwarp", 8), ("cluster", 0), ("cluster", 32), ("tile", 8)):
        pd = torch.zeros_like(qd)
        ctx.force_loop(qd, pd, pl, loop=5, variant=variant, group=group)
        assert np.abs(pd.cpu().numpy()[:, :3] - p_o).max() / np.abs(p_o).max() < TOL_FP64, variant
    # momenta are a permutation of the lattice-order result
    assert np.abs(p_o - s.p[perm] * (5 / s.steps)).max() / s.scale < 1e-13


def test_row_range_build_for_ghosts(ctx, torch, sysS):
    # rows only for "owned" particles [r0,r1); all particles remain candidates
    s = sysS
    qd, _ = s.device_arrays(torch, "aos4")
    r0, r1 = s.pn // 5, s.pn // 2
    pl = ctx.makepair(qd, rows=(r0, r1))
    nop, ptr, lst = list_to_host(pl)
    nop_o, ptr_o, lst_o = s.full
    assert np.all(nop[:r0] == 0) and np.all(nop[r1:] == 0)
    assert np.array_equal(nop[r0:r1], nop_o[r0:r1])
    for i in (r0, (r0 + r1) // 2, r1 - 1):
        assert sorted(lst[ptr[i]:ptr[i] + nop[i]]) == lst_o[ptr_o[i]:ptr_o[i] + nop_o[i]].tolist()


# ------------------------------------------------------------------------ force, FP64
FORCE_CASES = [("aos4", "subwarp", 8), ("aos4", "warp", 32), ("aos4", "thread", 1),
               ("aos4", "subwarp", 4), ("aos4", "subwarp", 16), ("aos4", "subwarp", 2),
               ("aos4", "tile", 8), ("aos4", "tile", 32), ("aos4", "tile", 4),
               ("aos3", "subwarp", 8), ("aos3", "warp", 32), ("aos3", "tile", 16),
               ("soa", "subwarp", 8), ("soa", "thread", 1), ("soa", "tile", 8),
               ("aos4", "cluster", 0), ("aos3", "cluster", 0), ("soa", "cluster", 0),
               ("aos4", "cluster", 32), ("aos3", "cluster", 32), ("aos4", "cluster", 16), ("soa", "cluster", 16)]


@pytest.mark.parametrize("layout,variant,group", FORCE_CASES)
def test_force_fp64_config_A(ctx, torch, sysA, golden, layout, variant, group):
    s = sysA
    qd, pd = s.device_arrays(torch, layout)
    pn = s.pn if layout == "soa" else None
    pl = ctx.makepair(qd, layout=layout, pn=pn, clusters=(variant == "cluster"))
    ctx.force_loop(qd, pd, pl, loop=100, layout=layout, variant=variant, group=group, pn=pn)
    ctx.sync()
    assert s.err(pd, layout) < TOL_FP64
    ph = s.p_host(pd, layout)
    g = golden(0.5)  # full-precision sample dumped from the REAL reference (force_sorted x100)
    assert np.abs(ph[g["sample_idx"]] - g["p_sorted_sample"]).max() / float(g["p_sorted_absmax"]) < TOL_FP64
    rows = np.array(golden_rows(0.5))  # the published ref_data/density0.5.dat
    assert np.abs(np.vstack([ph[:5], ph[-5:]]) - rows).max() < 1e-10
    if layout == "aos4":
        assert torch.all(pd[:, 3] == 77.5)  # .w preserved


@pytest.mark.parametrize("layout,variant,group", [("aos4", "subwarp", 8), ("aos4", "tile", 8),
                                                   ("aos3", "warp", 32), ("soa", "subwarp", 16),
                                                   ("aos4", "cluster", 0), ("aos4", "auto", 0)])
def test_force_fp64_config_B(ctx, torch, sysB, golden, layout, variant, group):
    s = sysB
    qd, pd = s.device_arrays(torch, layout)
    pn = s.pn if layout == "soa" else None
    pl = ctx.makepair(qd, layout=layout, pn=pn, clusters=variant in ("cluster", "auto"))
    ctx.force_loop(qd, pd, pl, loop=100, layout=layout, variant=variant, group=group, pn=pn,
                   use_graph=True)
    ctx.sync()
    assert s.err(pd, layout) < TOL_FP64
    ph = s.p_host(pd, layout)
    g = golden(1.0)
    assert np.abs(ph[g["sample_idx"]] - g["p_sorted_sample"]).max() / float(g["p_sorted_absmax"]) < TOL_FP64
    from lj_gpu_b200 import print_results
    with open(os.path.join(GOLDEN, "density1.dat")) as f:
        assert print_results(ph) == f.read()  # byte-identical to ref_data/density1.dat


def test_int4_and_scalar_list_paths_agree(ctx, torch, sysS):
    # different lane -> entry mapping, so different summation order: agreement to rounding
    s = sysS
    for layout in ("aos4", "aos3", "soa"):
        qd, _ = s.device_arrays(torch, layout)
        pn = s.pn if layout == "soa" else None
        pl = ctx.makepair(qd, layout=layout, pn=pn)
        for group in (4, 8, 16, 32):
            a, b = torch.zeros_like(qd), torch.zeros_like(qd)
            ctx.force_loop(qd, a, pl, loop=s.steps, layout=layout, variant="subwarp", group=group, pn=pn,
                           list_scalar=2)
            ctx.force_loop(qd, b, pl, loop=s.steps, layout=layout, variant="subwarp", group=group, pn=pn)
            assert s.err(a, layout) < TOL_FP64
            assert s.err(b, layout) < TOL_FP64
            assert (a - b).abs().max().item() / s.scale < 1e-14, (layout, group)
    # a list that is NOT 16-byte aligned silently takes the scalar path
    qd, pd = s.device_arrays(torch, "aos4")
    pl = ctx.makepair(qd)
    shifted = torch.empty(pl.sorted_list.numel() + 1, dtype=torch.int32, device="cuda")
    shifted[1:] = pl.sorted_list
    pl.sorted_list = shifted[1:]
    ctx.force_loop(qd, pd, pl, loop=s.steps, variant="subwarp", group=8, list_scalar=2)
    assert s.err(pd, "aos4") < TOL_FP64


def test_force_pointer64_and_thread_block(ctx, torch, sysS):
    s = sysS
    qd, _ = s.device_arrays(torch, "aos4")
    pl = ctx.makepair(qd, pointer64=True)
    for tb in (64, 256, 1024):
        for variant, group in (("subwarp", 8), ("tile", 8), ("warp", 32)):
            pd = torch.zeros_like(qd)
            ctx.force_loop(qd, pd, pl, loop=s.steps, variant=variant, group=group, threads_per_block=tb)
            assert s.err(pd, "aos4") < TOL_FP64, (tb, variant)


def test_force_ell(ctx, torch, sysA, oracle):
    s = sysA
    qd, pd = s.device_arrays(torch, "aos4")
    pl = ctx.makepair(qd, sort_rows=True)
    tl = ctx.make_transposed_pairlist(pl)
    tl_o, max_np = oracle.transpose_list(s.full[2], s.full[0], s.full[1])
    assert pl.max_partners == max_np
    assert np.array_equal(tl.cpu().numpy()[:max_np * s.pn], tl_o[:max_np * s.pn])  # zero padded too
    ctx.force_loop(qd, pd, pl, loop=100, ell=True)
    assert s.err(pd, "aos4") < TOL_FP64
    q3, p3 = s.device_arrays(torch, "aos3")
    ctx.force_loop(q3, p3, pl, loop=100, ell=True, layout="aos3")
    assert s.err(p3, "aos3") < TOL_FP64


@pytest.mark.parametrize("layout,group", [("aos4", 8), ("aos4", 32), ("aos3", 1), ("soa", 8)])
def test_force_newton3_half_list(ctx, torch, sysA, layout, group):
    s = sysA
    qd, pd = s.device_arrays(torch, layout)
    pn = s.pn if layout == "soa" else None
    pl = ctx.makepair(qd, half=True, layout=layout, pn=pn)
    assert pl.number_of_pairs == 2268138  # P_half of config A
    ctx.force_loop(qd, pd, pl, loop=100, layout=layout, variant="n3", group=group, pn=pn)
    assert s.err(pd, layout) < TOL_FP64


@pytest.mark.parametrize("which", ["A", "B"])
@pytest.mark.parametrize("layout,group", [("aos4", 8), ("aos4", 32), ("aos3", 4), ("soa", 8), ("aos4", -1),
                                          ("soa", -1)])
def test_force_mixed(ctx, torch, sysA, sysB, which, layout, group):
    s = sysA if which == "A" else sysB
    qd, pd = s.device_arrays(torch, layout)
    pn = s.pn if layout == "soa" else None
    cluster = group < 0   # group -1: the cluster-list mixed kernel
    pl = ctx.makepair(qd, layout=layout, pn=pn, clusters=cluster)
    ctx.force_loop(qd, pd, pl, loop=100, layout=layout, group=max(group, 0), precision="mixed", pn=pn,
                   variant="cluster" if cluster else "auto")
    e = s.err(pd, layout)
    assert e < TOL_MIXED, e
    assert e > 1e-14  # it really is the FP32 path


@pytest.mark.parametrize("which", ["A", "B"])
def test_float4_layout_mixed(ctx, torch, sysA, sysB, oracle, which):
    """The reference's q_f4 / p_f4 buffers (cuda/force_cuda.cu:25, never timed there): float4 in,
    float4 out, FP32 pair math.  Checker = the FP64 oracle on the SAME float-valued positions."""
    s = sysA if which == "A" else sysB
    qf = np.zeros((s.pn, 4), np.float32); qf[:, :3] = s.q.astype(np.float32); qf[:, 3] = 3.5
    q64 = np.ascontiguousarray(qf[:, :3].astype(np.float64))
    nop_o, ptr_o, lst_o = oracle.makepair(q64, full=True)
    p_o = np.zeros_like(q64)
    oracle.force_gather(q64, p_o, nop_o, ptr_o, lst_o, steps=100)
    qd = torch.from_numpy(qf).cuda()
    pd = torch.zeros_like(qd); pd[:, 3] = -2.0
    pl = ctx.makepair(qd, clusters=True)                      # list build straight from float4
    nop, ptr, lst = list_to_host(pl)
    assert np.array_equal(nop, nop_o) and np.array_equal(s.lo.sort_rows(nop, ptr, lst), lst_o)
    for variant in ("auto", "cluster"):
        pd[:, :3] = 0
        ctx.force_loop(qd, pd, pl, loop=100, precision="mixed", variant=variant, group=0)
        got = pd.cpu().numpy()
        assert np.abs(got[:, :3] - p_o).max() / np.abs(p_o).max() < TOL_MIXED
        assert np.all(got[:, 3] == -2.0)
    from lj_gpu_b200 import LJError
    with pytest.raises(LJError):
        ctx.force_step(qd, pd, pl, precision="fp64")           # float4 is a mixed-precision layout


def test_row_order_is_irrelevant_and_gather_is_reproducible(ctx, torch, sysA):
    s = sysA
    qd, _ = s.device_arrays(torch, "aos4")
    pl = ctx.makepair(qd)
    ref = torch.zeros_like(qd)
    ctx.force_loop(qd, ref, pl, loop=5, group=8)
    again = torch.zeros_like(qd)
    ctx.force_loop(qd, again, pl, loop=5, group=8)
    assert torch.equal(ref, again)  # no atomics in the gather path: bit-reproducible
    before = s.lo.sort_rows(*list_to_host(pl))
    ctx.random_shfl(pl, seed=10)
    nop, ptr, lst = list_to_host(pl)
    assert not np.array_equal(lst, s.full[2])
    assert np.array_equal(s.lo.sort_rows(nop, ptr, lst), before)  # a permutation within rows
    for variant, group in (("subwarp", 8), ("tile", 8), ("warp", 32)):
        pd = torch.zeros_like(qd)
        ctx.force_loop(qd, pd, pl, loop=5, variant=variant, group=group)
        assert (pd - ref)[:, :3].abs().max().item() / ref[:, :3].abs().max().item() < 1e-13


def test_force_cutoff_semantics(ctx, torch, oracle):
    # r2 == CL2 contributes, r2 just above does not (cuda/kernel.cuh:30 `if (r2 > CL2) df = 0`)
    pts = np.array([[0, 0, 0, 0], [3.0, 0, 0, 0], [0, 40, 0, 0], [np.nextafter(3.0, 4), 40, 0, 0]], np.float64)
    qd = torch.from_numpy(pts).cuda()
    pl = ctx.makepair(qd)
    nop, ptr, lst = oracle.makepair(pts[:, :3], full=True)
    po = np.zeros((4, 3))
    oracle.force_gather(pts[:, :3].copy(), po, nop, ptr, lst)
    assert po[0, 0] != 0 and po[2, 0] == 0
    for variant, group, prec in (("subwarp", 8, "fp64"), ("tile", 8, "fp64"), ("thread", 1, "fp64"),
                                 ("subwarp", 8, "mixed")):
        pd = torch.zeros_like(qd)
        ctx.force_step(qd, pd, pl, variant=variant, group=group, precision=prec)
        ph = pd.cpu().numpy()[:, :3]
        assert ph[2, 0] == 0 and ph[3, 0] == 0
        assert abs(ph[0, 0] - po[0, 0]) <= (1e-14 if prec == "fp64" else 1e-6) * abs(po[0, 0])
        assert ph[1, 0] == -ph[0, 0]


def test_force_row_range(ctx, torch, sysS):
    s = sysS
    qd, pd = s.device_arrays(torch, "aos4")
    pl = ctx.makepair(qd)
    r0, r1 = 100, s.pn - 333
    for variant in ("subwarp", "tile"):
        pd.zero_()
        ctx.force_loop(qd, pd, pl, loop=s.steps, variant=variant, group=8, rows=(r0, r1))
        ph = pd.cpu().numpy()[:, :3]
        assert np.all(ph[:r0] == 0) and np.all(ph[r1:] == 0)
        assert np.abs(ph[r0:r1] - s.p[r0:r1]).max() / s.scale < TOL_FP64


def test_cluster_list_identity_row_ranges_and_invalidation(ctx, torch, sysS):
    from lj_gpu_b200 import LJError
    s = sysS
    qd, pd = s.device_arrays(torch, "aos4")
    plain = ctx.makepair(qd)
    with pytest.raises(LJError):                       # no mirror for these arrays
        ctx.force_step(qd, pd, plain, variant="cluster")
    pl = ctx.makepair(qd, clusters=True)
    launches = ctx.launches
    ctx.force_loop(qd, pd, pl, loop=s.steps, variant="cluster")
    assert ctx.launches - launches == s.steps and s.err(pd, "aos4") < TOL_FP64
    with pytest.raises(LJError):                       # the mirror belongs to pl, not to plain
        ctx.force_step(qd, pd, plain, variant="cluster")
    # row sub-ranges on 4-row boundaries use the mirror, others fall back to the per-row kernel
    for r0, r1 in ((128, s.pn - 400), (0, s.pn), (4, 8), (s.pn - 4, s.pn)):
        pd.zero_()
        ctx.force_loop(qd, pd, pl, loop=s.steps, variant="cluster", rows=(r0, r1))
        ph = pd.cpu().numpy()[:, :3]
        assert np.all(ph[:r0] == 0) and np.all(ph[r1:] == 0)
        assert np.abs(ph[r0:r1] - s.p[r0:r1]).max() / s.scale < TOL_FP64
    pd.zero_()
    ctx.force_loop(qd, pd, pl, loop=s.steps, variant="auto", rows=(3, s.pn - 2))   # AUTO: per-row kernel
    ph = pd.cpu().numpy()[:, :3]
    assert np.abs(ph[3:s.pn - 2] - s.p[3:s.pn - 2]).max() / s.scale < TOL_FP64 and np.all(ph[:3] == 0)
    with pytest.raises(LJError):
        ctx.force_step(qd, pd, pl, variant="cluster", rows=(3, s.pn - 2))
    # partial row ranges at build time (ghost particles as candidates only)
    own = (s.pn // 3) // 4 * 4 + 2
    plr = ctx.makepair(qd, clusters=True, rows=(0, own))
    pd.zero_()
    ctx.force_loop(qd, pd, plr, loop=s.steps, variant="cluster", rows=(0, own))
    ph = pd.cpu().numpy()[:, :3]
    assert np.abs(ph[:own] - s.p[:own]).max() / s.scale < TOL_FP64 and np.all(ph[own:] == 0)
    # a shuffle or an explicit invalidate drops the mirror
    pl2 = ctx.makepair(qd, clusters=True)
    ctx.random_shfl(pl2)
    with pytest.raises(LJError):
        ctx.force_step(qd, pd, pl2, variant="cluster")
    pl3 = ctx.makepair(qd, clusters=True)
    ctx.list_invalidate()
    with pytest.raises(LJError):
        ctx.force_step(qd, pd, pl3, variant="cluster")
    pd.zero_()
    ctx.force_loop(qd, pd, pl3, loop=s.steps)          # AUTO still works (per-row kernel)
    assert s.err(pd, "aos4") < TOL_FP64


def test_bad_arguments_are_rejected(ctx, torch, sysS):
    from lj_gpu_b200 import LJError
    from lj_gpu_b200 import _capi
    qd, pd = sysS.device_arrays(torch, "aos4")
    pl = ctx.makepair(qd)
    with pytest.raises(LJError) as e:
        ctx.force_step(qd, pd, pl, threads_per_block=48)  # reference CLI range is 64..1024
    assert e.value.status == _capi.LJ_ERR_BAD_ARG
    with pytest.raises(LJError):
        ctx.force_step(qd, pd, pl, group=3)
    mis = torch.zeros(sysS.pn * 4 + 1, dtype=torch.float64, device="cuda")[1:].view(sysS.pn, 4)
    with pytest.raises(LJError):
        ctx.force_step(mis, pd, pl)  # double4 must be 32-byte aligned
    with pytest.raises(LJError):
        ctx.makepair(qd, search_len=-1.0)
    bad = ctx.makepair(qd)
    bad.sorted_list[5] = sysS.pn + 3
    with pytest.raises(LJError) as e:
        ctx.check_loadedpair(bad)
    assert e.value.status == _capi.LJ_ERR_INVALID_LIST


# ------------------------------------------------------------------------ measure() / memory layer
def test_measure_host_buffers_reproduce_goldens(ctx, torch, sysA):
    from lj_gpu_b200 import print_results
    s = sysA
    for layout, kw in (("aos3", dict(variant="warp")), ("aos4", dict(variant="tile", group=8, rebuild_every=20)),
                       ("aos4", dict(variant="auto", use_graph=True, rebuild_every=20))):
        w = 3 if layout == "aos3" else 4
        qh = np.zeros((s.pn, w)); qh[:, :3] = s.q
        ph = np.zeros((s.pn, w))
        m = ctx.measure(qh, ph, layout=layout, loop=100, **kw)
        assert m.number_of_pairs == 4536276 and m.max_partners == 78
        assert m.list_builds == (5 if kw.get("rebuild_every") else 1)
        with open(os.path.join(GOLDEN, "density0.5.dat")) as f:
            assert print_results(ph) == f.read()
        assert np.abs(ph[:, :3] - s.p).max() / s.scale < TOL_FP64
        assert m.h2d_bytes == 2 * qh.nbytes and m.d2h_bytes == ph.nbytes
    # the reference's own flow: host-built list uploaded by copy_to_gpu; half list -> Newton-3
    nop, ptr, lst = s.half
    qh = np.zeros((s.pn, 4)); qh[:, :3] = s.q
    ph = np.zeros((s.pn, 4))
    m = ctx.measure(qh, ph, layout="aos4", loop=100, half=True, host_list=(nop, ptr.astype(np.int32), lst))
    assert m.list_builds == 0 and m.number_of_pairs == len(lst)
    assert np.abs(ph[:, :3] - s.p).max() / s.scale < TOL_FP64


def test_cuda_ptr_semantics(ctx, torch):
    # the reference's disabled unit test (cuda/cuda_ptr.cuh:106-147): copy round trip, set_val ranges
    n = 256
    fl = ctx.cuda_ptr(np.float32, n)
    fl.host[:] = 3.0
    fl.host2dev()
    fl.host[:] = 0.0
    fl.dev2host()
    ctx.sync()
    assert np.all(fl.host == 3.0)
    fl.set_val(12.0)
    fl.set_val(0.0, 10, 87)
    expect = np.full(n, 12.0, np.float32); expect[10:97] = 0.0
    assert np.array_equal(fl.host, expect)  # host mirror
    fl.host[:] = -1
    fl.dev2host(5, 100)
    ctx.sync()
    assert np.array_equal(fl.host[5:105], expect[5:105]) and np.all(fl.host[:5] == -1)
    fl.deallocate()
    # pageable upload/download through the staging ring, larger than one ring half
    from lj_gpu_b200 import _capi
    import ctypes as C
    big = np.arange(12 * 1024 * 1024, dtype=np.int64)
    d = torch.empty(big.size, dtype=torch.int64, device="cuda")
    lib = _capi.load()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert lib.lj_upload(ctx.h, d.data_ptr(), big.ctypes.data, big.nbytes, st) == 0
    back = np.zeros_like(big)
    assert lib.lj_download(ctx.h, back.ctypes.data, d.data_ptr(), big.nbytes, st) == 0
    ctx.sync()
    assert np.array_equal(back, big) and torch.equal(d[-3:].cpu(), torch.from_numpy(big[-3:]))


def test_cpp_driver_prints_the_goldens():
    exe = os.path.join(ROOT, "lj_gpu_b200", "driver", "force_b200")
    if not os.path.exists(exe):
        pytest.skip("driver not built")
    for args, gold in ((["--test"], "density0.5.dat"), (["--test", "--density", "1.0", "256"], "density1.dat")):
        r = subprocess.run([exe] + args, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        with open(os.path.join(GOLDEN, gold)) as f:
            assert r.stdout == f.read()
        assert "[sec] (without Host<->Device)" in r.stderr and "force_kernel_warp_unroll2_double3" in r.stderr
    r = subprocess.run([exe, "32"], capture_output=True, text=True)
    assert r.returncode == 1 and "THREAD_BLOCK size is too large or small." in r.stderr
    # the OpenACC SoA program on six separate arrays: CSR run, then transposed-list run
    r = subprocess.run([exe, "--soa6"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    with open(os.path.join(GOLDEN, "density0.5.dat")) as f:
        gold = f.read()
    assert r.stdout == gold + gold
    assert "acc_reactless_soa" in r.stderr and "acc_reactless_memopt_soa" in r.stderr


def test_cpp_driver_pair_cache(tmp_path, oracle):
    """--cache: first run builds on the GPU and writes the reference's text cache, second run loads
    it (the reference's own flow: host list uploaded) and prints the same goldens."""
    from lj_gpu_b200 import loadpair
    exe = os.path.join(ROOT, "lj_gpu_b200", "driver", "force_b200")
    if not os.path.exists(exe):
        pytest.skip("driver not built")
    gold = open(os.path.join(GOLDEN, "density0.5.dat")).read()
    r1 = subprocess.run([exe, "--test", "--cache"], capture_output=True, text=True, timeout=300, cwd=tmp_path)
    assert r1.returncode == 0, r1.stderr
    assert r1.stdout == gold and "Now make pairlist .cache_pair_all.dat." in r1.stderr
    nop, ptr, lst = loadpair(str(tmp_path / ".cache_pair_all.dat"), 62500)
    q = oracle.init_fcc(0.5, 50.0)
    nop_o, ptr_o, lst_o = oracle.makepair(q, full=True)
    assert np.array_equal(nop, nop_o) and np.array_equal(ptr, ptr_o) and np.array_equal(lst, lst_o)
    r2 = subprocess.run([exe, "--test", "--cache"], capture_output=True, text=True, timeout=300, cwd=tmp_path)
    assert r2.returncode == 0, r2.stderr
    assert r2.stdout == gold and ".cache_pair_all.dat is successfully loaded." in r2.stderr


# ------------------------------------------------------------------------ full size (config C)
def test_full_size_1M_properties_and_oracle(ctx, torch, oracle):
    """BASELINE config 3: FCC rho=1.0, 63 cells/side -> N=1,000,188.  Full oracle comparison
    for the list (cell-list restatement) and 3 force steps, plus size-independent properties."""
    from lj_gpu_b200 import init_fcc
    from oracle import ljoracle as lo
    q = init_fcc(1.0, 100.1)
    pn = q.shape[0]
    assert pn == 1000188
    q4 = np.zeros((pn, 4)); q4[:, :3] = q
    qd = torch.from_numpy(q4).cuda()
    half = ctx.makepair(qd, half=True)
    full = ctx.makepair(qd, clusters=True)
    assert full.number_of_pairs == 2 * half.number_of_pairs          # every pair listed both ways
    nop_o, ptr_o, lst_o = oracle.makepair(q, full=True)
    nop, ptr, lst = list_to_host(full)
    assert full.number_of_pairs == len(lst_o)
    assert np.array_equal(nop, nop_o) and np.array_equal(ptr, ptr_o)
    assert np.array_equal(lo.sort_rows(nop, ptr, lst), lst_o)        # bit-exact, all 1.4e8 entries
    p_o = np.zeros_like(q)
    oracle.force_gather(q, p_o, nop_o, ptr_o, lst_o, steps=3)
    scale = np.abs(p_o).max()
    results = {}
    for variant, group in (("subwarp", 8), ("tile", 8), ("warp", 32), ("cluster", 0)):
        pd = torch.zeros_like(qd)
        ctx.force_loop(qd, pd, full, loop=3, variant=variant, group=group)
        ph = pd.cpu().numpy()[:, :3]
        assert np.abs(ph - p_o).max() / scale < TOL_FP64
        results[variant] = pd
        # Newton's third law: the directed contributions cancel pairwise -> total momentum ~ 0
        assert np.abs(ph.sum(axis=0)).max() < 1e-9 * scale * np.sqrt(pn)
    # linearity in the step count with static positions: p(6 steps) == 2 * p(3 steps)
    pd6 = torch.zeros_like(qd)
    ctx.force_loop(qd, pd6, full, loop=6, variant="subwarp", group=8)
    assert (pd6 - 2 * results["subwarp"])[:, :3].abs().max().item() / scale < 1e-14
    # half list + Newton-3 and mixed precision at full size
    pn3 = torch.zeros_like(qd)
    ctx.force_loop(qd, pn3, half, loop=3, variant="n3", group=8)
    assert np.abs(pn3.cpu().numpy()[:, :3] - p_o).max() / scale < TOL_FP64
    for variant in ("subwarp", "cluster"):
        pmx = torch.zeros_like(qd)
        ctx.force_loop(qd, pmx, full, loop=3, group=8, precision="mixed", variant=variant)
        assert np.abs(pmx.cpu().numpy()[:, :3] - p_o).max() / scale < TOL_MIXED


# ------------------------------------------------------------------------ six-array SoA
def test_six_array_soa_interface_of_the_openacc_program(ctx, torch, sysS):
    """openacc/force_oacc_soa.cpp keeps qx,qy,qz,px,py,pz as six separate allocations (:17-22):
    list build and force loop on them, separately allocated (gathered into a library-owned block)
    and as equally spaced views of one allocation (read in place); CSR, ELL, cell-tile, mixed."""
    s = sysS
    qs = [torch.from_numpy(np.ascontiguousarray(s.q[:, c])).cuda() for c in range(3)]
    junk = torch.empty(1237, device="cuda")                       # breaks any regular spacing
    ps = [torch.zeros(s.pn, dtype=torch.float64, device="cuda") for _ in range(3)]
    pl = ctx.makepair_soa6(*qs, tiles=True)
    nop, ptr, lst = list_to_host(pl)
    assert pl.number_of_pairs == len(s.full[2]) and np.array_equal(nop[:s.pn], s.full[0])
    assert np.array_equal(s.lo.sort_rows(nop, ptr, lst), s.full[2])
    for kw, tol in ((dict(variant="subwarp", group=8), TOL_FP64), (dict(variant="celltile"), TOL_FP64),
                    (dict(variant="celltile", precision="mixed"), TOL_MIXED)):
        for p in ps:
            p.zero_()
        ctx.force_loop_soa6(*qs, *ps, pl, loop=s.steps, **kw)
        ph = torch.stack(ps, dim=1).cpu().numpy()
        assert np.abs(ph - s.p).max() / s.scale < tol, kw
    ctx.make_transposed_pairlist(pl)
    for p in ps:
        p.zero_()
    ctx.force_loop_soa6(*qs, *ps, pl, loop=s.steps, ell=True)      # force_reactless_memopt
    assert np.abs(torch.stack(ps, dim=1).cpu().numpy() - s.p).max() / s.scale < TOL_FP64
    # one allocation, equal spacing: no copies (the p views are updated in place)
    stride = s.pn + 8
    qb = torch.zeros(3 * stride, dtype=torch.float64, device="cuda")
    pb = torch.zeros(3 * stride, dtype=torch.float64, device="cuda")
    qv = [qb[c * stride:c * stride + s.pn] for c in range(3)]
    pv = [pb[c * stride:c * stride + s.pn] for c in range(3)]
    for c in range(3):
        qv[c].copy_(qs[c])
    pl2 = ctx.makepair_soa6(*qv)
    assert pl2.number_of_pairs == pl.number_of_pairs
    ctx.force_loop_soa6(*qv, *pv, pl2, loop=s.steps, variant="subwarp", group=8)
    assert np.abs(torch.stack(pv, dim=1).cpu().numpy() - s.p).max() / s.scale < TOL_FP64
    del junk


# ------------------------------------------------------------------------ cell-tile mirror
def _rows_as_sets(nop, ptr, lst, pn):
    rows = np.repeat(np.arange(pn, dtype=np.int64), nop[:pn])
    order = np.lexsort((lst, rows))
    return rows[order] * pn + lst[order]


@pytest.mark.parametrize("layout", ["aos4", "aos3", "soa"])
def test_celltile_force_and_list_match_oracle(ctx, torch, sysS, layout):
    """lj_build_list(LJ_LIST_TILES) + LJ_VARIANT_CELLTILE: the list the tile fill pass writes is the
    oracle's list, the shared-memory force kernel reproduces the oracle's momenta, and (same list
    order, same lane mapping) its result is bit-identical to the per-row kernel with group = 8."""
    s = sysS
    qd, pd = s.device_arrays(torch, layout)
    npn = s.pn if layout == "soa" else None
    pl = ctx.makepair(qd, layout=layout, pn=npn, tiles=True)
    nop, ptr, lst = list_to_host(pl)
    assert pl.number_of_pairs == len(s.full[2])
    assert np.array_equal(nop[:s.pn], s.full[0]) and np.array_equal(ptr[:s.pn], s.full[1])
    assert np.array_equal(_rows_as_sets(nop, ptr, lst, s.pn), _rows_as_sets(s.full[0], s.full[1], s.full[2], s.pn))
    launches = ctx.launches
    ctx.force_loop(qd, pd, pl, loop=s.steps, layout=layout, pn=npn, variant="celltile")
    assert ctx.launches - launches == 2 * s.steps          # position permute + force kernel
    assert s.err(pd, layout) < TOL_FP64
    pr = torch.zeros_like(pd)
    if layout == "aos4":
        pr[:, 3] = 77.5
        assert float(pd[:, 3].min()) == 77.5 == float(pd[:, 3].max())   # .w of p survives
    ctx.force_loop(qd, pr, pl, loop=s.steps, layout=layout, pn=npn, variant="subwarp", group=8)
    assert torch.equal(pd, pr)


@pytest.mark.parametrize("layout", ["aos4", "aos3", "soa"])
def test_celltile_mixed_precision_matches_oracle(ctx, torch, sysS, layout):
    """LJ_VARIANT_CELLTILE + LJ_PREC_MIXED: fixed-point records in shared memory, FP32 pair
    arithmetic, FP64 momenta; within the stated 1e-5 of the oracle, .w of p untouched."""
    s = sysS
    qd, pd = s.device_arrays(torch, layout)
    npn = s.pn if layout == "soa" else None
    pl = ctx.makepair(qd, layout=layout, pn=npn, tiles=True)
    launches = ctx.launches
    ctx.force_loop(qd, pd, pl, loop=s.steps, layout=layout, pn=npn, variant="celltile", precision="mixed")
    assert ctx.launches - launches == 2 * s.steps          # fixed-point permute + force kernel
    err = s.err(pd, layout)
    assert 0 < err < TOL_MIXED                              # FP32 arithmetic really ran
    if layout == "aos4":
        assert float(pd[:, 3].min()) == 77.5 == float(pd[:, 3].max())
    # the CUDA-graph replay of the same loop gives the same bits
    pg = torch.zeros_like(pd)
    if layout == "aos4":
        pg[:, 3] = 77.5
    ctx.force_loop(qd, pg, pl, loop=s.steps, layout=layout, pn=npn, variant="celltile", precision="mixed",
                   use_graph=True)
    assert torch.equal(pg, pd)


def test_celltile_mixed_precision_borderline_pairs_are_decided_in_fp64(ctx, torch, oracle, sysS):
    """Pairs at r2 == CL2 (contributes, cuda/kernel.cuh:30) and a hair outside (does not): FP32
    on fixed-point coordinates cannot tell them apart, the kernel re-decides them from the FP64
    positions.  A misclassified pair would show as ~1e-3 relative, far above the 1e-5 bound."""
    s = sysS
    a = np.array([-7.0, -7.0, -7.0])
    extra = np.stack([a, a + [3.0, 0.0, 0.0], a - [0.0, 3.0 + 4e-9, 0.0], a + [0.0, 0.0, 3.0 - 4e-9]])
    q = np.ascontiguousarray(np.concatenate([s.q, extra]))
    pn = len(q)
    nop_o, ptr_o, lst_o = oracle.makepair(q, full=True)
    p_o = np.zeros_like(q)
    oracle.force_gather(q, p_o, nop_o, ptr_o, lst_o, steps=s.steps)
    assert p_o[pn - 4, 0] != 0 and p_o[pn - 4, 1] == 0 and p_o[pn - 4, 2] != 0   # in, out, in
    qd = torch.from_numpy(q).cuda()
    pl = ctx.makepair(qd, layout="aos3", tiles=True)
    for prec, tol in (("fp64", TOL_FP64), ("mixed", TOL_MIXED)):
        pd = torch.zeros_like(qd)
        ctx.force_loop(qd, pd, pl, loop=s.steps, layout="aos3", variant="celltile", precision=prec)
        ph = pd.cpu().numpy()
        assert np.abs(ph - p_o).max() / np.abs(p_o).max() < tol
        # the four extra atoms only see each other: their momenta are small, check them on their own scale
        assert np.abs(ph[pn - 4:] - p_o[pn - 4:]).max() / np.abs(p_o[pn - 4:]).max() < 100 * tol
        assert ph[pn - 4, 1] == 0.0


def test_celltile_random_cloud_and_clusters_of_particles(ctx, torch, oracle):
    """No lattice: a uniform random cloud with a denser blob and an empty region (tiles with very
    different row counts, empty cells, empty tiles), particle order without spatial coherence."""
    from lj_gpu_b200 import LJError
    rng = np.random.default_rng(12)
    a = rng.uniform(0.0, 18.0, size=(2600, 3))
    a = a[~((a[:, 0] > 6) & (a[:, 0] < 11) & (a[:, 1] < 9))]          # a hole
    blob = rng.normal(loc=(14.0, 14.0, 4.0), scale=1.6, size=(220, 3))
    q = np.ascontiguousarray(np.concatenate([a, blob]))
    rng.shuffle(q)
    pn = len(q)
    nop_o, ptr_o, lst_o = oracle.makepair(q, full=True)
    qd = torch.from_numpy(q).cuda()
    pl = ctx.makepair(qd, layout="aos3", tiles=True)
    nop, ptr, lst = list_to_host(pl)
    assert np.array_equal(nop[:pn], nop_o)
    assert np.array_equal(_rows_as_sets(nop, ptr, lst, pn), _rows_as_sets(nop_o, ptr_o, lst_o, pn))
    # forces: the two GPU kernels bit for bit on this list (close random pairs make huge but finite
    # numbers; both kernels evaluate the same expression in the same order)
    p1 = torch.zeros_like(qd); p2 = torch.zeros_like(qd)
    ctx.force_step(qd, p1, pl, layout="aos3", variant="celltile", dt=1e-9)
    ctx.force_step(qd, p2, pl, layout="aos3", variant="subwarp", group=8, dt=1e-9)
    assert torch.equal(p1, p2)
    # far too dense for the shared-memory rings (rows of > 1000 partners): the build still delivers
    # the list, there is silently no mirror, AUTO runs the per-row kernel
    blob = rng.normal(loc=(9.0, 9.0, 9.0), scale=0.8, size=(900, 3))
    q = np.ascontiguousarray(np.concatenate([a, blob]))
    nop_o, ptr_o, lst_o = oracle.makepair(q, full=True)
    qd = torch.from_numpy(q).cuda()
    pl = ctx.makepair(qd, layout="aos3", tiles=True)
    nop, ptr, lst = list_to_host(pl)
    assert np.array_equal(_rows_as_sets(nop, ptr, lst, len(q)), _rows_as_sets(nop_o, ptr_o, lst_o, len(q)))
    p1 = torch.zeros_like(qd)
    with pytest.raises(LJError):
        ctx.force_step(qd, p1, pl, layout="aos3", variant="celltile", dt=1e-9)
    ctx.force_step(qd, p1, pl, layout="aos3", dt=1e-9)


def test_tile_engine_rows_longer_than_the_staging_area(ctx, torch, oracle):
    """Search length 4.4 at rho = 1.0: up to 368 partners per row, more than the 256 entries per row
    the replay pass of the tile engine stages in shared memory at most (longer rows are written
    directly; where a (row, pencil) window exceeds 64 records the round-1 engine serves the build) and
    46 trips per row in the cell-tile force kernel.  List against the oracle, both force kernels on it
    against the oracle's gather and against each other bit for bit."""
    from lj_gpu_b200 import init_fcc
    q = init_fcc(1.0, 15.0)
    pn = len(q)
    search, cutoff = 4.4, 4.0
    nop_o, ptr_o, lst_o = oracle.makepair(q, search_len=search, full=True)
    assert nop_o.max() > 256
    q4 = np.zeros((pn, 4)); q4[:, :3] = q
    qd = torch.from_numpy(q4).cuda()
    pl = ctx.makepair(qd, search_len=search, tiles=True)
    nop, ptr, lst = list_to_host(pl)
    assert np.array_equal(nop[:pn], nop_o)
    assert np.array_equal(_rows_as_sets(nop, ptr, lst, pn), _rows_as_sets(nop_o, ptr_o, lst_o, pn))
    p1 = torch.zeros_like(qd); p2 = torch.zeros_like(qd)
    ctx.force_loop(qd, p1, pl, loop=5, variant="celltile", cl2=cutoff * cutoff)
    ctx.force_loop(qd, p2, pl, loop=5, variant="subwarp", group=8, cl2=cutoff * cutoff)
    assert torch.equal(p1, p2)
    po = np.zeros((pn, 3))
    oracle.force_gather(q, po, nop_o, ptr_o, lst_o, steps=5, cl2=cutoff * cutoff, static_q=True)
    assert np.abs(p1.cpu().numpy()[:, :3] - po).max() / np.abs(po).max() < TOL_FP64


def test_celltile_moving_particles_rebuild_and_row_ranges(ctx, torch, sysS):
    """The mirror is a LIST: it stays valid while particles move (positions are re-permuted every
    step), it is rebuilt with the list, and it serves exactly the row range it was built for."""
    from lj_gpu_b200 import LJError
    s = sysS
    qd, pd = s.device_arrays(torch, "aos4")
    qr, pr = qd.clone(), pd.clone()
    # the reference run uses the per-row kernel on a list of its own from the same engine: the order of a row's
    # entries is unspecified by contract and differs between the engines, bit equality needs the same order
    plr = ctx.makepair(qr, tiles=True)
    pl = ctx.makepair(qd, tiles=True)                    # (one mirror per context: this build replaces plr's)
    for step in range(12):                                # kick + drift, list reused (skin 0.3)
        ctx.force_step(qd, pd, pl, variant="celltile")
        ctx.drift(qd, pd)
        ctx.force_step(qr, pr, plr, variant="subwarp", group=8)
        ctx.drift(qr, pr)
        if step == 5:                                     # rebuild both lists in place mid-run
            ctx.rebuild(qr, plr, tiles=True)
            ctx.rebuild(qd, pl, tiles=True)
    assert torch.equal(qd, qr) and torch.equal(pd, pr)
    # rebuild() reuses the flags the list was built with (the mirror is rebuilt with the list) ...
    ctx.rebuild(qd, pl)
    ctx.force_step(qd, pd, pl, variant="celltile")
    # ... a rebuild with the flag switched off leaves no mirror: explicit request fails, AUTO falls back
    ctx.rebuild(qd, pl, tiles=False)
    with pytest.raises(LJError):
        ctx.force_step(qd, pd, pl, variant="celltile")
    ctx.force_step(qd, pd, pl)
    # partial row range at build time (ghost particles are candidates only)
    qd, pd = s.device_arrays(torch, "aos4")
    own = s.pn // 3 + 1
    plo = ctx.makepair(qd, tiles=True, rows=(0, own))
    ctx.force_loop(qd, pd, plo, loop=s.steps, variant="celltile", rows=(0, own))
    ph = pd.cpu().numpy()[:, :3]
    assert np.abs(ph[:own] - s.p[:own]).max() / s.scale < TOL_FP64 and np.all(ph[own:] == 0)
    with pytest.raises(LJError):                          # another row range: no mirror for it
        ctx.force_step(qd, pd, plo, variant="celltile", rows=(0, own - 1))
    ctx.force_step(qd, pd, plo, rows=(0, own - 1))       # AUTO: per-row kernel
    # half lists and float4 positions have no mirror; the flag is ignored, the build still works
    plh = ctx.makepair(qd, half=True, tiles=True)
    assert plh.number_of_pairs == len(s.half[2])


def test_celltile_large_system_matches_per_row_kernel(ctx, torch):
    """N = 108k (rho = 1.0, L = 48): thousands of tiles, every SM walks several column segments."""
    from lj_gpu_b200 import init_fcc
    q = init_fcc(1.0, 48.0)
    pn = len(q)
    q4 = np.zeros((pn, 4)); q4[:, :3] = q
    qd = torch.from_numpy(q4).cuda()
    pl = ctx.makepair(qd, tiles=True)
    p1 = torch.zeros_like(qd); p2 = torch.zeros_like(qd)
    ctx.force_loop(qd, p1, pl, loop=3, variant="celltile")
    ctx.force_loop(qd, p2, pl, loop=3, variant="subwarp", group=8)
    assert torch.equal(p1, p2)
    pm = torch.zeros_like(qd)                             # mixed precision on the same mirror
    ctx.force_loop(qd, pm, pl, loop=3, variant="celltile", precision="mixed")
    assert 0 < ((pm - p1).abs().max() / p1.abs().max()).item() < TOL_MIXED
    plain = ctx.makepair(qd)                              # cluster-engine fill: same list as a set
    assert plain.number_of_pairs == pl.number_of_pairs
    assert torch.equal(plain.number_of_partners, pl.number_of_partners)
    rows = torch.repeat_interleave(torch.arange(pn, device="cuda"), plain.number_of_partners[:pn].long())
    ka = torch.sort(rows * pn + plain.sorted_list[:plain.number_of_pairs].long()).values
    kb = torch.sort(rows * pn + pl.sorted_list[:pl.number_of_pairs].long()).values
    assert torch.equal(ka, kb)


# ------------------------------------------------------------------------ SURVEY 8 leftovers (round 2)
def test_newton3_on_the_half_ell_table(ctx, torch, sysA, oracle):
    """force_kernel_memopt2_with_aar / memopt3_with_aar (cuda/kernel.cuh:344-423): thread per i on the
    HALF column-major ELL table, reaction on j by atomics.  Checker: the oracle's force_sorted on the
    same half list (= cpu_ref/force_soa.cpp:163-195)."""
    s = sysA
    for layout in ("aos4", "aos3", "soa"):
        qd, pd = s.device_arrays(torch, layout)
        pn = s.pn if layout == "soa" else None
        pl = ctx.makepair(qd, half=True, layout=layout, pn=pn, sort_rows=True)
        tl = ctx.make_transposed_pairlist(pl)
        tl_o, max_np = oracle.transpose_list(s.half[2], s.half[0], s.half[1])
        assert pl.max_partners == max_np
        assert np.array_equal(tl.cpu().numpy()[:max_np * s.pn], tl_o[:max_np * s.pn])
        ctx.force_loop(qd, pd, pl, loop=100, ell=True, layout=layout, pn=pn)   # half list -> Newton-3
        assert s.err(pd, layout) < TOL_FP64
        if layout == "aos4":
            assert torch.all(pd[:, 3] == 77.5)


def test_row_major_padded_ell(ctx, torch, sysA, sysS):
    """make_sorted_list2d() (cuda/force_cuda.cu:242-253) with a width that holds the longest row: the
    table equals the CSR rows zero padded, the gather on it reproduces the oracle, and the reference's
    own width (NUM_NEIGH = 60 < max_partners = 78 at rho = 0.5) is refused instead of overlapping."""
    from lj_gpu_b200 import LJError, _capi
    s = sysA
    qd, pd = s.device_arrays(torch, "aos4")
    pl = ctx.makepair(qd, sort_rows=True)
    assert pl.max_partners == 78
    with pytest.raises(LJError) as e:
        ctx.make_sorted_list2d(pl, width=60)
    assert e.value.status == _capi.LJ_ERR_CAPACITY
    t2 = ctx.make_sorted_list2d(pl).cpu().numpy().reshape(s.pn, pl.max_partners)
    nop, ptr, lst = s.full
    for i in (0, 1, s.pn // 2, s.pn - 1):
        assert np.array_equal(t2[i, :nop[i]], lst[ptr[i]:ptr[i] + nop[i]]) and np.all(t2[i, nop[i]:] == 0)
    assert int((t2 != 0).sum()) <= len(lst)
    for group in (1, 8, 32):
        pd.zero_()
        ctx.force_loop(qd, pd, pl, loop=100, ell_rows=True, group=group)
        assert s.err(pd, "aos4") < TOL_FP64, group
    # wider than needed, other layouts
    for layout in ("aos3", "soa"):
        q2, p2 = sysS.device_arrays(torch, layout)
        pn = sysS.pn if layout == "soa" else None
        pl2 = ctx.makepair(q2, layout=layout, pn=pn)
        ctx.make_sorted_list2d(pl2, width=pl2.max_partners + 5)
        ctx.force_loop(q2, p2, pl2, loop=sysS.steps, ell_rows=True, layout=layout, pn=pn)
        assert sysS.err(p2, layout) < TOL_FP64


@pytest.mark.parametrize("which", ["A", "B"])
def test_float3_layout_mixed(ctx, torch, sysA, sysB, oracle, which):
    """The reference's q_f3 / p_f3 buffers (cuda/force_cuda.cu:24): packed float3 in and out.
    Checker = the FP64 oracle on the SAME float-valued positions (as for float4)."""
    s = sysA if which == "A" else sysB
    qf = np.ascontiguousarray(s.q.astype(np.float32))
    q64 = np.ascontiguousarray(qf.astype(np.float64))
    nop_o, ptr_o, lst_o = oracle.makepair(q64, full=True)
    p_o = np.zeros_like(q64)
    oracle.force_gather(q64, p_o, nop_o, ptr_o, lst_o, steps=100, static_q=True)
    qd = torch.from_numpy(qf).cuda()
    assert qd.shape[1] == 3 and qd.dtype == torch.float32
    pd = torch.zeros_like(qd)
    pl = ctx.makepair(qd)                                          # list build straight from float3
    nop, ptr, lst = list_to_host(pl)
    assert np.array_equal(nop, nop_o) and np.array_equal(s.lo.sort_rows(nop, ptr, lst), lst_o)
    ctx.force_loop(qd, pd, pl, loop=100, precision="mixed")
    assert np.abs(pd.cpu().numpy() - p_o).max() / np.abs(p_o).max() < TOL_MIXED
    from lj_gpu_b200 import LJError
    with pytest.raises(LJError):
        ctx.force_step(qd, pd, pl, precision="fp64")              # float3 is a mixed-precision layout


def test_force_step_parts_interior_plus_boundary_equal_one_step(ctx, torch, sysS):
    """lj_force_step_part: INTERIOR tiles read only positions inside the list's row range (garbage in
    the other positions does not reach them), BOUNDARY refreshes those positions and runs the rest;
    together they are exactly one lj_force_step."""
    from lj_gpu_b200 import LJError
    s = sysS
    qd, pd = s.device_arrays(torch, "aos4")
    own = (s.pn // 2) // 4 * 4
    pl = ctx.makepair(qd, tiles=True, rows=(0, own))
    ref = torch.zeros_like(pd)
    ctx.force_loop(qd, ref, pl, loop=3, variant="celltile", rows=(0, own))
    good = qd.clone()
    pd.zero_()
    touched_by_interior = None
    for _ in range(3):
        qd[own:, :3] = 1.0e5                                  # "ghosts not here yet"
        before = pd.clone()
        ctx.force_step(qd, pd, pl, variant="celltile", rows=(0, own), part="interior")
        touched_by_interior = (pd != before).any(dim=1)
        qd.copy_(good)                                         # "the halo has arrived"
        ctx.force_step(qd, pd, pl, variant="celltile", rows=(0, own), part="boundary")
    assert torch.equal(pd, ref)
    n_int = int(touched_by_interior.sum().item())
    assert 0 < n_int < own                                     # both parts had rows
    assert np.abs(pd.cpu().numpy()[:own, :3] - s.p[:own] * (3 / s.steps)).max() / s.scale < TOL_FP64
    # mixed precision goes through the same two parts
    pm, pr = torch.zeros_like(pd), torch.zeros_like(pd)
    ctx.force_step(qd, pr, pl, variant="celltile", rows=(0, own), precision="mixed")
    ctx.force_step(qd, pm, pl, variant="celltile", rows=(0, own), precision="mixed", part="interior")
    ctx.force_step(qd, pm, pl, variant="celltile", rows=(0, own), precision="mixed", part="boundary")
    assert torch.equal(pm, pr)
    # a whole-system list has no boundary tiles; no mirror -> an error, not a silent fallback
    full = ctx.makepair(qd, tiles=True)
    pa, pb = torch.zeros_like(pd), torch.zeros_like(pd)
    ctx.force_step(qd, pa, full, variant="celltile")
    ctx.force_step(qd, pb, full, variant="celltile", part="interior")
    ctx.force_step(qd, pb, full, variant="celltile", part="boundary")
    assert torch.equal(pa, pb)
    plain = ctx.makepair(qd)
    with pytest.raises(LJError):
        ctx.force_step(qd, pd, plain, part="interior")


# ------------------------------------------------------------------------ lj_list_mirror (round 2)
def test_list_mirror_of_a_caller_supplied_shuffled_list(ctx, torch, oracle):
    """The reference always hands the kernel a host-built, random_shfl()-ed or cache-loaded list
    (cuda/force_cuda.cu:392-397).  lj_list_mirror builds the cell-tile mirror for such a list: the
    cell-tile kernel then consumes the caller's row order (bit-identical to the per-row kernel with 8
    lanes per row on the same arrays) and matches the oracle."""
    from lj_gpu_b200 import LJError, PairList, init_fcc
    q = init_fcc(1.0, 48.0)
    pn = len(q)
    nop_o, ptr_o, lst_o = oracle.makepair(q, full=True)
    oracle.shuffle_rows(lst_o, nop_o, ptr_o, seed=10)               # the reference's random_shfl (std::shuffle)
    p_o = np.zeros_like(q)
    oracle.force_gather(q, p_o, nop_o, ptr_o, lst_o, steps=7, static_q=True)
    q4 = np.zeros((pn, 4)); q4[:, :3] = q
    qd = torch.from_numpy(q4).cuda()
    pl = PairList(torch.from_numpy(nop_o).cuda(), torch.from_numpy(ptr_o.astype(np.int32)).cuda(),
                  torch.from_numpy(lst_o).cuda(), len(lst_o), int(nop_o.max()))
    # no mirror yet: explicit request fails, AUTO runs the per-row kernel
    pd = torch.zeros_like(qd)
    with pytest.raises(LJError):
        ctx.force_step(qd, pd, pl, variant="celltile")
    assert ctx.list_mirror(qd, pl) == 0                              # every row fits its tile's region
    assert pl.token != 0
    launches = ctx.launches
    ctx.force_loop(qd, pd, pl, loop=7, variant="celltile")
    assert ctx.launches - launches == 14                             # permute + cell-tile kernel per step
    pr = torch.zeros_like(qd)
    ctx.force_loop(qd, pr, pl, loop=7, variant="subwarp", group=8)
    assert torch.equal(pd, pr)
    assert np.abs(pd.cpu().numpy()[:, :3] - p_o).max() / np.abs(p_o).max() < TOL_FP64
    pm = torch.zeros_like(qd)
    ctx.force_loop(qd, pm, pl, loop=7, variant="celltile", precision="mixed")
    assert 1e-14 < np.abs(pm.cpu().numpy()[:, :3] - p_o).max() / np.abs(p_o).max() < TOL_MIXED
    # the stale-list hazard: ANOTHER list written into the same three buffers.  A caller that does not
    # vouch for the mirror (token 0) gets the per-row kernel and the right answer; the old token would not.
    half_rows = nop_o.copy(); half_rows[::2] = 0                      # drop every second row's partners
    pl.number_of_partners.copy_(torch.from_numpy(half_rows).cuda())
    p2 = np.zeros_like(q)
    oracle.force_gather(q, p2, half_rows, ptr_o, lst_o, steps=1)
    fresh = PairList(pl.number_of_partners, pl.pointer, pl.sorted_list, len(lst_o), int(nop_o.max()))
    pd.zero_()
    ctx.force_step(qd, pd, fresh)                                     # AUTO, token 0
    assert np.abs(pd.cpu().numpy()[:, :3] - p2).max() / np.abs(p2).max() < TOL_FP64
    with pytest.raises(LJError):
        ctx.force_step(qd, pd, fresh, variant="celltile")             # explicit request needs the token too


def test_list_mirror_rows_outside_their_region_take_the_per_row_kernel(ctx, torch, oracle):
    """A list built with a LONGER search length than the mirror's cell grid assumes: some rows have a
    partner outside the 5 x 5 pencil region of their tile.  Those rows are left out of the mirror and
    completed by the per-row kernel; the result is the oracle's either way."""
    from lj_gpu_b200 import PairList, init_fcc
    q = init_fcc(0.7, 20.0)
    pn = len(q)
    nop_o, ptr_o, lst_o = oracle.makepair(q, search_len=4.1, full=True)
    p_o = np.zeros_like(q)
    oracle.force_gather(q, p_o, nop_o, ptr_o, lst_o, steps=3, static_q=True)
    q3 = torch.from_numpy(q).cuda()
    pl = PairList(torch.from_numpy(nop_o).cuda(), torch.from_numpy(ptr_o.astype(np.int32)).cuda(),
                  torch.from_numpy(lst_o).cuda(), len(lst_o), int(nop_o.max()))
    outside = ctx.list_mirror(q3, pl, search_len=3.3, layout="aos3")
    assert 0 < outside < pn
    pd = torch.zeros_like(q3)
    launches = ctx.launches
    ctx.force_loop(q3, pd, pl, loop=3, layout="aos3", variant="celltile")
    assert ctx.launches - launches == 9                               # permute + cell-tile + per-row completion
    assert np.abs(pd.cpu().numpy() - p_o).max() / np.abs(p_o).max() < TOL_FP64
    # with the search length the list was built with, every row fits
    assert ctx.list_mirror(q3, pl, search_len=4.1, layout="aos3") == 0
    pd.zero_()
    ctx.force_loop(q3, pd, pl, loop=3, layout="aos3", variant="celltile")
    assert np.abs(pd.cpu().numpy() - p_o).max() / np.abs(p_o).max() < TOL_FP64


def test_cpp_driver_cached_list_runs_the_cell_tile_kernel(tmp_path, oracle):
    """force_b200 --cache at N >= 3e5: the first run builds on the GPU and writes the reference's text cache,
    the second loads it (the reference's own flow, cuda/force_cuda.cu:203-227), mirrors the LOADED list
    (lj_list_mirror inside lj_measure) and prints the oracle's momenta."""
    exe = os.path.join(ROOT, "lj_gpu_b200", "driver", "force_b200")
    if not os.path.exists(exe):
        pytest.skip("driver not built")
    args = [exe, "--density", "1.0", "--L", "70.0", "--variant", "auto", "--layout", "aos4", "--steps", "10", "--cache",
            "--print"]
    r1 = subprocess.run(args, capture_output=True, text=True, timeout=600, cwd=tmp_path)
    assert r1.returncode == 0, r1.stderr
    assert "Now make pairlist .cache_pair_all.dat." in r1.stderr
    r2 = subprocess.run(args, capture_output=True, text=True, timeout=600, cwd=tmp_path)
    assert r2.returncode == 0, r2.stderr
    assert ".cache_pair_all.dat is successfully loaded." in r2.stderr
    assert "cell-tile mirror of the loaded list: yes" in r2.stderr
    q = oracle.init_fcc(1.0, 70.0)
    assert len(q) >= 300000
    nop, ptr, lst = oracle.makepair(q, full=True)
    p = np.zeros_like(q)
    oracle.force_gather(q, p, nop, ptr, lst, steps=10, static_q=True)
    from lj_gpu_b200 import print_results
    assert r2.stdout == print_results(p) and r1.stdout == r2.stdout
