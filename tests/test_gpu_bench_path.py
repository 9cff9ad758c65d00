"""GPU parity of the path bench.py times: lj_build_list(LJ_LIST_TILES) -> k_tile_fill's list and
mirror -> lj_celltile_force (FP64 and mixed precision), against the CPU oracle.

  * BASELINE config 3 (C: N = 1,000,188, rho = 1.0): list bit-exact (all 1.37e8 entries), FP64 after
    the reference's LOOP = 100 steps <= 1e-12, mixed precision <= 1e-5, CUDA-graph replay and
    lj_measure() (the e2e call of bench.py) reproduce the same numbers;
  * configs A and B (the reference's own sizes) forced through the cell-tile kernel print
    ref_data/density0.5.dat / density1.dat byte for byte;
  * BASELINE config 4 (D: N = 16,078,716, 2.2e9 list entries, int64 pointer[]): >= 256 sampled rows
    bit-exact against the oracle's brute-force rows (each sampled particle against all 16 M), the
    momenta of those rows after FP64 / mixed / Newton-3 / cell-tile steps against the oracle's
    arithmetic on them.

Tolerances (BASELINE.json north_star): lists bit-exact after per-row sorting; momenta norm-wise
relative |dp|_max / |p|_max <= 1e-12 (FP64), <= 1e-5 (mixed), after the reference's step count.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

TOL_FP64 = 1e-12
TOL_MIXED = 1e-5


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


def _aos4(torch, q):
    q4 = np.zeros((len(q), 4))
    q4[:, :3] = q
    return torch.from_numpy(q4).cuda()


def _sorted_keys(torch, pl, pn):
    """(row, j) keys of the device list, sorted on the device: row * pn + j ascending."""
    rows = torch.repeat_interleave(torch.arange(pn, device="cuda"), pl.number_of_partners[:pn].long())
    return torch.sort(rows * pn + pl.sorted_list[:pl.number_of_pairs].long()).values


# ------------------------------------------------------------------------ config C, the bench workload
@pytest.fixture(scope="module")
def sysC(oracle, torch):
    """Config C with the oracle's list and its momenta after LOOP = 100 steps on static positions."""
    from lj_gpu_b200 import init_fcc
    q = init_fcc(1.0, 100.1)
    assert len(q) == 1000188
    nop, ptr, lst = oracle.makepair(q, full=True)          # cell-list restatement, pinned at A and B
    p = np.zeros_like(q)
    oracle.force_gather(q, p, nop, ptr, lst, steps=100, static_q=True)
    # the static-q shortcut is the same oracle: one explicit step-by-step run for 3 steps agrees bit for bit
    p3a, p3b = np.zeros_like(q), np.zeros_like(q)
    oracle.force_gather(q, p3a, nop, ptr, lst, steps=3)
    oracle.force_gather(q, p3b, nop, ptr, lst, steps=3, static_q=True)
    assert np.array_equal(p3a, p3b)
    return dict(q=q, nop=nop, ptr=ptr, lst=lst, p=p, scale=np.abs(p).max())


def test_config_C_tile_fill_list_is_the_oracle_list(ctx, torch, sysC):
    """The list k_tile_fill writes (the bench's list build, LJ_LIST_TILES) == oracle.makepair, all
    136,733,458 entries, for both tile widths the bench uses."""
    s = sysC
    pn = len(s["q"])
    qd = _aos4(torch, s["q"])
    want = torch.from_numpy(np.repeat(np.arange(pn, dtype=np.int64), s["nop"]) * pn + s["lst"]).cuda()
    for tiles in (True, "wide"):
        pl = ctx.makepair(qd, tiles=tiles)
        assert pl.number_of_pairs == len(s["lst"]) == 136733458
        assert np.array_equal(pl.number_of_partners.cpu().numpy(), s["nop"])
        assert np.array_equal(pl.pointer.cpu().numpy().astype(np.int64), s["ptr"])
        assert torch.equal(_sorted_keys(torch, pl, pn), want)      # oracle rows are ascending in j
        assert pl.max_partners == int(s["nop"].max())
        del pl


@pytest.mark.parametrize("use_graph", [False, True])
def test_config_C_celltile_fp64_100_steps(ctx, torch, sysC, use_graph):
    s = sysC
    qd = _aos4(torch, s["q"])
    pd = torch.zeros_like(qd)
    pd[:, 3] = 77.5
    pl = ctx.makepair(qd, tiles=True)
    launches = ctx.launches
    ctx.force_loop(qd, pd, pl, loop=100, variant="celltile", use_graph=use_graph)
    ctx.sync()
    if not use_graph:
        assert ctx.launches - launches == 200                      # position permute + force kernel per step
    ph = pd.cpu().numpy()
    err = np.abs(ph[:, :3] - s["p"]).max() / s["scale"]
    assert err < TOL_FP64, err
    assert np.all(ph[:, 3] == 77.5)                                # .w of p untouched
    # AUTO on this list IS the cell-tile kernel (what bench.py launches), and it is bit-identical to
    # the per-row kernel with 8 lanes per row on the same list order
    pa, pr = torch.zeros_like(qd), torch.zeros_like(qd)
    ctx.force_loop(qd, pa, pl, loop=5, variant="auto")
    ctx.force_loop(qd, pr, pl, loop=5, variant="subwarp", group=8)
    pc = torch.zeros_like(qd)
    ctx.force_loop(qd, pc, pl, loop=5, variant="celltile")
    assert torch.equal(pa, pc) and torch.equal(pa, pr)


def test_config_C_celltile_mixed_100_steps(ctx, torch, sysC):
    s = sysC
    qd = _aos4(torch, s["q"])
    for tiles in ("wide", True):
        pl = ctx.makepair(qd, tiles=tiles)
        pd = torch.zeros_like(qd)
        ctx.force_loop(qd, pd, pl, loop=100, variant="celltile", precision="mixed")
        err = np.abs(pd.cpu().numpy()[:, :3] - s["p"]).max() / s["scale"]
        assert 1e-14 < err < TOL_MIXED, err                        # FP32 pair arithmetic really ran
        del pl


def test_config_C_measure_is_the_bench_e2e_call(ctx, torch, sysC):
    """lj_measure() on host buffers with the bench's cadence (rebuild every 20 steps, AUTO ->
    cell-tile, CUDA graph): 100 steps, 5 list builds, the oracle's momenta."""
    s = sysC
    pn = len(s["q"])
    qh = np.zeros((pn, 4)); qh[:, :3] = s["q"]
    ph = np.zeros((pn, 4))
    m = ctx.measure(qh, ph, layout="aos4", loop=100, rebuild_every=20, variant="auto", use_graph=True)
    assert m.number_of_pairs == len(s["lst"]) and m.list_builds == 5
    assert np.abs(ph[:, :3] - s["p"]).max() / s["scale"] < TOL_FP64


def test_config_C_kernel_timing_counts_the_dominant_kernel_alone(ctx, torch, sysC):
    """lj_kernel_timing (the roofline leg of bench.py): one event pair per launch of lj_celltile_force, its
    summed duration below the time of the whole loop (which also permutes the positions every step), nothing
    recorded for per-row launches, under a stream capture or once switched off; results unchanged."""
    s = sysC
    qd = _aos4(torch, s["q"])
    pl = ctx.makepair(qd, tiles=True)
    pd = torch.zeros_like(qd)
    ctx.force_loop(qd, pd, pl, loop=3, variant="celltile")            # warm
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.current_stream()
    pd.zero_()
    ctx.kernel_timing(True)
    e0.record(st)
    ctx.force_loop(qd, pd, pl, loop=100, variant="celltile")          # more launches than event pairs in the ring
    e1.record(st)
    torch.cuda.synchronize()
    ms, n = ctx.kernel_timing_read()
    assert n == 100 and 0.0 < ms < e0.elapsed_time(e1)
    assert ms / n > 0.05                                              # a 1M-atom force step is not microseconds
    ctx.force_loop(qd, pd.clone(), pl, loop=2, variant="subwarp", group=8)      # not the dominant kernel
    ctx.force_loop(qd, pd.clone(), pl, loop=4, variant="celltile", use_graph=True)  # captured launches are skipped
    torch.cuda.synchronize()
    assert ctx.kernel_timing_read()[1] == 100 + 1                     # (+ the graph path's warm launch outside the capture)
    ctx.kernel_timing(False)
    ctx.force_loop(qd, pd.clone(), pl, loop=2, variant="celltile")
    assert ctx.kernel_timing_read()[1] == 101
    assert np.abs(pd.cpu().numpy()[:, :3] - s["p"]).max() / s["scale"] < TOL_FP64


# ------------------------------------------------------------------------ configs A and B through the cell-tile kernel
@pytest.mark.parametrize("rho,gold", [(0.5, "density0.5.dat"), (1.0, "density1.dat")])
@pytest.mark.parametrize("layout", ["aos4", "aos3", "soa"])
def test_reference_configs_through_celltile_print_the_goldens(ctx, torch, oracle, golden, rho, gold, layout):
    from lj_gpu_b200 import init_fcc, print_results
    q = init_fcc(rho, 50.0)
    pn = len(q)
    if layout == "aos3":
        qh = q.copy()
    elif layout == "aos4":
        qh = np.zeros((pn, 4)); qh[:, :3] = q
    else:
        qh = np.ascontiguousarray(q.T)
    qd = torch.from_numpy(qh).cuda()
    pd = torch.zeros_like(qd)
    npn = pn if layout == "soa" else None
    pl = ctx.makepair(qd, layout=layout, pn=npn, tiles=True)
    ctx.force_loop(qd, pd, pl, loop=100, layout=layout, pn=npn, variant="celltile")
    a = pd.cpu().numpy()
    ph = a[:, :3] if layout != "soa" else a[:, :pn].T
    with open(os.path.join(GOLDEN, gold)) as f:
        assert print_results(ph) == f.read()                       # byte-identical to ref_data/
    g = golden(rho)   # full-precision samples dumped from the REAL reference (force_sorted x 100)
    assert np.abs(ph[g["sample_idx"]] - g["p_sorted_sample"]).max() / float(g["p_sorted_absmax"]) < TOL_FP64
    # mixed precision on the same mirror: the printed goldens only resolve 1e-10 absolute, so compare norm-wise
    pm = torch.zeros_like(qd)
    ctx.force_loop(qd, pm, pl, loop=100, layout=layout, pn=npn, variant="celltile", precision="mixed")
    b = pm.cpu().numpy()
    pmh = b[:, :3] if layout != "soa" else b[:, :pn].T
    assert np.abs(pmh[g["sample_idx"]] - g["p_sorted_sample"]).max() / float(g["p_sorted_absmax"]) < TOL_MIXED


# ------------------------------------------------------------------------ config D, 16 M atoms
def test_config_D_sampled_rows_bit_exact_and_forces(ctx, torch, oracle):
    """BASELINE config 4: N = 16,078,716 (159 cells/side), full list 2.2e9 entries > 2^31 -> int64
    pointer[].  256 sampled rows (+ the first and last particles) against the oracle's brute-force
    rows over all 16 M particles: bit-exact after per-row sorting; the momenta of those rows after
    3 steps for the FP64 per-row kernel, the cell-tile kernel (FP64 and mixed), the per-row mixed
    kernel, and Newton-3 on the half list (whose sampled rows are checked bit-exact too)."""
    from lj_gpu_b200 import init_fcc
    from lj_gpu_b200.decomp import lattice_spacing
    q = init_fcc(1.0, (159 + 0.05) * lattice_spacing(1.0))
    pn = len(q)
    assert pn == 16078716
    rng = np.random.RandomState(4)
    rows = np.unique(np.concatenate([rng.choice(pn, 256, replace=False), [0, 1, 2, pn - 2, pn - 1]]))
    assert len(rows) >= 256
    nop_o, ptr_o, lst_o = oracle.rows_brute(q, rows, full=True)
    steps = 3
    p_o = oracle.force_rows(q, rows, nop_o, ptr_o, lst_o, steps=steps)
    qd = _aos4(torch, q)
    rows_d = torch.from_numpy(rows).cuda()

    def check_rows(pl, nop_w, ptr_w, lst_w):
        nop = pl.number_of_partners[rows_d].cpu().numpy()
        ptr = pl.pointer[rows_d].cpu().numpy().astype(np.int64)
        assert np.array_equal(nop, nop_w)
        for k in range(len(rows)):
            got = np.sort(pl.sorted_list[ptr[k]:ptr[k] + nop[k]].cpu().numpy())
            assert np.array_equal(got, lst_w[ptr_w[k]:ptr_w[k + 1]]), rows[k]

    pl = ctx.makepair(qd, pointer64=True, tiles=True)
    assert pl.number_of_pairs > 2 ** 31 and pl.pointer.dtype == torch.int64
    check_rows(pl, nop_o, ptr_o, lst_o)
    # the sampled momenta are compared on the scale of the whole system (norm-wise relative)
    results = {}
    for name, kw, tol in (("fp64 per-row g8", dict(variant="subwarp", group=8), TOL_FP64),
                          ("fp64 warp/i", dict(variant="warp", group=32), TOL_FP64),
                          ("fp64 cell-tile", dict(variant="celltile"), TOL_FP64),
                          ("mixed cell-tile", dict(variant="celltile", precision="mixed"), TOL_MIXED),
                          ("mixed per-row g4", dict(variant="subwarp", group=4, precision="mixed"), TOL_MIXED)):
        pd = torch.zeros_like(qd)
        ctx.force_loop(qd, pd, pl, loop=steps, **kw)
        scale = float(pd[:, :3].abs().max().item())
        got = pd[rows_d, :3].cpu().numpy()
        err = np.abs(got - p_o).max() / scale
        assert err < tol, (name, err)
        if "mixed" not in name:   # Newton's third law on the full list: directed contributions cancel pairwise
            assert float(pd[:, :3].sum(0).abs().max().item()) < 1e-9 * scale * np.sqrt(pn)
        results[name] = pd if name == "fp64 per-row g8" else None
    p_ref = results["fp64 per-row g8"]
    pc = torch.zeros_like(qd)
    ctx.force_loop(qd, pc, pl, loop=steps, variant="celltile")
    assert torch.equal(pc, p_ref)                                  # all 16 M rows, bit for bit
    del pl, pc
    torch.cuda.empty_cache()
    # half list + Newton-3 scatter
    nop_h, ptr_h, lst_h = oracle.rows_brute(q, rows, full=False)
    half = ctx.makepair(qd, half=True, pointer64=True)
    check_rows(half, nop_h, ptr_h, lst_h)
    pn3 = torch.zeros_like(qd)
    ctx.force_loop(qd, pn3, half, loop=steps, variant="n3", group=8)
    scale = float(p_ref[:, :3].abs().max().item())
    assert np.abs(pn3[rows_d, :3].cpu().numpy() - p_o).max() / scale < TOL_FP64
    assert float((pn3 - p_ref)[:, :3].abs().max().item()) / scale < TOL_FP64   # every row vs the gather kernel
