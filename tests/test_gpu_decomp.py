"""Decomposed (z-slab) path on real GPUs: needs >= 2 devices, skipped otherwise (run with
`gpurun --gpus 2`; the log of the last run is committed under profiles/).  Two ranks, NCCL and
CUDA-IPC halo transports, per-row and cell-tile kernels, halo-then-force and overlapped schedules.

Checker = the CPU oracle on the UNDECOMPOSED system (tests/ may use oracle/): the owned momenta of
all ranks, concatenated, must equal the oracle's within 1e-12; with moving particles (kick + drift
+ list rebuilds, the ghost exchange ordered by the device-side flag handshake) positions and
momenta must equal the oracle's kick + a numpy drift."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DENSITY, L, STEPS = 1.0, 40.0, 25
MD_STEPS, MD_DT, MD_REBUILD = 40, 0.004, 8


def _worker(rank, world, port, mode, out_dir):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from lj_gpu_b200 import decomp
    parts = mode.split("-")
    tiles = "tiles" in parts          # cell-tile mirror per slab
    overlap = "overlap" in parts      # interior while the halo flies, boundary after its event
    md = "md" in parts
    system = decomp.DecomposedSystem(DENSITY, L, halo_mode=parts[0], tiles=tiles)
    # ghosts start as garbage: they must come from the exchange
    system.q[system.slab.n_own:] = 1e6
    torch.cuda.synchronize(); dist.barrier()
    fkw = dict(variant="celltile") if tiles else dict(variant="tile", group=8)
    if md:
        system.run_md(MD_STEPS, MD_DT, MD_REBUILD, overlap=overlap, **fkw)
    else:
        system.run(STEPS, rebuild_every=10, first_step=1, overlap=overlap, **fkw)
    torch.cuda.synchronize()
    p = system.gather("p")
    q = system.gather("q")
    if rank == 0:
        np.save(os.path.join(out_dir, "p_%s.npy" % mode), p)
        np.save(os.path.join(out_dir, "q_%s.npy" % mode), q)
    dist.barrier()
    dist.destroy_process_group()


def _spawn(mode, tmp_path):
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, mode, str(tmp_path)), nprocs=2, join=True)
    return np.load(tmp_path / ("p_%s.npy" % mode)), np.load(tmp_path / ("q_%s.npy" % mode))


@pytest.fixture(scope="module")
def static_oracle(oracle):
    q = oracle.init_fcc(DENSITY, L)
    nop, ptr, lst = oracle.makepair(q, full=True)
    p = np.zeros_like(q)
    oracle.force_gather(q, p, nop, ptr, lst, steps=STEPS, static_q=True)
    return q, p


@pytest.mark.parametrize("mode", ["nccl", "nccl-overlap", "p2p", "p2p-overlap", "p2p-tiles", "p2p-tiles-overlap",
                                  "nccl-tiles-overlap"])
def test_two_gpu_decomposition_matches_the_oracle(mode, tmp_path, static_oracle):
    got, q_got = _spawn(mode, tmp_path)
    q, ref = static_oracle
    assert got.shape == ref.shape
    assert np.array_equal(q_got, q)                                 # static run: positions untouched
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-12


@pytest.fixture(scope="module")
def md_oracle(oracle):
    """kick (oracle gather on the current list) + drift (numpy), list rebuilt at the fixed cadence."""
    q = oracle.init_fcc(DENSITY, L)
    p = np.zeros_like(q)
    nop, ptr, lst = oracle.makepair(q, full=True)
    for k in range(MD_STEPS):
        if k > 0 and k % MD_REBUILD == 0:
            nop, ptr, lst = oracle.makepair(q, full=True)
        oracle.force_gather(q, p, nop, ptr, lst, steps=1, dt=MD_DT)
        q += p * MD_DT
    return q, p


@pytest.mark.parametrize("mode", ["p2p-tiles-overlap-md", "p2p-overlap-md", "nccl-tiles-overlap-md"])
def test_two_gpu_moving_particles_match_the_oracle(mode, tmp_path, md_oracle, oracle):
    """Drift + rebuild on slabs: rank r reads rank r+-1's q while that rank is about to overwrite it;
    the flag handshake (lj_flag_set / lj_halo_pull_sync / lj_flag_wait) orders the two."""
    p_got, q_got = _spawn(mode, tmp_path)
    q, p = md_oracle
    q0 = oracle.init_fcc(DENSITY, L)
    assert np.abs(q - q0).max() > 1e-3                               # the particles really moved
    assert np.abs(q_got - q).max() < 1e-11
    assert np.abs(p_got - p).max() / np.abs(p).max() < 1e-11        # 40 dependent steps: rounding compounds


# ------------------------------------------------------------------------ the same behind the C ABI, ONE process
def _decomp_capi(q, ngpus, steps, rebuild_every, md=False, dt=0.001, prec=0, devices=None, overlap=1):
    """lj_decomp_* through ctypes: one process, `ngpus` contexts on `ngpus` devices (or on `devices`)."""
    import ctypes as C

    from lj_gpu_b200 import _capi
    lib = _capi.load()
    pn = len(q)
    slab = (C.c_int64 * (ngpus + 1))()
    halo = C.c_int64(0)
    assert lib.lj_decomp_plan_fcc(DENSITY, L, ngpus, 3.3, slab, C.byref(halo)) == 0
    a = _capi.LjDecompArgs()
    qc = np.ascontiguousarray(q, np.float64)
    a.ngpus, a.q_xyz_host, a.pn = ngpus, qc.ctypes.data, pn
    a.slab_begin, a.halo_rows = C.cast(slab, C.c_void_p), halo.value
    a.search_len, a.cutoff, a.dt, a.precision = 3.3, 3.0, dt, prec
    if devices is not None:
        dev = (C.c_int32 * ngpus)(*devices)
        a.devices = C.cast(dev, C.c_void_p)
    d = C.c_void_p()
    rc = lib.lj_decomp_create(C.byref(d), C.byref(a))
    assert rc == 0, lib.lj_decomp_last_error(d)
    rc = (lib.lj_decomp_md if md else lib.lj_decomp_step)(d, steps, rebuild_every, overlap)
    assert rc == 0, lib.lj_decomp_last_error(d)
    p = np.zeros((pn, 3)); qo = np.zeros((pn, 3))
    assert lib.lj_decomp_gather(d, p.ctypes.data, qo.ctypes.data) == 0, lib.lj_decomp_last_error(d)
    pairs, launches = lib.lj_decomp_pairs(d), lib.lj_decomp_launch_count(d)
    assert lib.lj_decomp_destroy(d) == 0
    return p, qo, pairs, launches


def test_capi_decomposition_two_contexts_two_devices_one_process(static_oracle, md_oracle, oracle):
    """lj_decomp_create/step/md/gather: two lj_ctx on two devices driven from ONE host thread (what a C++
    caller of the reference's driver would do) -- the per-device kernel attributes, the device guard at
    every entry point and the peer-memory flag handshake all have to hold.  Checker: the oracle."""
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    q, ref = static_oracle
    nop, _, lst = oracle.makepair(q, full=True)
    p, qo, pairs, launches = _decomp_capi(q, 2, STEPS, 10)
    assert pairs == len(lst) and launches > 4 * STEPS
    assert np.array_equal(qo, q)
    assert np.abs(p - ref).max() / np.abs(ref).max() < 1e-12
    pm, _, _, _ = _decomp_capi(q, 2, STEPS, 10, prec=1)                # mixed precision on the slabs
    assert 1e-14 < np.abs(pm - ref).max() / np.abs(ref).max() < 1e-5
    qm, pmd = md_oracle
    p, qo, _, _ = _decomp_capi(q, 2, MD_STEPS, MD_REBUILD, md=True, dt=MD_DT)
    assert np.abs(qo - qm).max() < 1e-11
    assert np.abs(p - pmd).max() / np.abs(pmd).max() < 1e-11
    # one slab is the plain single-GPU run through the same entry points
    p1, _, _, _ = _decomp_capi(q, 1, STEPS, 10)
    assert np.abs(p1 - ref).max() / np.abs(ref).max() < 1e-12


@pytest.mark.parametrize("nslabs", [2, 3])
def test_capi_decomposition_slabs_share_one_device(nslabs, static_oracle, md_oracle, oracle):
    """The whole decomposed path on a ONE-GPU box: `nslabs` contexts on device 0 (lj_decomp_args.devices), so
    slabs, ghost plan, device-side flag handshake, INTERIOR / BOUNDARY part launches, rebuilds on ghosted slabs
    and the drift ordering all run exactly as on `nslabs` devices -- only the peer pointers are local.
    Checker: the oracle on the undecomposed system."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    q, ref = static_oracle
    nop, _, lst = oracle.makepair(q, full=True)
    dev = [0] * nslabs
    for overlap in (1, 0):
        p, qo, pairs, launches = _decomp_capi(q, nslabs, STEPS, 10, devices=dev, overlap=overlap)
        assert pairs == len(lst) and launches > 2 * nslabs * STEPS
        assert np.array_equal(qo, q)
        assert np.abs(p - ref).max() / np.abs(ref).max() < 1e-12
    pm, _, _, _ = _decomp_capi(q, nslabs, STEPS, 10, prec=1, devices=dev)
    assert 1e-14 < np.abs(pm - ref).max() / np.abs(ref).max() < 1e-5
    qm, pmd = md_oracle
    p, qo, _, _ = _decomp_capi(q, nslabs, MD_STEPS, MD_REBUILD, md=True, dt=MD_DT, devices=dev)
    assert np.abs(qo - qm).max() < 1e-11
    assert np.abs(p - pmd).max() / np.abs(pmd).max() < 1e-11


def test_cpp_driver_gpus_2(static_oracle):
    from conftest import ROOT
    import subprocess
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    exe = os.path.join(ROOT, "lj_gpu_b200", "driver", "force_b200")
    if not os.path.exists(exe):
        pytest.skip("driver not built")
    r = subprocess.run([exe, "--gpus", "2", "--density", str(DENSITY), "--L", str(L), "--steps", str(STEPS),
                        "--rebuild-every", "10", "--print"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert "force_decomposed_2gpus" in r.stderr and "(without Host<->Device)" in r.stderr
    from lj_gpu_b200 import print_results
    assert r.stdout == print_results(static_oracle[1])


def test_cpp_driver_gpus_3_on_one_device(static_oracle):
    """force_b200 --gpus 3 --one-device: the C++ driver's decomposed run with its three slabs on device 0."""
    from conftest import ROOT
    import subprocess
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    exe = os.path.join(ROOT, "lj_gpu_b200", "driver", "force_b200")
    if not os.path.exists(exe):
        pytest.skip("driver not built")
    r = subprocess.run([exe, "--gpus", "3", "--one-device", "--density", str(DENSITY), "--L", str(L), "--steps", str(STEPS),
                        "--rebuild-every", "10", "--print"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert "force_decomposed_3gpus" in r.stderr and "slabs=3" in r.stderr
    from lj_gpu_b200 import print_results
    assert r.stdout == print_results(static_oracle[1])
