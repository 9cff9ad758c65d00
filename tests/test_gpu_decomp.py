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
