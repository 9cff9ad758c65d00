"""Decomposed (z-slab) path on real GPUs: needs >= 2 devices, skipped otherwise.  Two ranks, NCCL
and CUDA-IPC halo transports; owned momenta of all ranks must equal the single-GPU result."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DENSITY, L, STEPS = 1.0, 40.0, 25


def _worker(rank, world, port, mode, out_dir):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from lj_gpu_b200 import decomp
    tiles = mode.endswith("-tiles")   # cell-tile mirror per slab, halo-then-force schedule
    system = decomp.DecomposedSystem(DENSITY, L, halo_mode=mode.split("-")[0], tiles=tiles)
    # ghosts start as garbage: they must come from the exchange
    system.q[system.slab.n_own:] = 1e6
    torch.cuda.synchronize(); dist.barrier()
    if tiles:
        system.run(STEPS, rebuild_every=10, first_step=1, overlap=False, variant="celltile")
    else:
        system.run(STEPS, rebuild_every=10, first_step=1, variant="tile", group=8)
    torch.cuda.synchronize()
    p = system.gather_p()
    if rank == 0:
        np.save(os.path.join(out_dir, "p_%s.npy" % mode), p)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["nccl", "p2p", "p2p-tiles"])
def test_two_gpu_decomposition_matches_single_gpu(mode, tmp_path):
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, mode, str(tmp_path)), nprocs=2, join=True)
    from lj_gpu_b200 import LJContext, init_fcc
    ctx = LJContext(0)
    q = init_fcc(DENSITY, L)
    q4 = np.zeros((len(q), 4)); q4[:, :3] = q
    qd = torch.from_numpy(q4).cuda(); pd = torch.zeros_like(qd)
    pl = ctx.makepair(qd)
    ctx.force_loop(qd, pd, pl, loop=STEPS, variant="subwarp", group=8)
    ref = pd.cpu().numpy()[:, :3]
    got = np.load(tmp_path / ("p_%s.npy" % mode))
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-12
