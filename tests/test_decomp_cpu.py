"""Host logic of the z-slab decomposition (lj_gpu_b200/decomp.py) on CPU: world_size 2 and 3 over
gloo.  The exchange pattern is the one the GPU path runs over NCCL; the per-rank compute is done
here by the CPU oracle (test infrastructure) so that the test needs no GPU:

  every rank fills only its OWNED positions -> halo exchange (gloo) -> local list + force by the
  oracle on [owned | ghosts] -> owned momenta gathered -> equal to the single-domain result.
"""
import os
import socket

import numpy as np
import pytest

from lj_gpu_b200 import decomp

DENSITY, L, STEPS = 0.8, 17.3, 3


def test_slab_bookkeeping():
    # a slab thinner than the halo would need ghosts from rank +-2: refused, not silently wrong
    with pytest.raises(ValueError):
        decomp.make_slab(0, 8, DENSITY, 40.0)           # 23 layers / 8 ranks = 2-3 layers < halo of 3
    for world, box in ((1, 40.0), (2, 40.0), (3, 40.0), (4, 40.0), (8, 60.0)):
        s = decomp.lattice_spacing(DENSITY)
        n = int(box / s)
        slabs = [decomp.make_slab(r, world, DENSITY, box) for r in range(world)]
        assert slabs[0].lo == 0 and slabs[-1].hi == 4 * n ** 3
        for a, b in zip(slabs, slabs[1:]):
            assert a.hi == b.lo                      # contiguous, disjoint, complete
            assert a.n_hi == b.n_own - max(b.n_own - a.halo * a.layer, 0) or a.n_hi == a.halo * a.layer
        for sl in slabs:
            ib, ie = sl.interior_rows()
            assert 0 <= ib <= ie <= sl.n_own
            rows = [(0, ib), (ib, ie), (ie, sl.n_own)]
            assert sum(e - b for b, e in rows) == sl.n_own
            assert sl.n_local == sum(e - b for b, e in sl.global_ranges())
            # halo thick enough: a ghost layer further out cannot hold a neighbour
            assert sl.halo * s - 0.5 * s - 0.1 >= 3.3 - 1e-12 or sl.halo * s >= 3.3 + 0.5 * s + 0.1 - 1e-12
        # the plans of neighbouring ranks mirror each other
        nb = {sl.rank: sl for sl in slabs}
        plans = {r: decomp.halo_plan(nb[r], nb) for r in range(world)}
        for r in range(world):
            for peer, kind, b, e in plans[r]:
                other = [(k, bb, ee) for pr, k, bb, ee in plans[peer] if pr == r and k != kind]
                assert any(ee - bb == e - b for _, bb, ee in other)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    from oracle.ljoracle import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        oracle = Oracle()
        oracle.set_num_threads(1)
        q_all = oracle.init_fcc(DENSITY, L)
        slabs = {r: decomp.make_slab(r, world, DENSITY, L) for r in range(world)}
        slab = slabs[rank]
        # only the owned segment is filled; ghosts must arrive through the exchange
        q_local = torch.zeros(slab.n_local, 4, dtype=torch.float64)
        q_local[:slab.n_own, :3] = torch.from_numpy(q_all[slab.lo:slab.hi])
        plan = decomp.halo_plan(slab, slabs)
        decomp.exchange_halo(q_local, plan, dist)
        expect = decomp.local_positions(slab, q_all)
        assert np.array_equal(q_local[:, :3].numpy(), expect), "ghost positions differ"
        # per-rank compute by the oracle on local indices
        ql = np.ascontiguousarray(q_local[:, :3].numpy())
        nop, ptr, lst = oracle.makepair(ql, full=True)
        p = np.zeros_like(ql)
        oracle.force_gather(ql, p, nop, ptr, lst, steps=STEPS)
        np.save(os.path.join(out_dir, "p_%d.npy" % rank), p[:slab.n_own])
        np.save(os.path.join(out_dir, "nop_%d.npy" % rank), nop[:slab.n_own])
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_and_decomposed_forces_match_single_domain(world, tmp_path, oracle):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    q = oracle.init_fcc(DENSITY, L)
    nop, ptr, lst = oracle.makepair(q, full=True)
    p = np.zeros_like(q)
    oracle.force_gather(q, p, nop, ptr, lst, steps=STEPS)
    p_dec = np.concatenate([np.load(tmp_path / ("p_%d.npy" % r)) for r in range(world)])
    nop_dec = np.concatenate([np.load(tmp_path / ("nop_%d.npy" % r)) for r in range(world)])
    assert p_dec.shape == p.shape
    assert np.array_equal(nop_dec, nop)              # every owned row sees all its neighbours
    assert np.abs(p_dec - p).max() / np.abs(p).max() < 1e-13
