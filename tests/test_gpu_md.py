"""MD step around the force call: kick = lj_force_step, drift, skin-triggered
rebuild, energies.  The reference has no such loop (its q is static); the checker is the CPU oracle
for the kick and the list plus plain numpy for drift / displacement / energy."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RHO, L, DT, STEPS = 0.8, 15.5, 0.002, 220
CUT, SEARCH = 3.0, 3.3


def numpy_energy(q, p, nop, ptr, lst):
    i = np.repeat(np.arange(len(nop)), nop)
    d = q[lst] - q[i]
    r2 = d[:, 2] * d[:, 2] + (d[:, 1] * d[:, 1] + d[:, 0] * d[:, 0])
    x3 = (1.0 / r2[r2 <= CUT * CUT]) ** 3
    return 0.5 * (p * p).sum(), 0.5 * (4.0 * (x3 * x3 - x3)).sum()


@pytest.mark.parametrize("layout", ["aos4", "soa"])
def test_md_run_matches_cpu_and_conserves_energy(oracle, layout):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from lj_gpu_b200 import LJContext
    ctx = LJContext(0)
    q0 = oracle.init_fcc(RHO, L)
    pn = len(q0)
    # ---- CPU: oracle kick + numpy drift, rebuild when max displacement > skin/2
    q, p = q0.copy(), np.zeros_like(q0)
    nop, ptr, lst = oracle.makepair(q, search_len=SEARCH, full=True)
    ke0, pe0 = numpy_energy(q, p, nop, ptr, lst)
    q_ref, rebuilds = q.copy(), 0
    for s in range(STEPS):
        oracle.force_gather(q, p, nop, ptr, lst, steps=1, dt=DT, cl2=CUT * CUT)
        q += p * DT
        if ((q - q_ref) ** 2).sum(1).max() > (0.5 * (SEARCH - CUT)) ** 2:
            nop, ptr, lst = oracle.makepair(q, search_len=SEARCH, full=True)
            q_ref = q.copy()
            rebuilds += 1
    ke1, pe1 = numpy_energy(q, p, nop, ptr, lst)
    assert rebuilds >= 1                      # the trigger really fired
    # ---- GPU through the C ABI
    if layout == "aos4":
        qh = np.zeros((pn, 4)); qh[:, :3] = q0
        kw = {}
    else:
        qh = np.ascontiguousarray(q0.T)
        kw = dict(layout="soa", pn=pn)
    qd = torch.from_numpy(qh).cuda(); pd = torch.zeros_like(qd)
    pl = ctx.makepair(qd, search_len=SEARCH, capacity=int(1.2 * len(lst)) + 4096, **kw)
    gke0, gpe0 = ctx.energy(qd, pd, pl, **kw)
    assert gke0 == 0.0 and abs(gpe0 - pe0) <= 1e-12 * abs(pe0)
    n_reb = ctx.md_run(qd, pd, pl, STEPS, dt=DT, search_len=SEARCH, cutoff=CUT, **kw)
    assert n_reb == rebuilds
    qg = qd.cpu().numpy(); pg = pd.cpu().numpy()
    qg, pg = (qg[:, :3], pg[:, :3]) if layout == "aos4" else (qg.T, pg.T)
    assert np.abs(qg - q).max() < 1e-9 and np.abs(pg - p).max() / np.abs(p).max() < 1e-9
    gke1, gpe1 = ctx.energy(qd, pd, pl, **kw)
    assert abs(gke1 - ke1) <= 1e-9 * abs(ke1) and abs(gpe1 - pe1) <= 1e-9 * abs(pe1)
    # first-order symplectic Euler at dt = 0.002 on an unshifted, truncated potential, starting from a
    # strained lattice: the energy error is O(dt) and bounded (0.6 % here); CPU and GPU agree on it
    e0, e1 = gke0 + gpe0, gke1 + gpe1
    assert abs(e1 - e0) / abs(e0) < 1e-2, (e0, e1)
    assert abs((e1 - e0) - ((ke1 + pe1) - (ke0 + pe0))) <= 1e-9 * abs(e0)
    assert gke1 > 1.0                          # the system really moved (lattice + jitter relaxes)
    ctx.close()
