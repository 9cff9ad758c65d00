"""debug: where does the time of a decomposed cell-tile step go? (torchrun, 2 ranks)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from lj_gpu_b200 import decomp
s = (1.0 / 4.0) ** (-1.0 / 3.0)
system = decomp.DecomposedSystem(1.0, (80 + 0.05) * s, halo_mode=os.environ.get("LJ_HALO", "p2p"), tiles=True)
ctx = system.ctx
def timed(fn, reps=20):
    fn(); fn(); torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
n_own = system.slab.n_own
t_ct = timed(lambda: ctx.force_step(system.q, system.p, system.pl, rows=(0, n_own), variant="celltile"))
t_sw = timed(lambda: ctx.force_step(system.q, system.p, system.pl, rows=(0, n_own), variant="subwarp", group=8))
def halo_only():
    ev = system.halo(); system.compute.wait_event(ev)
t_h = timed(halo_only)
t_ser = timed(lambda: system.step(overlap=False, variant="auto"))
t_ser_sw = timed(lambda: system.step(overlap=False, variant="subwarp", group=8))
print("rank %d n_own=%d pn=%d: celltile %.4f  subwarp %.4f  halo %.4f  step(serial,auto) %.4f  step(serial,subwarp) %.4f ms"
      % (rank, n_own, system.q.shape[0], t_ct, t_sw, t_h, t_ser, t_ser_sw), flush=True)
dist.barrier(); dist.destroy_process_group()
