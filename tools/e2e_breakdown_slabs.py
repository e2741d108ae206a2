"""Where a slab rank's end-to-end step goes (wall clock per C-ABI call, rank 0 prints): run under torchrun with N ranks.
usage: torchrun --nproc-per-node N tools/e2e_breakdown_slabs.py [n_fish]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from fish_abm_b200 import synth
from fish_abm_b200.slab import DistributedSchool
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16_777_216
w = synth.world_size_for(n)
q100 = np.loadtxt(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "quality.csv"), delimiter=",", dtype=np.int32)
d = synth.uniform_state(n, w, w, seed=0x5EED); q = synth.tile_quality(q100, w, w)
S = DistributedSchool(dist, rank, world, lr, transport="p2p", n=n, w=w, h=w, seed=0x5EED)
S.set_state(**d); S.set_quality(q); S.step(5); S.flush()
cap = int(S.slab_info()["capacity"])
shapes = dict(ids=((cap,), np.int32), coords=((cap, 2), np.float64), dir=((cap,), np.float64), speed=((cap,), np.float64),
              speed_factor=((cap,), np.float64), lag=((cap,), np.int32), sample_rate=((cap,), np.int32), food_intake=((cap,), np.int32))
bufs = [{k: torch.empty(shp, dtype=torch.from_numpy(np.empty(0, dt)).dtype).pin_memory().numpy() for k, (shp, dt) in shapes.items()} for _ in range(2)]
own = S.get_state_owned(out=bufs[0]); S.set_state_owned(**own); S.step(1); own = S.get_state_owned(out=bufs[1])
acc = {}
def timed(name, f):
    t0 = time.perf_counter(); r = f(); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0; return r
torch.cuda.synchronize(); dist.barrier()
K = 6
t0 = time.perf_counter()
for it in range(K):
    timed("set_state_owned(enqueue)", lambda: S.set_state_owned(**own, wait=False))
    timed("wait", lambda: S.wait())
    timed("step", lambda: S.step(1))
    own = timed("get_state_owned", lambda: S.get_state_owned(out=bufs[it & 1]))
    timed("sum_array", lambda: S.sum_array())
torch.cuda.synchronize(); dist.barrier()
tot = time.perf_counter() - t0
if rank == 0:
    print(f"{world} ranks, {len(own['ids'])} fish on rank 0 ({56 * len(own['ids']) / 1e6:.0f} MB each way): {tot / K * 1e3:.2f} ms per e2e step = {n * K / tot:.3e} agent-steps/s")
    for k, v in acc.items(): print(f"  {k:26s} {v / K * 1e3:8.3f} ms")
S.close()
dist.destroy_process_group()
