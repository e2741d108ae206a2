"""Where the end-to-end step (host buffers in, one step, host buffers out) spends its time: wall clock of each C-ABI call of
bench.py's e2e loop at the benchmark state, pinned host buffers.  usage: python tools/e2e_breakdown.py [n_fish]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from fish_abm_b200 import School, synth
from fish_abm_b200.school import _p
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
w = synth.world_size_for(n)
q100 = np.loadtxt(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "quality.csv"), delimiter=",", dtype=np.int32)
d = synth.uniform_state(n, w, w, seed=0x5EED); q = synth.tile_quality(q100, w, w)
S = School(n=n, w=w, h=w, seed=0x5EED); S.set_state(**d); S.set_quality(q)
S.step(5); S.flush()
order = ("coords", "dir", "speed", "speed_factor", "lag", "sample_rate", "food_intake")
pinned = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in S.get_state().items()}
host = {k: v.numpy() for k, v in pinned.items()}
out = {k: torch.empty_like(v).pin_memory().numpy() for k, v in pinned.items()}
sumq = np.zeros(1, np.int64)
bufs = [host, out]
acc = {"set_state": 0.0, "step": 0.0, "get_state": 0.0, "sum_quality": 0.0}
def call(name, f):
    t0 = time.perf_counter(); S._ck(f()); acc[name] += time.perf_counter() - t0
def one():
    src, dst = bufs
    call("set_state", lambda: S.L.fishabm_set_state(S.ctx, *(_p(src[k]) for k in order)))
    call("step", lambda: S.L.fishabm_step(S.ctx, 1))
    call("get_state", lambda: S.L.fishabm_get_state(S.ctx, *(_p(dst[k]) for k in order)))
    call("sum_quality", lambda: S.L.fishabm_sum_quality(S.ctx, _p(sumq)))
    bufs.reverse()
one(); torch.cuda.synchronize()
for k in acc: acc[k] = 0.0
K = 10
t0 = time.perf_counter()
for _ in range(K): one()
torch.cuda.synchronize()
tot = time.perf_counter() - t0
mb = sum(v.nbytes for v in host.values()) / 1e6
print(f"n={n}: {tot/K*1e3:.3f} ms per e2e step ({n*K/tot:.3e} agent-steps/s), {mb:.1f} MB each way")
for k, v in acc.items(): print(f"  {k:12s} {v/K*1e3:7.3f} ms")
# raw PCIe: the same bytes as plain copies
dev = {k: torch.empty_like(v, device="cuda") for k, v in pinned.items()}
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(K):
    for k in order: dev[k].copy_(pinned[k], non_blocking=True)
torch.cuda.synchronize(); h2d = (time.perf_counter() - t0) / K
t0 = time.perf_counter()
for _ in range(K):
    for k in order: pinned[k].copy_(dev[k], non_blocking=True)
torch.cuda.synchronize(); d2h = (time.perf_counter() - t0) / K
print(f"  plain copies of the same arrays: H2D {h2d*1e3:.3f} ms ({mb/h2d/1e3:.1f} GB/s), D2H {d2h*1e3:.3f} ms ({mb/d2h/1e3:.1f} GB/s)")
