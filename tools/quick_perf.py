"""Quick A/B timing of the neighbour kernel at the benchmark state (C3 by default): ms per fused step and per neighbour launch.
usage: [FISHABM_LIB=...] [FISHABM_NEIGHBOUR=v3] python tools/quick_perf.py [n_fish]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from fish_abm_b200 import School, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
w = synth.world_size_for(n)
q100 = np.loadtxt(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "quality.csv"), delimiter=",", dtype=np.int32)
d = synth.uniform_state(n, w, w, seed=0x5EED); q = synth.tile_quality(q100, w, w)
S = School(n=n, w=w, h=w, seed=0x5EED); S.set_state(**d); S.set_quality(q)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(5): S.step(1)
S.flush()
tot = nb = 0.0; k = 0
STEPS = int(os.environ.get("STEPS", 20))
series = []
for _ in range(STEPS):
    flush.fill_(1); torch.cuda.synchronize()
    S.step(1); tot += S.last_step_ms(); a, b = S.last_neighbour_ms(); nb += a; k += b; series.append(round(a * 1e3, 1))
print("  neighbour launch by step:", series[::5])
wt = wn = 0.0; wk = 0
for _ in range(20):   # L2-warm: no flush between steps (what a resident 1M-fish school sees in production)
    S.step(1); wt += S.last_step_ms(); a, b = S.last_neighbour_ms(); wn += a; wk += b
print(f"  L2-warm: step {wt/20*1e3:.1f} us, neighbour launch {wn/wk*1e3:.1f} us")
S.enable_phase_timing(True); S.step(10); ph = S.last_phase_ms(); S.enable_phase_timing(False); S.flush()
print(f"lib={os.environ.get('FISHABM_LIB','default')} variant={os.environ.get("FISHABM_NEIGHBOUR","v4")} n={n}: step {tot/STEPS*1e3:.1f} us, neighbour launch {nb/k*1e3:.1f} us, "
      f"phases/step (L2 warm) {({p: round(v/10*1e3,1) for p, v in ph.items() if p != 'steps'})}")
