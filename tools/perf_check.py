"""Developer timing (GPU box): neighbour kernel designs at the benchmark state.  Same protocol as bench.py: 5 warm-up
steps, then 30 timed steps (device events inside the library), no L2 flush.  VARIANTS=v3,v4  N=1000000"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fish_abm_b200 import School, synth
q100 = np.loadtxt(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "quality.csv"), delimiter=",", dtype=np.int32)
n = int(os.environ.get("N", 1_000_000)); w = synth.world_size_for(n)
d = synth.uniform_state(n, w, w, seed=0x5EED); q = synth.tile_quality(q100, w, w)
for variant in os.environ.get("VARIANTS", "v3,v4").split(","):
    os.environ["FISHABM_NEIGHBOUR"] = variant
    S = School(n=n, w=w, h=w, seed=0x5EED); S.set_state(**d); S.set_quality(q)
    S.step(5); S.flush()
    for rep in range(2):
        S.step(30); ms = S.last_step_ms(); nb, l = S.last_neighbour_ms()
        print(f"{variant}: N={n} 30 steps {ms:.3f} ms -> {n*30/ms*1e3:.3e} agent-steps/s; neighbour {nb/l*1e3:.1f} us/launch x{l}; rest {(ms-nb)/30*1e3:.1f} us/step", flush=True)
    S.enable_stats(True); S.step(1); print("   stats", S.last_pass_stats()); S.close()
