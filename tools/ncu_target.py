"""Short run for ncu captures: a few fused steps of the benchmark state (N fish, default 1M)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fish_abm_b200 import School, synth
q100 = np.loadtxt(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "quality.csv"), delimiter=",", dtype=np.int32)
n = int(os.environ.get("N", 1_000_000)); w = synth.world_size_for(n)
d = synth.uniform_state(n, w, w, seed=0x5EED); q = synth.tile_quality(q100, w, w)
S = School(n=n, w=w, h=w, seed=1); S.set_state(**d); S.set_quality(q)
S.step(int(os.environ.get("STEPS", 6))); S.flush()
print("ok", S.last_step_ms())
