import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fish_abm_b200 import School, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
w = synth.world_size_for(n)
q100 = np.loadtxt(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "quality.csv"), delimiter=",", dtype=np.int32)
d = synth.uniform_state(n, w, w, seed=0x5EED); q = synth.tile_quality(q100, w, w)
S = School(n=n, w=w, h=w, seed=0x5EED); S.set_state(**d); S.set_quality(q)
S.step(25); S.flush()
S.enable_stats(True)
for _ in range(3):
    S.step(1); print(S.last_pass_stats())
