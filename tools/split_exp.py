import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np
from fish_abm_b200 import School, synth
q100 = np.loadtxt("/root/repo/tests/golden/quality.csv", delimiter=",", dtype=np.int32)
for n in (30_000, 100_000, 300_000, 1_000_000):
    w = synth.world_size_for(n); d = synth.uniform_state(n, w, w, seed=1); q = synth.tile_quality(q100, w, w)
    for sp in (1, 2, 3, 4, 8):
        os.environ["FISHABM_SPLIT"] = str(sp)
        S = School(n=n, w=w, h=w, seed=1); S.set_state(**d); S.set_quality(q); S.step(3); S.flush()
        S.step(40); nb, l = S.last_neighbour_ms(); print(f"N={n:8d} cells={(w//200)**2:6d} split={sp}: neighbour {nb/l*1e3:7.1f} us", flush=True); S.close()
