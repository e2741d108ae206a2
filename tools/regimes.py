"""Developer timing (GPU box): the step in regimes other than the benchmark state."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fish_abm_b200 import Ensemble, School, synth
q100 = np.loadtxt(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "quality.csv"), delimiter=",", dtype=np.int32)

def run(tag, n, w, steps, lagged=0.6, **kw):
    d = synth.uniform_state(n, w, w, seed=1, lagged_fraction=lagged)
    S = School(n=n, w=w, h=w, seed=1, **kw); S.set_state(**d); S.set_quality(synth.tile_quality(q100, w, w))
    S.step(3); S.flush()
    S.step(steps); ms = S.last_step_ms(); nb, l = S.last_neighbour_ms()
    print(f"{tag:34s} N={n:8d} W={w:6d}: {ms/steps*1e3:9.1f} us/step  {n*steps/ms*1e3:.3e} agent-steps/s  (neighbour {nb/l*1e3:.1f} us)", flush=True)
    S.close()

run("C1 default school size", 60, 2000, 500)
E = Ensemble(1, n=60, seed=1); E.initialize(5.0); E.set_quality(q100); E.step(10); E.step(500)
print(f"{'C1 as a 1-school ensemble launch':34s} N=      60 W=  2000: {E.last_step_ms()/500*1e3:9.1f} us/step  {60*500/E.last_step_ms()*1e3:.3e} agent-steps/s", flush=True); E.close()
run("C2 10k in the reference window", 10_000, 2000, 100)
run("100k, rho=5e-4", 100_000, synth.world_size_for(100_000), 100)
run("C3 1M, 60% lagged", 1_000_000, 44_800, 30)
run("C3 1M, nobody lagged (t=0 of a run)", 1_000_000, 44_800, 10, lagged=0.0)
run("dense: 200k in 4000^2 (500/cell)", 200_000, 4000, 5)
run("sparse: 1M in 200k^2 (1/cell)", 1_000_000, 200_000, 30)
