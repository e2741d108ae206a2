"""Short ensemble run for ncu captures.  usage: python tools/ens_target.py cohesive|dispersed [schools] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fish_abm_b200 import Ensemble, synth
mode = sys.argv[1] if len(sys.argv) > 1 else "cohesive"
R = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
n = 200
q100 = np.loadtxt(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "quality.csv"), delimiter=",", dtype=np.int32)
E = Ensemble(R, n=n, seed=0x5EED, speed_norm=float(n))
if mode == "cohesive":
    E.initialize(5.0)
else:
    st = [synth.uniform_state(n, 2000, 2000, seed=5 + r % 64) for r in range(R)]
    E.set_state(**{k: np.stack([s[k] for s in st]) for k in st[0]})
E.set_quality(q100)
E.step(30)
E.step(steps)
print(mode, R, "schools:", E.last_step_ms() / steps, "ms per step,", R * n * steps / (E.last_step_ms() * 1e-3), "agent-steps/s")
