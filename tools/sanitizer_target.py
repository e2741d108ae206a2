"""Small run touching every kernel (whole school, forced slab, in-process peer-memory slabs, ensemble, queries) for
compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fish_abm_b200 import Ensemble, School, synth
from fish_abm_b200.slab import LocalSlabs
q100 = np.loadtxt(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "quality.csv"), delimiter=",", dtype=np.int32)
n, w = 6000, 2000
d = synth.uniform_state(n, w, w, seed=3, speed=10.0)
d["coords"][:600] = (np.array([900.0, 900.0]) + np.random.default_rng(1).uniform(0, 150, (600, 2))).astype(np.float32)  # >256 candidates: multi-batch
with School(n=n, w=w, h=w, seed=5) as S:
    S.set_state(**d); S.set_quality(q100)
    S.zone_counts(); S.social(); S.get_speed(); S.in_zone(200, 150.0); S.feed(np.arange(10, dtype=np.int32))
    S.move(); S.sample(); S.step(3); ref = S.get_state(); S.sum_array(); S.get_quality()
with School(n=n, w=w, h=w, seed=5, force_slab=True) as S:
    S.set_state(**d); S.set_quality(q100); S.step(3); S.get_state(); own = S.get_state_owned(); S.set_state_owned(**own); S.step(1)
os.environ["FISHABM_SLAB_TIMEOUT_MS"] = "60000"
with LocalSlabs(2, transport="p2p", n=n, w=w, h=w, seed=5) as P:
    P.set_state(**d); P.set_quality(q100); P.step(3); P.get_state()
with LocalSlabs(3, n=n, w=w, h=w, seed=5) as P:
    P.set_state(**d); P.set_quality(q100); P.step(2); P.get_state()
with Ensemble(6, n=200, seed=9) as E:
    E.initialize(5.0); E.set_quality(q100); E.step(5); E.run_logged(12, intake_every=5); E.zone_counts(); E.move(); E.sample(); E.get_state()
print("sanitizer target done")
