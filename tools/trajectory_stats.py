"""Trajectory statistics of the GPU path as replicate ensembles, next to the numbers the reference ships as data
(tests/golden/reference_stats.json).  usage: trajectory_stats.py [replicates=512] [out.json]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fish_abm_b200 import Ensemble
from fish_abm_b200.stats import depletion_stats, intake_stats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
q100 = np.loadtxt(os.path.join(ROOT, "tests", "golden", "quality.csv"), delimiter=",", dtype=np.int32)
ref = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_stats.json")))
R = int(sys.argv[1]) if len(sys.argv) > 1 else 512
out = {"replicates": R, "depletion": {}, "intake": {}}
for rolling in (True, False):
    name = "rolling" if rolling else "norolling"
    out["depletion"][name] = {}
    for sp in ("2.5", "5", "7.5", "10", "12.5", "15", "17.5", "20"):
        t0 = time.time()
        with Ensemble(R, n=60, seed=2026, speed_norm=60.0, rolling=rolling) as E:
            E.initialize(float(sp)); E.set_quality(q100)
            sumq, _ = E.run_logged(2000)
            ms = E.last_step_ms()
        st = depletion_stats(np.concatenate([np.full((1, R), 3499), sumq]).T)
        out["depletion"][name][sp] = st
        r = ref["depletion"][name][sp]
        print(f"{name:9s} speed {sp:5s}: final {st['final_mean']:7.1f} +- {st['final_sd']:5.1f} (ref {r['final_mean']:7.1f} +- {r['final_sd']:5.1f})  slope median {st['slope_median']:.3f} (ref {r['slope_median']:.3f})   [{ms:.0f} ms GPU, {time.time()-t0:.1f} s]", flush=True)
    with Ensemble(R, n=60, seed=77, speed_norm=60.0, rolling=rolling) as E:
        E.initialize(5.0); E.set_quality(q100)
        _, intake = E.run_logged(5000, log_sumq=False, intake_every=10)
    st = intake_stats(intake[-1].reshape(-1))
    out["intake"][name] = st
    r = ref["intake"][name]
    print(f"{name:9s} intake after 5000 steps: mean {st['final_mean']:.2f} var {st['final_var']:.1f} (ref mean {r['final_mean']:.2f} var {r['final_var']:.1f})", flush=True)
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], "w"), indent=1)
