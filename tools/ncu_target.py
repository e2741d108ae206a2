"""Short run for ncu captures: a few fused steps of the benchmark state (N fish, default 1M).  After the captured steps the
same school is run again with the pair counters on, so that the focal / candidate / interacting counts of every neighbour
launch (launch k of both runs is the same work) can be put beside the capture: gpurun_out/ncu_target_stats.json."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fish_abm_b200 import School, synth
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
q100 = np.loadtxt(os.path.join(ROOT, "tests", "golden", "quality.csv"), delimiter=",", dtype=np.int32)
n = int(os.environ.get("N", 1_000_000)); w = synth.world_size_for(n); steps = int(os.environ.get("STEPS", 6))
d = synth.uniform_state(n, w, w, seed=0x5EED); q = synth.tile_quality(q100, w, w)
S = School(n=n, w=w, h=w, seed=1); S.set_state(**d); S.set_quality(q)
S.step(steps); S.flush()
print("ok", S.last_step_ms())
S.close()
if os.environ.get("NCU_TARGET_STATS", "1") != "0":
    S = School(n=n, w=w, h=w, seed=1); S.set_state(**d); S.set_quality(q)
    S.enable_stats(True)
    per = []
    for _ in range(steps):
        S.step(1); per.append(S.last_pass_stats())
    S.flush(); per.append(S.last_pass_stats())
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"n": n, "steps": steps, "per_neighbour_launch": per}, open(os.path.join(ROOT, "gpurun_out", "ncu_target_stats.json"), "w"))
