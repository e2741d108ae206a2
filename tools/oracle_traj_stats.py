"""Developer evidence (CPU, oracle): final sum(quality) after 2000 steps, N=60, for the reference's sequential update
(orc *_seq) and the synchronous update the GPU implements (*_sync), same draws.  Writes tests/golden/oracle_traj_stats.json.  usage: oracle_traj_stats.py reps speeds..."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.pyoracle import Oracle, State, load_quality_csv, draw

q100 = load_quality_csv(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "quality.csv"))
reps = int(sys.argv[1]); speeds = [float(s) for s in sys.argv[2:]]
ref = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "reference_stats.json")))["depletion"]["rolling"]

def init_state(seed, rep, n=60, speed=5.0):
    # initialize(), fish_mod.cpp:183-206 with the library's streams (csrc/fishabm.cu fishabm_initialize)
    d = np.array([draw(seed, rep, 0, i, 2, -45, 45) for i in range(n)], np.float64)
    x = np.array([1000 + draw(seed, rep, 0, i, 6, 0, 200) - 100 for i in range(n)], np.float64)
    y = np.array([1000 + draw(seed, rep, 1, i, 6, 0, 200) - 100 for i in range(n)], np.float64)
    return State(np.stack([x, y], 1), d, np.full(n, speed), np.ones(n), np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32))

out = {"source": "oracle/fish_oracle.c orc_run, N=60, quality.csv, 2000 steps, initialize() law; seq = the reference's in-place id-ordered "
                 "update (pinned against the compiled fish_mod.cpp), sync = the frozen-state update the GPU implements", "speeds": {}}
for sp in speeds:
    fin = {True: [], False: []}
    for sync in (False, True):
        for r in range(reps):
            O = Oracle(init_state(123, r, speed=sp), 2000, 2000, q100.copy(), seed=123, replicate=r, speed_norm=60.0)
            sumq, _ = O.run(2000, sync=sync)
            fin[sync].append(sumq[-1])
    key = str(sp).rstrip("0").rstrip(".")
    out["speeds"][key] = {"replicates": reps, "seq_mean": float(np.mean(fin[False])), "seq_sd": float(np.std(fin[False], ddof=1)),
                          "sync_mean": float(np.mean(fin[True])), "sync_sd": float(np.std(fin[True], ddof=1))}
    json.dump(out, open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "oracle_traj_stats.json"), "w"), indent=1)
    print(f"speed {sp}: reference data {ref[key]['final_mean']:.0f} +- {ref[key]['final_sd']:.0f} | oracle seq {np.mean(fin[False]):.0f} +- {np.std(fin[False], ddof=1):.0f} | oracle sync {np.mean(fin[True]):.0f} +- {np.std(fin[True], ddof=1):.0f}  ({reps} reps)", flush=True)
