import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
from conftest import golden_state
from fish_abm_b200 import School
from oracle.pyoracle import Oracle, State, load_quality_csv
from test_gpu_parity import as_dict, to32
g = np.load("tests/golden/golden_ref.npz"); q100 = load_quality_csv("tests/golden/quality.csv")
case = sys.argv[1] if len(sys.argv) > 1 else "E"
st = golden_state(g, f"{case}_in")
for variant in ("v1", "v3", "v4"):
    os.environ["FISHABM_NEIGHBOUR"] = variant
    S = School(n=st.n, seed=int(g["seed"])); S.set_state(**as_dict(st)); S.set_quality(q100); S.step_index = int(g[f"{case}_step"])
    O = Oracle(to32(st), 2000, 2000, q100.copy(), seed=int(g["seed"]))
    so = O.social(step=int(g[f"{case}_step"]))
    turn, na, nl, nt = S.social()
    err = np.abs(np.deg2rad(turn - so["turn"]))
    bad = np.argsort(-err)[:8]
    print(variant, "max", err.max())
    for b in bad:
        print(f"  id {b} dir {st.dir[b]:.3f} gpu {turn[b]:.5f} orc {so['turn'][b]:.5f} n {na[b]},{nl[b]},{nt[b]} orc {so['n_avoid'][b]},{so['n_align'][b]},{so['n_attract'][b]} ties {so['n_ties'][b]} knife {so['knife'][b]}")
