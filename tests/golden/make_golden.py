"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libfishref_<num>.so, built from
/root/reference/fish_mod.cpp by oracle/build_ref.sh).  Run in the development container, where the reference
exists; the fixtures travel to the GPU box, the reference does not.

    python tests/golden/make_golden.py
"""
import hashlib
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fish_abm_b200 import synth  # noqa: E402
from oracle.pyoracle import RefLib, State, load_quality_csv  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
REF_Q = "/root/reference/quality.csv"
SEED = 20261019


def state_fields(prefix, st):
    return {f"{prefix}_{k}": np.array(getattr(st, k)) for k in
            ("coords", "dir", "speed", "speed_factor", "lag", "sample_rate", "food_intake")}


def case(name, num, st: State, q100, step, out):
    R = RefLib(num)
    R.set_world(2000, 2000)
    R.set_state(st)
    R.set_quality(q100)
    R.rng(SEED, 0, step)
    out.update(state_fields(f"{name}_in", st))
    out[f"{name}_step"] = np.int64(step)
    out[f"{name}_zone"] = np.stack([R.in_zone(330, 15), R.in_zone(330, 100), R.in_zone(330, 200), R.in_zone(180, 200)], 1)
    out[f"{name}_zone_misc"] = np.stack([R.in_zone(200, 150.0), R.in_zone(90, 37.5), R.in_zone(360, 200)], 1)
    out[f"{name}_social"] = R.social(330)
    out[f"{name}_get_speed"] = R.get_speed_all()
    k = min(st.n, 64)
    ids = np.arange(k, dtype=np.int32)
    out[f"{name}_move_focal_ids"] = ids
    out[f"{name}_move_focal"] = np.stack([R.move_focal(int(i)) for i in ids])
    R.move()
    out.update(state_fields(f"{name}_moved", R.get_state()))
    R.sample()
    out.update(state_fields(f"{name}_sampled", R.get_state()))
    out[f"{name}_quality_after"] = R.get_quality(100, 100)
    out[f"{name}_sum_after"] = np.int64(R.sum_array())


def main():
    for num in (60, 200, 1000):
        assert RefLib.available(num), "run oracle/build_ref.sh first"
    q100 = load_quality_csv(REF_Q)
    assert q100.shape == (100, 100) and q100.sum() == 3499
    # the food environment is data, not source: it ships as a fixture with its checksum
    with open(REF_Q, "rb") as f:
        raw = f.read()
    with open(os.path.join(OUT, "quality.csv"), "wb") as f:
        f.write(raw)
    out = {"quality_sha256": np.array(hashlib.sha256(raw).hexdigest()), "seed": np.int64(SEED)}

    # A: the reference's own initial school (initialize(), fish_mod.cpp:183-206) at N=60
    R = RefLib(60)
    R.set_world(2000, 2000)
    R.rng(SEED, 0, 0)
    R.initialize(60, 5.0, 7)
    case("A", 60, R.get_state(), q100, 0, out)
    # A2: the same school after 100 reference steps (polarised, partly feeding)
    R.set_world(2000, 2000); R.rng(SEED, 0, 0); R.initialize(60, 5.0, 7); R.set_quality(q100)
    R.run(100, 0, log_sumq=False)
    case("A2", 60, R.get_state(), R.get_quality(100, 100), 100, out)
    # B: generic positions, N=200, arbitrary headings, mixed lag
    case("B", 200, State(**synth.uniform_state(200, 2000, 2000, seed=11)), q100, 3, out)
    # C: N=1000 (rho = 2.5e-4), integer headings
    case("C", 1000, State(**synth.uniform_state(1000, 2000, 2000, seed=12, integer_dirs=True)), q100, 5, out)
    # D: integer lattice, N=200 (coincident fish, exact-radius ties)
    case("D", 200, State(**synth.lattice_state(200, 2000, 2000, seed=13)), q100, 7, out)
    # E: out-of-range raw headings (getangle's single wrap)
    d = synth.uniform_state(200, 2000, 2000, seed=14)
    rng = np.random.default_rng(15)
    d["dir"] = rng.uniform(-290, 650, 200).astype(np.float32).astype(np.float64)
    d["lag"][:] = 0
    # a dense patch so that all three zones are populated
    d["coords"][:100] = (1000 + rng.uniform(-60, 60, (100, 2))).astype(np.float32)
    case("E", 200, State(**d), q100, 9, out)

    # scalar functions
    R = RefLib(60); R.set_world(2000, 2000); R.rng(SEED, 0, 0)
    rng = np.random.default_rng(21)
    vecs = [rng.uniform(-180, 180, rng.integers(1, 9)) for _ in range(200)]
    out["avg_flat"] = np.concatenate(vecs)
    out["avg_len"] = np.array([len(v) for v in vecs])
    out["avg_res0"] = np.array([R.averageturn(v, False) for v in vecs])
    out["avg_res1"] = np.array([R.averageturn(v[:2] if len(v) >= 2 else np.r_[v, v], True) for v in vecs])
    ga_in = np.column_stack([rng.uniform(0, 2000, (500, 2)), rng.uniform(-290, 650, 500), rng.uniform(0, 2000, (500, 2))])
    out["getangle_in"] = ga_in
    out["getangle_res"] = np.array([R.getangle(r[:2], r[2], r[3:5]) for r in ga_in])
    ca = np.r_[rng.uniform(-400, 800, 100), [-360, 0, 360, 720, -0.0]]
    out["correctangle_in"] = ca
    out["correctangle_res"] = np.array([R.correctangle(a) for a in ca])
    cc = np.r_[rng.uniform(-50, 2050, (100, 2)), [[2000, 0], [0, 2000], [2000.5, -0.5]]]
    out["correct_coords_in"] = cc
    out["correct_coords_res"] = np.array([R.correct_coords(r) for r in cc])

    # trajectory: 300 reference steps of the default school, sum_array per step and intake every 10
    R.set_world(2000, 2000); R.rng(SEED, 0, 0); R.initialize(60, 5.0, 7); R.set_quality(q100)
    out.update(state_fields("traj_in", R.get_state()))
    _, sumq, intake = R.run(300, 0, log_sumq=True, intake_every=10)
    out["traj_sumq"] = sumq
    out["traj_intake"] = intake
    out.update(state_fields("traj_out", R.get_state()))

    # libstdc++'s own uniform_int_distribution on the oracle's word stream
    with tempfile.TemporaryDirectory() as td:
        exe = os.path.join(td, "lemire_check")
        subprocess.run(["g++", "-O1", "-I", os.path.join(ROOT, "oracle"), os.path.join(ROOT, "oracle", "lemire_check.cpp"), "-o", exe], check=True)
        txt = subprocess.run([exe, str(SEED), "400"], check=True, capture_output=True, text=True).stdout
    out["lemire"] = np.array([[int(x) for x in ln.split()] for ln in txt.strip().splitlines()], np.int64)

    np.savez_compressed(os.path.join(OUT, "golden_ref.npz"), **out)
    print("wrote", os.path.join(OUT, "golden_ref.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
