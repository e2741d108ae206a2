"""Developer check (GPU box): CUDA path vs the oracle on frozen states; prints differences."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fish_abm_b200 import School, synth
from oracle.pyoracle import Oracle, State, load_quality_csv

q100 = load_quality_csv(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "quality.csv"))

def check(n, w, seed=1, integer_dirs=False, steps=3):
    print(f"=== n={n} w={w} seed={seed}")
    d = synth.uniform_state(n, w, w, seed=seed, integer_dirs=integer_dirs)
    q = synth.tile_quality(q100, w, w)
    S = School(n=n, w=w, h=w, seed=99)
    S.set_state(**d); S.set_quality(q)
    g = S.get_state()
    for k in d: assert np.array_equal(g[k], d[k]), k
    O = Oracle(State(**{k: v.copy() for k, v in d.items()}), w, w, q.copy(), seed=99)
    t0 = time.time(); zc = O.zone_counts(); so = O.social(step=0); t1 = time.time()
    gz = S.zone_counts(); turn, na, nl, nt = S.social()
    print("oracle secs", t1 - t0)
    print("count mismatches", (gz != zc).sum(axis=0), "of", n)
    ok = (so['n_ties'] == 0)
    dt = np.abs(turn - so['turn'])[ok]
    print("turn max diff deg", dt.max(), "rad", np.deg2rad(dt.max()), "mean", dt.mean(), "rw", so['randomwalk'].sum())
    bad = np.argsort(-np.abs(turn - so['turn']))[:3]
    for b in bad: print("  worst", b, turn[b], so['turn'][b], so['n_avoid'][b], so['n_align'][b], so['n_attract'][b])
    print("gen in_zone(200,150)", (S.in_zone(200, 150.0) != O.in_zone(200, 150.0)).sum())
    # stepping: unfused move/sample vs oracle sync
    for s in range(steps):
        S.move(); O.move_sync(s)
        g = S.get_state()
        dd = np.abs(g['dir'] - O.state.dir).max(); dc = np.abs(g['coords'] - O.state.coords)
        dc = np.minimum(dc, w - dc).max()
        dsf = np.abs(g['speed_factor'] - O.state.speed_factor).max()
        print(f" step {s} move: ddir {dd:.3e} deg ({np.deg2rad(dd):.2e} rad) dpos {dc:.3e} dsf {dsf:.2e}")
        # re-sync oracle to the device state so that errors do not compound (frozen-state parity)
        O.state = State(**{k: v.copy() for k, v in g.items()})
        S.sample(); O.sample(s)
        g = S.get_state()
        print("   sample: lag", (g['lag'] != O.state.lag).sum(), "rate", (g['sample_rate'] != O.state.sample_rate).sum(),
              "intake", (g['food_intake'] != O.state.food_intake).sum(), "sf", np.abs(g['speed_factor'] - O.state.speed_factor).max(),
              "q", (S.get_quality() != O.quality).sum(), S.sum_array(), O.sum_quality())
    # fused vs unfused bit-identity
    A = School(n=n, w=w, h=w, seed=5); B = School(n=n, w=w, h=w, seed=5)
    for X in (A, B): X.set_state(**d); X.set_quality(q)
    A.step(4)
    for _ in range(4): B.move(); B.sample()
    ga, gb = A.get_state(), B.get_state()
    print("fused==unfused:", all(np.array_equal(ga[k], gb[k]) for k in ga), np.array_equal(A.get_quality(), B.get_quality()))
    S.enable_stats(True); S.zone_counts(); print(S.last_pass_stats())
    return S

check(2000, 2000, seed=1)
check(10000, 2000, seed=2, integer_dirs=True)
check(20000, 6400, seed=3)
# timing
n = 1_000_000; w = synth.world_size_for(n)
d = synth.uniform_state(n, w, w, seed=4); q = synth.tile_quality(q100, w, w)
S = School(n=n, w=w, h=w, seed=1); S.set_state(**d); S.set_quality(q)
S.step(5)
for k in (10, 50):
    S.step(k); ms = S.last_step_ms(); nb, l = S.last_neighbour_ms()
    print(f"1M: {k} steps {ms:.2f} ms -> {n*k/ms*1e3:.3e} agent-steps/s; neighbour {nb:.2f} ms in {l} launches")
