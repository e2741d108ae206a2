"""Live cross-check of the C restatement against the compiled, unmodified reference on fresh random states.
Skipped where oracle/_ref/ is absent (it is built only where /root/reference exists; the golden fixtures cover
the same ground everywhere else)."""
import numpy as np
import pytest

from fish_abm_b200 import synth
from oracle.pyoracle import Oracle, RefLib, State

pytestmark = pytest.mark.skipif(not (RefLib.available(200) and RefLib.available(1000)), reason="oracle/_ref not built")
FIELDS = ("coords", "dir", "speed", "speed_factor", "lag", "sample_rate", "food_intake")


@pytest.mark.parametrize("num,seed", [(200, 101), (200, 102), (1000, 103)])
def test_frozen_state_queries(q100, num, seed):
    st = State(**synth.uniform_state(num, 2000, 2000, seed=seed))
    R = RefLib(num); R.set_world(2000, 2000); R.set_state(st); R.set_quality(q100); R.rng(77, 0, 4)
    O = Oracle(st.copy(), 2000, 2000, q100.copy(), seed=77)
    want = np.stack([R.in_zone(330, 15), R.in_zone(330, 100), R.in_zone(330, 200), R.in_zone(180, 200)], 1)
    assert np.array_equal(O.zone_counts(), want)
    so = O.social(step=4)
    ok = so["n_ties"] == 0
    assert np.array_equal(so["turn"][ok], R.social()[ok])


@pytest.mark.parametrize("num,seed", [(200, 201), (1000, 202)])
def test_three_sequential_steps(q100, num, seed):
    st = State(**synth.uniform_state(num, 2000, 2000, seed=seed))
    R = RefLib(num); R.set_world(2000, 2000); R.set_state(st); R.set_quality(q100)
    O = Oracle(st.copy(), 2000, 2000, q100.copy(), seed=88)
    R.rng(88, 0, 0)
    R.run(3, 0, log_sumq=False)
    O.run(3, 0, sync=False, log_sumq=False)
    got = R.get_state()
    for f in FIELDS:
        assert np.array_equal(getattr(got, f), getattr(O.state, f)), f
    assert np.array_equal(R.get_quality(100, 100), O.quality)


def test_isolated_fish_random_walk(q100):
    """a fish with nobody in view turns by distrib_randomwalk (fish_mod.cpp:372-374): same draw on both sides"""
    d = synth.uniform_state(200, 2000, 2000, seed=301)
    d["coords"][:] = np.float32(50.0) + d["coords"] * np.float32(0.01)   # everybody in one corner ...
    d["coords"][0] = (1000.0, 1000.0)                                     # ... except fish 0
    d["lag"][:] = 0
    st = State(**d)
    R = RefLib(200); R.set_world(2000, 2000); R.set_state(st); R.set_quality(q100); R.rng(5, 0, 2)
    O = Oracle(st.copy(), 2000, 2000, q100.copy(), seed=5)
    so = O.social(step=2)
    assert so["randomwalk"][0] == 1 and abs(so["turn"][0]) <= 45
    assert so["turn"][0] == R.social(ids=np.array([0]))[0]
    R.move(); O.move_seq(2)
    assert np.array_equal(R.get_state().dir, O.state.dir)
