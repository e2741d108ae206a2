"""Replicate ensembles (BASELINE.json configs[4]): replicate r of an ensemble must be bit-identical to a single school
run through the cell-list path with the same seed and replicate id, and match the oracle like any school."""
import numpy as np
import pytest

from fish_abm_b200 import Ensemble, FishAbmError, School, synth
from oracle.pyoracle import Oracle, State

pytestmark = pytest.mark.gpu
FIELDS = ("coords", "dir", "speed", "speed_factor", "lag", "sample_rate", "food_intake")


def stack_states(states):
    return {k: np.stack([s[k] for s in states]) for k in FIELDS}


def make_states(R, n, w, seed, lattice=False):
    if lattice:
        return [synth.lattice_state(n, w, w, seed=seed + r) for r in range(R)]
    return [synth.uniform_state(n, w, w, seed=seed + r, speed=5.0 + r) for r in range(R)]


@pytest.mark.parametrize("n,w,R,steps,lattice", [(200, 2000, 6, 40, True), (200, 2000, 5, 25, False), (60, 2000, 3, 60, True),
                                                 (700, 2000, 2, 6, False), (90, 600, 3, 30, False)])
def test_ensemble_replicates_equal_single_schools(q100, n, w, R, steps, lattice):
    q = synth.tile_quality(q100, w, w)
    states = make_states(R, n, w, 300, lattice)
    seed, rep0 = 4242, 5
    with Ensemble(R, n=n, w=w, h=w, seed=seed, replicate0=rep0) as E:
        E.set_state(**stack_states(states))
        E.set_quality(q)
        g0 = E.get_state()
        zc0 = E.zone_counts()
        turn0 = E.social()[0]
        E.step(steps // 2)
        E.step(steps - steps // 2)
        g = E.get_state()
        zc1 = E.zone_counts()
        sums = E.sum_array()
        qual = [E.get_quality(r) for r in range(R)]
    for r in range(R):
        with School(n=n, w=w, h=w, seed=seed, replicate0=rep0 + r) as S:
            S.set_state(**states[r])
            S.set_quality(q)
            for k in FIELDS:
                assert np.array_equal(g0[k][r], S.get_state()[k]), ("round trip", k)
            assert np.array_equal(zc0[r], S.zone_counts())
            assert np.array_equal(turn0[r], S.social()[0])
            S.step(steps)
            gs = S.get_state()
            for k in FIELDS:
                assert np.array_equal(g[k][r], gs[k]), (r, k)
            assert np.array_equal(zc1[r], S.zone_counts())
            assert np.array_equal(qual[r], S.get_quality())
            assert int(np.atleast_1d(sums)[r]) == S.sum_array()


def test_ensemble_unfused_calls_and_oracle(q100):
    """move() / sample() on an ensemble against the oracle's synchronous step, replicate by replicate"""
    R, n, w = 3, 200, 2000
    states = make_states(R, n, w, 500)
    with Ensemble(R, n=n, w=w, h=w, seed=9, replicate0=0) as E:
        E.set_state(**stack_states(states))
        E.set_quality(q100)
        zc = E.zone_counts()
        for s in range(2):
            E.move()
            E.sample()
        g = E.get_state()
        assert E.step_index == 2
    for r in range(R):
        O = Oracle(State(**{k: v.copy() for k, v in states[r].items()}), w, w, q100.copy(), seed=9, replicate=r)
        assert np.array_equal(zc[r], O.zone_counts())
        for s in range(2):
            O.step_sync(s)
        # two compounding fp32 steps: integers exact, floats within the step tolerances
        for k in ("lag", "sample_rate", "food_intake"):
            assert np.array_equal(g[k][r], getattr(O.state, k)), k
        d = np.abs(g["coords"][r] - O.state.coords)
        assert np.minimum(d, w - d).max() < 2e-4
        assert np.abs(np.deg2rad(g["dir"][r] - O.state.dir)).max() < 2e-5


def test_ensemble_loggers_match_stepwise_observation(q100):
    R, n, steps = 4, 200, 50
    E = Ensemble(R, n=n, seed=77)
    E.initialize(5.0)
    E.set_quality(q100)
    sumq, intake = E.run_logged(steps, intake_every=10)
    assert sumq.shape == (steps, R) and intake.shape == (5, R, n)
    assert E.step_index == steps
    E.close()
    F = Ensemble(R, n=n, seed=77)
    F.initialize(5.0)
    F.set_quality(q100)
    for z in range(steps):
        F.step(1)
        assert np.array_equal(F.sum_array(), sumq[z]), z
        if z % 10 == 0:
            assert np.array_equal(F.get_state()["food_intake"], intake[z // 10]), z
    F.close()
    assert np.all(np.diff(sumq, axis=0) <= 0) and sumq[0].max() <= 3499


def test_ensemble_food_conservation_and_initialize_law(q100):
    R, n = 64, 200
    with Ensemble(R, n=n, seed=5) as E:
        E.initialize(5.0)
        E.set_quality(q100)
        g0 = E.get_state()
        # initialize(), fish_mod.cpp:183-206: integer lattice centre +-100, integer headings -45..45, lag 0
        assert np.all(g0["coords"] == np.round(g0["coords"])) and g0["coords"].min() >= 900 and g0["coords"].max() <= 1100
        assert np.all(g0["dir"] == np.round(g0["dir"])) and np.abs(g0["dir"]).max() <= 45
        assert len({g0["coords"][r].tobytes() for r in range(R)}) == R  # replicates differ
        E.step(300)
        g = E.get_state()
        eaten = g["food_intake"].sum(axis=1)
        assert np.array_equal(3499 - E.sum_array(), eaten)
        assert eaten.min() > 0


def test_ensemble_limits():
    with pytest.raises(FishAbmError):
        Ensemble(2, n=2000)
    with Ensemble(2, n=50) as E:
        E.initialize(5.0)
        with pytest.raises(FishAbmError):
            E.in_zone(330, 100.0)
        with pytest.raises(FishAbmError):
            E.feed([0])
