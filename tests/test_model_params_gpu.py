"""Run-time model parameters (include/fishabm.h: fishabm_model_params; the reference's literals, fish_mod.cpp:28-33,385-386,
546-548,566-573): non-default values must match the oracle evaluated with the same values, and an ensemble whose replicates
each carry their own parameters (BASELINE.json configs[4]: a sweep) must equal single schools created with those parameters."""
import numpy as np
import pytest

from fish_abm_b200 import Ensemble, FishAbmError, School, synth
from oracle.pyoracle import Oracle, State

pytestmark = pytest.mark.gpu
FIELDS = ("coords", "dir", "speed", "speed_factor", "lag", "sample_rate", "food_intake")
MODEL = dict(lag_after_feed=4, sample_rate_fed=22, sample_rate_unfed=3, sample_draw_max=25, error_range=6, randomwalk_range=30,
             alone_align_min=3, fed_speed_factor=0.75, weight_align=1.2, weight_attract=0.9)


def test_non_default_model_matches_the_oracle(q100):
    n, w = 3000, 2000
    d = synth.uniform_state(n, w, w, seed=77)
    with School(n=n, w=w, h=w, seed=5, model=MODEL) as S:
        S.set_state(**d); S.set_quality(q100)
        O = Oracle(State(**{k: v.copy() for k, v in d.items()}), w, w, q100.copy(), seed=5, model=MODEL)
        so = O.social(step=0)
        ok = (so["n_ties"] == 0) & (so["knife"] == 0)
        assert np.abs(np.deg2rad(S.social()[0] - so["turn"]))[ok].max() < 1e-5      # the 1.2 / 0.9 combination weights
        for step in range(6):
            S.step(1)
            O.step_sync(step)
            g = S.get_state()
            for k in ("lag", "sample_rate", "food_intake"):
                assert np.array_equal(g[k], getattr(O.state, k)), (step, k)
            assert np.array_equal(S.get_quality(), O.quality), step
            # re-align the continuous state so that rounding differences cannot flip a later integer decision
            O.state.coords[:] = g["coords"]; O.state.dir[:] = g["dir"]; O.state.speed_factor[:] = g["speed_factor"]
        assert g["lag"].max() <= 4 and set(np.unique(g["sample_rate"])) <= set(range(0, 100))
        assert (g["speed_factor"] == np.float32(0.75)).any()                         # somebody fed


def test_default_model_is_the_reference(q100):
    """explicit defaults == no model argument, bit for bit"""
    n, w = 2000, 2000
    d = synth.uniform_state(n, w, w, seed=78)
    res = []
    for model in (None, dict(lag_after_feed=10, sample_rate_fed=30, sample_rate_unfed=0, sample_draw_max=35, error_range=10,
                             randomwalk_range=45, alone_align_min=5, fed_speed_factor=0.5, weight_align=1.7, weight_attract=1.7 - 1)):
        with School(n=n, w=w, h=w, seed=9, model=model) as S:
            S.set_state(**d); S.set_quality(q100)
            S.step(8)
            res.append((S.get_state(), S.get_quality()))
    for k in FIELDS:
        assert np.array_equal(res[0][0][k], res[1][0][k]), k
    assert np.array_equal(res[0][1], res[1][1])


def test_ensemble_sweep_equals_single_schools(q100):
    R, n, w, steps = 6, 200, 2000, 30
    states = [synth.lattice_state(n, w, w, seed=500 + r) for r in range(R)]
    models = [dict(sample_rate_fed=15 + 5 * r, fed_speed_factor=0.25 + 0.1 * r, lag_after_feed=6 + r) for r in range(R)]
    with Ensemble(R, n=n, w=w, h=w, seed=31, replicate0=2) as E:
        E.set_state(**{k: np.stack([s[k] for s in states]) for k in FIELDS})
        E.set_quality(q100)
        E.set_models(models)
        E.step(steps)
        g = E.get_state()
        sums = E.sum_array()
    assert len(set(int(v) for v in sums)) > 1    # the sweep changes the depletion
    for r in range(R):
        with School(n=n, w=w, h=w, seed=31, replicate0=2 + r, model=models[r]) as S:
            S.set_state(**states[r]); S.set_quality(q100)
            S.step(steps)
            gs = S.get_state()
            for k in FIELDS:
                assert np.array_equal(g[k][r], gs[k]), (r, k)
            assert int(sums[r]) == S.sum_array()


def test_set_models_on_a_single_school_and_validation(q100):
    n, w = 500, 2000
    d = synth.uniform_state(n, w, w, seed=79)
    with School(n=n, w=w, h=w, seed=3) as S:
        S.set_state(**d); S.set_quality(q100)
        S.set_models([MODEL])
        S.step(3)
        a = S.get_state()
        with pytest.raises(FishAbmError):
            S.set_models([dict(error_range=500)])
        with pytest.raises(FishAbmError):
            S.set_models([MODEL, MODEL])
    with School(n=n, w=w, h=w, seed=3, model=MODEL) as S:
        S.set_state(**d); S.set_quality(q100)
        S.step(3)
        b = S.get_state()
    for k in FIELDS:
        assert np.array_equal(a[k], b[k]), k
    with pytest.raises(FishAbmError):
        School(n=n, w=w, h=w, model=dict(fed_speed_factor=-1.0))
