"""Long runs: invariants that must hold after thousands of steps (every fish accounted for and inside the world, food
conserved, bounded integer state), and slab runs that must stay bit-identical to the whole school while hundreds of
fish migrate back and forth."""
import numpy as np
import pytest

from fish_abm_b200 import School, synth
from fish_abm_b200.slab import LocalSlabs

pytestmark = pytest.mark.gpu
FIELDS = ("coords", "dir", "speed", "speed_factor", "lag", "sample_rate", "food_intake")


def check_invariants(g, w, h, q_now, q0_sum):
    assert np.isfinite(g["coords"]).all() and np.isfinite(g["dir"]).all()
    assert (g["coords"] >= 0).all() and (g["coords"][:, 0] < w).all() and (g["coords"][:, 1] < h).all()
    assert g["lag"].min() >= 0 and g["lag"].max() <= 10
    assert g["sample_rate"].min() >= 0
    assert g["food_intake"].min() >= 0
    assert ((g["speed_factor"] == 0.5) | ((g["speed_factor"] >= 1.0) & (g["speed_factor"] <= 3.0))).all()
    assert q_now.min() >= 0
    assert int(q_now.sum()) + int(g["food_intake"].sum()) == q0_sum


def test_two_thousand_steps_keep_every_invariant(q100):
    n, w = 200_000, 20_000
    d = synth.uniform_state(n, w, w, seed=201)
    q = synth.tile_quality(q100, w, w)
    q0 = int(q.sum())
    with School(n=n, w=w, h=w, seed=202) as S:
        S.set_state(**d)
        S.set_quality(q)
        for k in range(4):
            S.step(500)
            g = S.get_state()
            check_invariants(g, w, w, S.get_quality(), q0)
            zc = S.zone_counts()
            assert (zc[:, 0] <= zc[:, 1]).all() and (zc[:, 1] <= zc[:, 2]).all() and (zc[:, 3] <= zc.max()).all()
        assert S.step_index == 2000
        assert g["food_intake"].sum() > n // 4          # the school has been feeding
        assert np.abs(g["dir"]).max() < 1200              # headings stay within the range the reference can reach
        moved = np.abs(g["coords"] - d["coords"]).max(axis=1)
        assert (moved > 1).mean() > 0.99


def test_slabs_stay_identical_over_600_steps(q100, monkeypatch):
    monkeypatch.setenv("FISHABM_SLAB_TIMEOUT_MS", "8000")
    n, w, steps = 40_000, 9 * 200 * 2, 600   # 18 cell columns over 4 slabs; speed 15: a fish crosses a slab in ~100 steps
    d = synth.uniform_state(n, w, w, seed=203, speed=15.0)
    q = synth.tile_quality(q100, w, w)
    with School(n=n, w=w, h=w, seed=204) as S:
        S.set_state(**d); S.set_quality(q); S.step(steps)
        g, qa = S.get_state(), S.get_quality()
    check_invariants(g, w, w, qa, int(q.sum()))
    import torch
    devices = list(range(4)) if torch.cuda.device_count() >= 4 else None   # one GPU per slab: peer memory; else the hooks
    with LocalSlabs(4, devices=devices, transport="p2p", n=n, w=w, h=w, seed=204) as P:
        P.set_state(**d); P.set_quality(q)
        owned0 = [s.slab_info()["owned"] for s in P.S]
        P.step(steps)
        gp = P.get_state()
        for k in FIELDS:
            assert np.array_equal(gp[k], g[k]), k
        assert np.array_equal(P.get_quality(), qa)
        owned1 = [s.slab_info()["owned"] for s in P.S]
        assert sum(owned1) == n and owned1 != owned0     # fish did migrate


def test_ensemble_replicate_stays_identical_to_a_school_over_1500_steps(q100):
    from fish_abm_b200 import Ensemble
    n, steps, seed = 200, 1500, 205
    with Ensemble(3, n=n, seed=seed, replicate0=11, speed_norm=float(n)) as E:
        E.initialize(5.0)
        E.set_quality(q100)
        g0 = E.get_state()
        E.step(steps)
        ge = E.get_state()
        qe = [E.get_quality(r) for r in range(3)]
    for r in (0, 2):
        with School(n=n, seed=seed, replicate0=11 + r, speed_norm=float(n)) as S:
            S.set_state(**{k: v[r] for k, v in g0.items()})
            S.set_quality(q100)
            S.step(steps)
            gs = S.get_state()
            for k in FIELDS:
                assert np.array_equal(ge[k][r], gs[k]), (r, k)
            assert np.array_equal(qe[r], S.get_quality())
            check_invariants(gs, 2000, 2000, S.get_quality(), 3499)
