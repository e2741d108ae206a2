"""fp64 check mode (fishabm_check_fp64): a brute-force, double-precision, cell-list-free evaluation of in_zone and
social() on the device.  (i) It is itself checked against the CPU oracle; (ii) it then checks the product path at the
full benchmark sizes on thousands of fish -- far more than the CPU oracle's all-pairs loops can cover in seconds."""
import numpy as np
import pytest

from fish_abm_b200 import School, synth
from oracle.pyoracle import Oracle, State

pytestmark = pytest.mark.gpu
TOL_RAD = 1e-5


def test_check_mode_equals_the_cpu_oracle(q100):
    n, w = 20_000, 6400
    d = synth.uniform_state(n, w, w, seed=301)
    d["dir"][:2000] = np.random.default_rng(1).uniform(-290, 650, 2000).astype(np.float32)   # exotic raw headings too
    q = synth.tile_quality(q100, w, w)
    ids = np.random.default_rng(2).choice(n, 1500, replace=False).astype(np.int32)
    with School(n=n, w=w, h=w, seed=302) as S:
        S.set_state(**d)
        S.set_quality(q)
        counts, turn, flags = S.check_fp64(ids)
    O = Oracle(State(**{k: v.copy() for k, v in d.items()}), w, w, q.copy(), seed=302)
    assert np.array_equal(counts, O.zone_counts(ids))
    so = O.social(step=0, ids=ids)
    assert np.array_equal(flags & 1, so["randomwalk"])
    assert np.array_equal((flags & 2) != 0, so["n_ties"] > 0)
    ok = (so["n_ties"] == 0) & (so["knife"] == 0)
    assert ok.mean() > 0.95
    assert np.abs(turn - so["turn"])[ok].max() < 1e-9          # degrees: same formulas, different summation order / libm


@pytest.mark.parametrize("n,k", [(1_000_000, 4096), (16_777_216, 384)])
def test_product_path_against_check_mode_at_benchmark_sizes(q100, n, k):
    """BASELINE configs[2] and [3] sizes: counts bit-exact, turns within 1e-5 rad, on k sampled fish each compared
    against ALL n fish"""
    w = synth.world_size_for(n)
    d = synth.uniform_state(n, w, w, seed=303)
    ids = np.random.default_rng(3).choice(n, k, replace=False).astype(np.int32)
    with School(n=n, w=w, h=w, seed=304) as S:
        S.set_state(**d)
        S.set_quality(synth.tile_quality(q100, w, w))
        S.step(2)                       # a state the product itself produced (moved, re-sorted, partly fed)
        counts, turn, flags = S.check_fp64(ids)
        zc = S.zone_counts()[ids]
        st = S.social()[0][ids]
    assert np.array_equal(zc, counts)
    ok = (flags & 6) == 0
    assert ok.mean() > 0.98
    assert np.abs(np.deg2rad(st - turn))[ok].max() < TOL_RAD


def test_check_mode_limits(q100):
    from fish_abm_b200 import FishAbmError
    d = synth.uniform_state(500, 2000, 2000, seed=305)
    with School(n=500, seed=1, force_slab=True) as S:
        S.set_state(**d)
        with pytest.raises(FishAbmError):
            S.check_fp64([0, 1])
    with School(n=500, seed=1) as S:
        S.set_state(**d)
        c, t, f = S.check_fp64([3, 499, 777])      # an id outside the school is marked, not an error
        assert f[2] == -1 and np.isnan(t[2]) and (c[2] == -1).all() and f[0] >= 0


def test_brute_fp64_mode_steps_like_the_product_path(q100):
    """FISHABM_MODE_BRUTE_FP64: every neighbour pass is the fp64 all-pairs evaluation.  Its counts are the product path's
    (bit-exact), its turns within 1e-5 rad, and a few steps of it keep every integer field identical to the cell-list path
    (continuous state re-aligned each step so that a rounding difference cannot flip a later decision)."""
    from fish_abm_b200 import FishAbmError, School, _lib, synth
    n, w = 3000, 2000
    d = synth.uniform_state(n, w, w, seed=404)
    with School(n=n, w=w, h=w, seed=12, mode=_lib.MODE_BRUTE_FP64) as B, School(n=n, w=w, h=w, seed=12) as S:
        for X in (B, S):
            X.set_state(**d); X.set_quality(q100)
        assert np.array_equal(B.zone_counts(), S.zone_counts())
        tb, ts = B.social()[0], S.social()[0]
        cc, ct, cf = S.check_fp64(np.arange(n, dtype=np.int32))
        ok = (cf & 6) == 0
        assert np.abs(np.deg2rad(tb - ts))[ok].max() < 1e-5
        assert np.abs(np.deg2rad(tb - ct))[ok].max() < 1e-6      # the mode IS the check evaluation (fp32 only at the output)
        for step in range(5):
            B.step(1); S.step(1)
            gb, gs = B.get_state(), S.get_state()
            for k in ("lag", "sample_rate", "food_intake"):
                assert np.array_equal(gb[k], gs[k]), (step, k)
            assert np.array_equal(B.get_quality(), S.get_quality())
            assert np.abs(np.deg2rad(gb["dir"] - gs["dir"]))[ok].max() < 1e-5
            B.set_state(**gs)
            B.step_index = S.step_index
    with pytest.raises(FishAbmError):
        School(n=100_000, w=8000, h=8000, mode=_lib.MODE_BRUTE_FP64)
