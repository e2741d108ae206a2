import numpy as np

from fish_abm_b200 import synth


def test_world_sizes_of_the_benchmark_configs():
    assert synth.world_size_for(1_000_000) == 44_800
    assert synth.world_size_for(16_777_216) == 183_200
    assert synth.world_size_for(60) == 600


def test_uniform_state_is_fp32_representable_and_in_range():
    d = synth.uniform_state(5000, 4400, 4400, seed=3)
    c = d["coords"]
    assert np.array_equal(c.astype(np.float32).astype(np.float64), c)
    assert c.min() >= 0 and c[:, 0].max() < 4400 and c[:, 1].max() < 4400
    off = c - 200 * np.floor(c / 200)
    assert np.array_equal(off.astype(np.float32).astype(np.float64), off)  # in-cell offsets exact too
    assert set(np.unique(d["lag"])) <= set(range(11))
    assert 0.5 < (d["lag"] > 0).mean() < 0.7
    assert d["sample_rate"].min() >= 0 and d["sample_rate"].max() <= 30


def test_tile_quality(q100):
    q = synth.tile_quality(q100, 4400, 2600)
    assert q.shape == (130, 220)
    assert np.array_equal(q[:100, :100], q100)
    assert np.array_equal(q[100:130, 100:200], q100[:30, :])


def test_lattice_state_is_integer():
    d = synth.lattice_state(200, 2000, 2000)
    assert np.array_equal(d["coords"], np.round(d["coords"]))
    assert np.abs(d["coords"] - 1000).max() <= 100


def test_stats_restate_the_r_scripts():
    """fish_abm_b200/stats.py: lm(y ~ x) slope per replicate (stats_script.r:3-14) and the uptake summary"""
    from fish_abm_b200.stats import depletion_slopes, depletion_stats, intake_stats
    x = np.arange(2001.0)
    curves = np.stack([3499 - 0.5 * x, 3499 - 0.7 * x + 3 * np.sin(x), np.full_like(x, 3499.0)])
    sl = depletion_slopes(curves)
    assert abs(sl[0] + 0.5) < 1e-12 and abs(sl[1] + 0.7) < 1e-3 and abs(sl[2]) < 1e-12
    for k in range(3):
        assert abs(sl[k] - np.polyfit(x, curves[k], 1)[0]) < 1e-9
    st = depletion_stats(curves)
    assert st["replicates"] == 3 and st["samples"] == 2001 and st["slope_median"] == np.median(sl)
    it = intake_stats(np.array([3, 4, 39, 40, 79]))
    assert it["fish"] == 5 and it["final_mean"] == 33.0 and sum(it["hist_breaks_0_80_4"]) == 5
    assert it["hist_breaks_0_80_4"][0] == 1 and it["hist_breaks_0_80_4"][1] == 1 and it["hist_breaks_0_80_4"][9] == 1 and it["hist_breaks_0_80_4"][10] == 1
