"""Host-side statistics / figures restated from the reference's auxiliary code: the social weight function of heatmap.cpp
(the same three formulas social() uses, fish_mod.cpp:330,342,349) and the speed-factor histogram."""
import numpy as np

from fish_abm_b200 import stats


def test_social_weight_is_the_three_formulas_of_social():
    d = np.array([0.0, 3.0, 14.999, 15.0, 60.0, 99.999, 100.0, 150.0, 199.999, 200.0, 500.0])
    w = stats.social_weight(d)
    assert np.allclose(w[:3], 1 / (0.1 * d[:3] ** 2 + 2))                    # fish_mod.cpp:330
    assert np.allclose(w[3:6], 1 / (0.004 * (d[3:6] - 15) ** 2 + 2))          # :342
    assert np.allclose(w[6:9], 1 / (0.004 * (d[6:9] - 200) ** 2 + 2))         # :349
    assert w[9] == 0 and w[10] == 0
    assert w.max() == 0.5 and np.isclose(stats.social_weight(15.0), 0.5) and np.isclose(stats.social_weight(199.9999), 0.5, atol=1e-6)


def test_heatmap_matches_paint_pixel_by_pixel():
    img = stats.social_heatmap()
    assert img.shape == (410, 410) and img.dtype == np.uint8
    rng = np.random.default_rng(0)
    for x, y in rng.integers(0, 410, (300, 2)):     # paint(x, y, image), heatmap.cpp:107-122, scalar restatement
        d = np.sqrt((205 - x) ** 2 + (205 - y) ** 2)
        if d < 15: s = 1 / (0.1 * d ** 2 + 2)
        elif d < 100: s = 1 / (0.004 * (d - 15) ** 2 + 2)
        elif d < 200: s = 1 / (0.004 * (d - 200) ** 2 + 2)
        else: s = 0.0
        assert img[y, x] == min(255, int(510 * s)), (x, y)
    assert img[205, 205] == 255 and img[0, 0] == 0      # s = 0.5 at the centre; the corners are beyond 200
    ring = img[205, 205 + 15]                            # the alignment zone starts at full weight again
    assert ring == 255


def test_speed_factor_hist():
    num = 60.0
    sf = np.r_[np.full(5, 0.5), 1 + 2 * np.array([0, 0, 3, 7, 7, 7]) / num]
    h = stats.speed_factor_hist(sf, num)
    assert h["fed"] == 5 and h["front_count_hist"] == [2, 0, 0, 1, 0, 0, 0, 3] and h["fish"] == 11


def test_displacement_hist_is_periodic():
    a = np.array([[1.0, 1.0], [1999.0, 5.0], [10.0, 10.0]])
    b = np.array([[4.0, 5.0], [2.0, 5.0], [10.0, 10.0]])
    h = stats.displacement_hist(a, b, 2000, 2000)
    assert h["resting"] == 1 and np.isclose(h["max"], 5.0) and h["hist"][3] == 1 and h["hist"][5] == 1
