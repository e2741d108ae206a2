"""The statistics the reference's R scripts compute on its logged runs, restated in numpy (no Rscript in the image).

depletion_stats: data/rolling_speed/stats_script.r:3-14 -- per replicate lm(y ~ x) of sum(quality) against the step index,
                 the slope is the "depletion rate"; plus the final value's mean / sd.
intake_stats   : data/rolling_int_uptake_speed5/rolling_int_uptake_stats_script.r:34-40 -- final cumulative food_intake
                 per fish: mean, variance (var.test compares variances), histogram with breaks seq(0,80,4).
Host-side numpy only."""
from __future__ import annotations

import numpy as np


def depletion_slopes(curves: np.ndarray) -> np.ndarray:
    """curves [replicates, samples] -> least-squares slope of each row against 0..samples-1"""
    y = np.asarray(curves, np.float64)
    x = np.arange(y.shape[1], dtype=np.float64)
    xc = x - x.mean()
    return (y - y.mean(axis=1, keepdims=True)) @ xc / (xc @ xc)


def depletion_stats(curves: np.ndarray) -> dict:
    y = np.asarray(curves, np.float64)
    sl = depletion_slopes(y)
    return {"replicates": int(y.shape[0]), "samples": int(y.shape[1]), "final_mean": float(y[:, -1].mean()),
            "final_sd": float(y[:, -1].std(ddof=1)) if y.shape[0] > 1 else 0.0, "slope_median": float(np.median(sl)),
            "slope_mean": float(sl.mean()), "slope_sd": float(sl.std(ddof=1)) if y.shape[0] > 1 else 0.0}


def intake_stats(final_intake: np.ndarray) -> dict:
    v = np.asarray(final_intake, np.float64).reshape(-1)
    hist, _ = np.histogram(v, bins=np.arange(0, 84, 4))
    return {"fish": int(v.size), "final_mean": float(v.mean()), "final_var": float(v.var(ddof=1)), "final_min": float(v.min()),
            "final_max": float(v.max()), "hist_breaks_0_80_4": [int(h) for h in hist]}


# ---- the social weight function and its figure (heatmap.cpp; the same three formulas as social(), fish_mod.cpp:330,342,349) ----
def social_weight(d):
    """weight a neighbour's bearing is multiplied by, as a function of its distance d (any array):
         d < 15        avoid    1 / (0.1 d^2 + 2)                 fish_mod.cpp:330,  heatmap.cpp:110
         15 <= d < 100 align    1 / (0.004 (d - 15)^2 + 2)        fish_mod.cpp:342,  heatmap.cpp:114
         100 <= d < 200 attract 1 / (0.004 (d - 200)^2 + 2)       fish_mod.cpp:349,  heatmap.cpp:118
         d >= 200      0 (outside every zone; heatmap.cpp leaves the pixel black)"""
    d = np.asarray(d, np.float64)
    out = np.zeros_like(d)
    m = d < 15
    out[m] = 1.0 / (0.1 * d[m] ** 2 + 2.0)
    m = (d >= 15) & (d < 100)
    out[m] = 1.0 / (0.004 * (d[m] - 15.0) ** 2 + 2.0)
    m = (d >= 100) & (d < 200)
    out[m] = 1.0 / (0.004 * (d[m] - 200.0) ** 2 + 2.0)
    return out


def social_heatmap(size: int = 410, centre: int = 205) -> np.ndarray:
    """paint() of heatmap.cpp:107-122 over its 410 x 410 square (main, heatmap.cpp:34-38): grey level (uint8) of every pixel
    = the C++ conversion of 510 * s to an 8-bit channel (truncation of the double, modulo 256 for s = 0.5 -> 255),
    s = social_weight(distance to (205, 205)).  The annotations / colour map / inversion of the figure are viewer-only."""
    y, x = np.mgrid[0:size, 0:size]
    d = np.sqrt((centre - x).astype(np.float64) ** 2 + (centre - y).astype(np.float64) ** 2)
    v = np.floor(510.0 * social_weight(d))
    return np.where(v > 255, 255, v).astype(np.uint8)   # cv::Vec3b saturates (saturate_cast<uchar>)


# ---- speed-factor statistics (north_star: "speed and uptake histograms") --------------------------------------------------
def speed_factor_hist(speed_factor, speed_norm: float | None = None) -> dict:
    """histogram of get_speed's value 1 + 2 * front / num (fish_mod.cpp:588-591) over the fish: speed_factor takes the value
    0.5 right after a successful feed() (:566), else 1 + 2k/num for an integer front count k.  Bins are those values:
    'fed' (0.5) and front counts 0..kmax."""
    v = np.asarray(speed_factor, np.float64).reshape(-1)
    num = float(speed_norm if speed_norm else v.size)
    fed = np.isclose(v, 0.5)
    k = np.rint((v[~fed] - 1.0) * num / 2.0).astype(np.int64)
    kmax = int(k.max()) if k.size else 0
    return {"fish": int(v.size), "fed": int(fed.sum()), "front_count_hist": np.bincount(np.clip(k, 0, None), minlength=kmax + 1).tolist(),
            "mean": float(v.mean()), "var": float(v.var(ddof=1)) if v.size > 1 else 0.0, "max": float(v.max())}


def displacement_hist(coords_before, coords_after, w: int, h: int, bins=None) -> dict:
    """histogram of the distance each fish covered in one step (speed * speed_factor for a moving fish, 0 for a lagged one,
    fish_mod.cpp:171-179), minimum-image in the periodic window"""
    a, b = np.asarray(coords_before, np.float64), np.asarray(coords_after, np.float64)
    d = np.abs(b - a)
    d[:, 0] = np.minimum(d[:, 0], w - d[:, 0]); d[:, 1] = np.minimum(d[:, 1], h - d[:, 1])
    dist = np.hypot(d[:, 0], d[:, 1])
    bins = np.arange(0.0, 16.0, 1.0) if bins is None else np.asarray(bins, np.float64)
    hist, _ = np.histogram(dist, bins=bins)
    return {"fish": int(dist.size), "resting": int((dist == 0).sum()), "mean": float(dist.mean()), "max": float(dist.max()),
            "bins": bins.tolist(), "hist": hist.tolist()}
