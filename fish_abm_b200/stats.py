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
