"""Extracts the trajectory statistics the reference ships as data into a small fixture,
tests/golden/reference_stats.json.  Run in the development container (reads /root/reference/data).

  depletion: data/rolling_speed/rolling_speed<S>.csv -- 30 replicates x 2001 samples of sum_array(quality) per step
             (fish_mod.cpp:149-154); per speed: mean / sd of the final value, and the per-replicate lm(y~x) slope
             ("depletion rate", data/rolling_speed/stats_script.r:3-14)
  intake   : data/rolling_int_uptake_speed5/rolling_int_uptake_speed5_<k>.csv -- 500 rows (every 10th of 5000 steps,
             fish_mod.cpp:132-141) x 60 fish of cumulative food_intake; final-row mean / variance / histogram with
             breaks seq(0,80,4) (rolling_int_uptake_stats_script.r:34-40)
"rolling" = the committed fish_mod.cpp (get_speed active); the non-rolling variant's code is not in the repository.
"""
import glob
import json
import os

import numpy as np

REF = "/root/reference/data"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_stats.json")


def slopes(m):
    x = np.arange(m.shape[1], dtype=np.float64)
    return np.array([np.polyfit(x, row, 1)[0] for row in m])


def main():
    out = {"source": "fritzfrancisco/fish_abm data/ (N=60, quality.csv, 2000x2000 window)", "depletion": {}, "intake": {}}
    for variant in ("rolling", "norolling"):
        dep = {}
        for f in sorted(glob.glob(f"{REF}/{variant}_speed/{variant}_speed*.csv")):
            name = os.path.basename(f)[len(variant) + 6:-4]
            if name.endswith("_"):
                continue
            m = np.loadtxt(f, delimiter=",")
            m = m[:, :2000] if m.shape[1] < 2001 else m
            sl = slopes(m)
            dep[name] = {"replicates": int(m.shape[0]), "samples": int(m.shape[1]), "start": float(m[:, 0].mean()),
                         "final_mean": float(m[:, -1].mean()), "final_sd": float(m[:, -1].std(ddof=1)),
                         "final_min": float(m[:, -1].min()), "final_max": float(m[:, -1].max()),
                         "slope_median": float(np.median(sl)), "slope_mean": float(sl.mean()), "slope_sd": float(sl.std(ddof=1)),
                         "mean_curve_every_100": [float(v) for v in m.mean(axis=0)[::100]]}
        out["depletion"][variant] = dep
        files = sorted(glob.glob(f"{REF}/{variant}_int_uptake_speed5/{variant}_int_uptake_speed5_*.csv"))
        fin = np.concatenate([np.loadtxt(f, delimiter=",")[-1] for f in files])
        hist, _ = np.histogram(fin, bins=np.arange(0, 84, 4))
        out["intake"][variant] = {"replicates": len(files), "fish": 60, "steps": 5000, "final_mean": float(fin.mean()),
                                  "final_var": float(fin.var(ddof=1)), "final_min": float(fin.min()), "final_max": float(fin.max()),
                                  "hist_breaks_0_80_4": [int(v) for v in hist]}
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
