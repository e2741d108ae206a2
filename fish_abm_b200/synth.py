"""Synthetic school states and world sizes for the benchmark and parity configurations
(SURVEY.md section 8d).  Host-side numpy only; nothing here touches the device or the oracle.

All positions are fp32-representable (and so are their in-cell offsets) so that a host fp64 checker and the
device see identical inputs.
"""
from __future__ import annotations

import math

import numpy as np

CELL = 200  # neighbour cell = largest interaction radius (fish_mod.cpp:275,589)
FOOD_CELL = 20  # Structure.res_factor (fish_mod.cpp:56)


def world_size_for(n: int, fish_per_cell: float = 20.0) -> int:
    """W = H = 200*ceil(sqrt(N/20)): about 20 fish per 200x200 neighbour cell (rho ~ 5e-4)."""
    return CELL * max(3, math.ceil(math.sqrt(n / fish_per_cell)))


def tile_quality(q100: np.ndarray, w: int, h: int) -> np.ndarray:
    """quality.csv (100x100 food cells of 20 units) tiled periodically to (h/20) x (w/20)."""
    rows, cols = h // FOOD_CELL, w // FOOD_CELL
    ry, rx = -(-rows // q100.shape[0]), -(-cols // q100.shape[1])
    return np.ascontiguousarray(np.tile(q100, (ry, rx))[:rows, :cols]).astype(np.int32)


def uniform_state(n: int, w: int, h: int, seed: int = 0x5EED, speed: float = 5.0, integer_dirs: bool = False,
                  lagged_fraction: float = 0.6):
    """Generic-position state: uniform positions (fp32), headings uniform in [0,360) (or integers in
    -45..45 as initialize() draws them, fish_mod.cpp:196), lag 0 w.p. 0.4 else 1..10 (the 59 % lagged steady
    state of the reference school), sample_rate 0..30, speed_factor 1."""
    rng = np.random.Generator(np.random.Philox(seed))
    coords = np.empty((n, 2), np.float64)
    coords[:, 0] = (rng.random(n, np.float32) * np.float32(w)).astype(np.float32)
    coords[:, 1] = (rng.random(n, np.float32) * np.float32(h)).astype(np.float32)
    np.clip(coords[:, 0], 0, np.nextafter(np.float32(w), np.float32(0)), out=coords[:, 0])
    np.clip(coords[:, 1], 0, np.nextafter(np.float32(h), np.float32(0)), out=coords[:, 1])
    if integer_dirs:
        d = rng.integers(-45, 46, n).astype(np.float64)
    else:
        d = (rng.random(n, np.float32) * np.float32(360.0)).astype(np.float32).astype(np.float64)
    lag = np.where(rng.random(n) < (1.0 - lagged_fraction), 0, rng.integers(1, 11, n)).astype(np.int32)
    rate = rng.integers(0, 31, n).astype(np.int32)
    return dict(coords=coords, dir=d, speed=np.full(n, float(speed)), speed_factor=np.ones(n),
                lag=lag, sample_rate=rate, food_intake=np.zeros(n, np.int32))


def lattice_state(n: int, w: int, h: int, seed: int = 7, spread: int = 100, speed: float = 5.0):
    """initialize()'s own law (fish_mod.cpp:183-206): integer lattice centre +- spread, integer headings
    -45..45, lag 0.  Exercises the NaN drops (coincident / collinear fish) and exact r = 15/100/200 ties."""
    rng = np.random.Generator(np.random.Philox(seed))
    coords = np.empty((n, 2), np.float64)
    coords[:, 0] = w // 2 + rng.integers(0, 2 * spread + 1, n) - spread
    coords[:, 1] = h // 2 + rng.integers(0, 2 * spread + 1, n) - spread
    d = rng.integers(-45, 46, n).astype(np.float64)
    return dict(coords=coords, dir=d, speed=np.full(n, float(speed)), speed_factor=np.ones(n),
                lag=np.zeros(n, np.int32), sample_rate=np.zeros(n, np.int32), food_intake=np.zeros(n, np.int32))
