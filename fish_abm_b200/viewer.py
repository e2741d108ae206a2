"""Optional viewer: the reference's picture of a step -- fish as 20-unit arrows with a centre dot
(drawfish, fish_mod.cpp:208-224) over the food grid painted q*127 per cell (:631-636, repainted per cell by feed :576-585),
blended 1.0 : 0.1 (addWeighted, :144) -- rendered on the HOST from fishabm_get_state / fishabm_get_quality every k steps.
Nothing here is on the update path; it needs OpenCV's Python module (cv2) only when called."""
from __future__ import annotations

import numpy as np


def environment_image(quality: np.ndarray, w: int, h: int) -> np.ndarray:
    """get_environment's painting (fish_mod.cpp:631-636): every food cell a grey square of value q * (255/2), int division"""
    rows, cols = quality.shape
    g = (np.clip(quality, 0, None) * (255 // 2)).astype(np.int64)
    g = np.clip(g, 0, 255).astype(np.uint8)   # Vec3b saturates
    img = np.repeat(np.repeat(g, h // rows, axis=0), w // cols, axis=1)
    return np.repeat(img[:, :, None], 3, axis=2)


def fish_colours(n: int, seed: int = 0) -> np.ndarray:
    """initialize() gives every fish a random BGR colour (rand() % 255, fish_mod.cpp:190-193)"""
    return np.random.default_rng(seed).integers(0, 255, (n, 3), dtype=np.int64)


def draw_frame(state: dict, quality: np.ndarray, w: int = 2000, h: int = 2000, colours: np.ndarray | None = None) -> np.ndarray:
    """one frame of fish_mod.avi: drawfish on black, then addWeighted(fish, 1, environment, 0.1)"""
    import cv2
    n = len(state["dir"])
    colours = fish_colours(n) if colours is None else colours
    fish = np.zeros((h, w, 3), np.uint8)
    c = np.cos(np.deg2rad(state["dir"])); s = np.sin(np.deg2rad(state["dir"]))
    xy = state["coords"]
    for i in range(n):
        col = tuple(int(v) for v in colours[i])
        centre = (int(xy[i, 0]), int(xy[i, 1]))
        tail = (int(xy[i, 0] - 10 * c[i]), int(xy[i, 1] - 10 * s[i]))
        head = (int(xy[i, 0] + 10 * c[i]), int(xy[i, 1] + 10 * s[i]))
        cv2.circle(fish, centre, 2, col, 1, cv2.LINE_AA)
        cv2.arrowedLine(fish, tail, head, col, 1, cv2.LINE_AA, 0, 0.2)
    return cv2.addWeighted(fish, 1.0, environment_image(quality, w, h), 0.1, 0.0)


def record(school, nsteps: int, every: int = 1, path: str | None = None, fps: int = 30, scale: float = 0.5):
    """step `school` (a fish_abm_b200.School) nsteps times; render every `every`-th step; optionally write a video like the
    reference's fish_mod.avi.  Returns the frames."""
    import cv2
    frames, writer = [], None
    colours = fish_colours(school.n)
    for z in range(0, nsteps, every):
        school.step(min(every, nsteps - z))
        f = draw_frame(school.get_state(), school.get_quality(), school.w, school.h, colours)
        if scale != 1.0:
            f = cv2.resize(f, None, fx=scale, fy=scale, interpolation=cv2.INTER_AREA)
        frames.append(f)
        if path:
            if writer is None:
                writer = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), fps, (f.shape[1], f.shape[0]))
            writer.write(f)
    if writer is not None:
        writer.release()
    return frames
