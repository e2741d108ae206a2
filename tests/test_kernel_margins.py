"""The two numerical margins the neighbour kernel's exact counts rest on (fish_abm_b200/csrc/neighbour_v4.cuh, kernels.cuh),
re-derived on the CPU with numpy's IEEE fp32 / fp16 arithmetic:

 * the shared y frame (yframe_shift): positions are in-cell fp32 offsets; shifting a candidate and the focal fish into the
   frame of a cell pair rounds each once, and the rounding bands of pair_math / pair_fast (BAND_R = 2e-4 around the radii,
   1e-4/d + 2e-6 around the bearing thresholds) must cover the resulting error in distance and bearing;
 * the packed-fp16 disc filter: every pair that could be counted (true distance < 200 + band) must pass it.
No GPU needed."""
import numpy as np

CELL = np.float32(200.0)
BAND_R = 2e-4


def frame_pairs(rng, n):
    """random (focal, candidate) pairs in adjacent cells: raw in-cell offsets (fp32), candidate cell (kx, ky) in -1..1,
    focal cell row parity (even rows compute in the frame of the odd row above them)"""
    fo = rng.uniform(0, 200, (n, 2)).astype(np.float32)
    qo = rng.uniform(0, 200, (n, 2)).astype(np.float32)
    # the interesting offsets: right at the cell edges and at binade boundaries
    edge = rng.integers(0, 6, n)
    fo[edge == 0, 1] = np.nextafter(np.float32(200), np.float32(0))
    qo[edge == 1, 1] = np.float32(0.0)
    fo[edge == 2, 1] = np.float32(99.99999)
    qo[edge == 3, 1] = np.float32(127.99999)
    k = rng.integers(-1, 2, (n, 2))
    even = rng.integers(0, 2, n).astype(bool)
    return fo, qo, k, even


def test_shared_frame_rounding_stays_inside_the_bands():
    rng = np.random.default_rng(11)
    fo, qo, k, even = frame_pairs(rng, 2_000_000)
    ysh = np.where(even, -CELL, np.float32(0)).astype(np.float32)            # yframe_shift
    # what the kernels compute (one fp32 rounding per shifted coordinate, then the differences)
    qx = (qo[:, 0] + (k[:, 0] * 200).astype(np.float32)).astype(np.float32)
    qy = (qo[:, 1] + ((k[:, 1] * 200).astype(np.float32) + ysh)).astype(np.float32)
    fy = (fo[:, 1] + ysh).astype(np.float32)
    dx32 = (qx - fo[:, 0]).astype(np.float32).astype(np.float64)
    dy32 = (qy - fy).astype(np.float32).astype(np.float64)
    # the exact geometry of the same fp32 offsets (what exact_pair_bits evaluates in fp64)
    dx = qo[:, 0].astype(np.float64) - fo[:, 0].astype(np.float64) + k[:, 0] * 200.0
    dy = qo[:, 1].astype(np.float64) - fo[:, 1].astype(np.float64) + k[:, 1] * 200.0
    ex, ey = np.abs(dx32 - dx), np.abs(dy32 - dy)
    assert ex.max() <= 1.6e-5 + 1.6e-5      # candidate shift + the difference's own rounding (|.| < 400: ulp 3.05e-5)
    assert ey.max() <= 1.6e-5 + 0.8e-5 + 1.6e-5
    d = np.hypot(dx, dy)
    err_d = np.abs(np.hypot(dx32, dy32) - d)
    near = d < 201.0
    assert err_d[near].max() < 0.25 * BAND_R     # the distance band (2e-4) has a factor of four to spare
    # bearing: an error e in position turns the direction by at most e / d; the band is 1e-4 / d + 2e-6
    ang = np.hypot(ex, ey)[near & (d > 1e-3)]
    assert (ang < 0.5 * 1e-4).all()


def half(x):
    return np.asarray(x, dtype=np.float64).astype(np.float16).astype(np.float64)


def test_fp16_disc_filter_never_rejects_a_pair_that_could_count():
    """HADD2 / HMUL2 / HFMA2 with round-to-nearest-even fp16 results, as in the kernel: stored coordinates
    RN_half((x - 100) / 4) and RN_half((y - hcy) / 4); test d2 < 2528"""
    rng = np.random.default_rng(12)
    n = 2_000_000
    fo, qo, k, even = frame_pairs(rng, n)
    # pairs on a ring around the interaction radius (the only ones the filter could wrongly reject), any direction
    r = rng.uniform(199.0, 200.0003, n)
    phi = rng.uniform(0, 2 * np.pi, n)
    ysh = np.where(even, -200.0, 0.0)
    fx, fy = fo[:, 0].astype(np.float64), fo[:, 1].astype(np.float64) + ysh
    qx, qy = fx + r * np.cos(phi), fy + r * np.sin(phi)
    keep = (qx > -200) & (qx < 400) & (qy > -400) & (qy < 400)      # inside a window
    fx, fy, qx, qy, r, even = fx[keep], fy[keep], qx[keep], qy[keep], r[keep], even[keep]
    for hcy in (0.0, 100.0):                                         # the two centres the kernel uses for y
        sel = even if hcy == 0.0 else ~even                           # (hcy = 100 only with the odd-row frame: y in [-200, 400))
        hfx, hfy = half((fx[sel] - 100) * 0.25), half((fy[sel] - hcy) * 0.25)
        hqx, hqy = half((qx[sel] - 100) * 0.25), half((qy[sel] - hcy) * 0.25)
        dxh, dyh = half(hqx - hfx), half(hqy - hfy)                  # HADD2 (one rounding each)
        d2 = half(dxh * dxh + half(dyh * dyh))                       # HMUL2, then HFMA2 (single rounding of the fma)
        passed = d2 < 2528.0
        assert passed.all(), f"{(~passed).sum()} pairs within 200.0003 units were rejected (hcy={hcy})"
    # and the filter does filter: pairs beyond 201.5 units never pass
    r2 = rng.uniform(201.5, 400.0, fx.size)
    qx2, qy2 = fx + r2 * np.cos(phi[keep]), fy + r2 * np.sin(phi[keep])
    ok = (qx2 > -200) & (qx2 < 400) & (qy2 > -400) & (qy2 < 400)
    hfx, hfy = half((fx[ok] - 100) * 0.25), half(fy[ok] * 0.25)
    hqx, hqy = half((qx2[ok] - 100) * 0.25), half(qy2[ok] * 0.25)
    dxh, dyh = half(hqx - hfx), half(hqy - hfy)
    with np.errstate(over="ignore"):
        d2 = half(dxh * dxh + half(dyh * dyh))
    assert not (d2 < 2528.0).any()
