"""A model check of scan_chained_kernel's protocol (fish_abm_b200/csrc/kernels.cuh): tiles take tickets in any order of
arrival, publish their tile sum, look back over the published words 32 at a time until they meet a tile that knows its
prefix, and publish their own.  The model interleaves the tiles' steps at random (a tile may only READ words that have been
published -- otherwise it keeps waiting, as the kernel's spin does) and checks every exclusive prefix, for many tile
counts and schedules, including stale words of the previous epoch.  No GPU needed."""
import numpy as np

SUM, PREFIX = 1, 2


def run(ntiles, rng, epoch, desc):
    totals = rng.integers(0, 50, ntiles)
    # state per tile: 0 = not started, 1 = published its sum (tile 0: its prefix), 2 = looking back, 3 = done
    state = np.zeros(ntiles, int)
    look = np.zeros(ntiles, int)
    acc = np.zeros(ntiles, int)
    result = np.full(ntiles, -1)
    started = 0                      # tickets are handed out in order: tile t cannot start before tile t-1 took its ticket
    while (state != 3).any():
        movable = [t for t in range(ntiles) if state[t] != 3 and t <= started]
        t = movable[rng.integers(len(movable))]
        if state[t] == 0:
            desc[t] = (epoch, PREFIX if t == 0 else SUM, int(totals[t]))
            if t == 0:
                result[0] = 0; state[0] = 3
            else:
                state[t] = 2; look[t] = t - 1
            started = max(started, t + 1)
        elif state[t] == 2:
            idx = [look[t] - l for l in range(32) if look[t] - l >= 0]
            words = [desc[i] for i in idx]
            if any(w[0] != epoch for w in words):
                continue             # some word is not published in THIS scan yet: the warp keeps spinning
            first = next((l for l, w in enumerate(words) if w[1] == PREFIX), None)
            upto = first if first is not None else 31
            acc[t] += sum(w[2] for l, w in enumerate(words) if l <= upto)
            if first is not None:
                result[t] = acc[t]
                desc[t] = (epoch, PREFIX, int(acc[t] + totals[t]))
                state[t] = 3
            else:
                look[t] -= 32
    expect = np.concatenate([[0], np.cumsum(totals)[:-1]])
    assert np.array_equal(result, expect), (ntiles, result, expect)


def test_chained_scan_protocol_under_random_schedules():
    rng = np.random.default_rng(5)
    for ntiles in (1, 2, 3, 25, 31, 32, 33, 64, 65, 100, 410):
        desc = [(0, 0, 0)] * ntiles                       # cleared at creation
        for epoch in range(1, 6):                         # words of earlier scans stay behind: the epoch invalidates them
            run(ntiles, rng, epoch, desc)
