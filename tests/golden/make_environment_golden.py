"""tests/golden/environment_ref.npz: food maps produced by the UNMODIFIED reference's create_environment
(fish_mod.cpp:482-531) replayed through oracle/_ref with a live seed count.  Run in the development container."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import RefLib  # noqa: E402

CASES = [(0, 11, 0), (1, 12, 0), (3, 13, 7), (25, 14, 0), (200, 15, 3)]  # (n_seeds, seed, replicate)


def main():
    R = RefLib(60)
    R.set_world(2000, 2000)
    out = {"cases": np.array(CASES, np.int64)}
    for k, (n_seeds, seed, rep) in enumerate(CASES):
        out[f"q{k}"] = R.create_environment(n_seeds, seed, rep).astype(np.int8)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "environment_ref.npz"), **out)
    print({k: int(v.sum()) for k, v in out.items() if k != "cases"})


if __name__ == "__main__":
    main()
