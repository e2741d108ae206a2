"""torchrun worker of tests/test_slab_gpu.py::test_nccl_slabs_equal_whole_school (and of the gloo host-logic test with
--dry): every rank runs its slab; rank 0 gathers the owned parts and saves the whole-school result."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fish_abm_b200 import synth  # noqa: E402
from fish_abm_b200.slab import DistributedSchool, merge_quality, merge_states, merge_owned  # noqa: E402

N, W, STEPS, SEED = 120_000, 16 * 200, 20, 91


def main():
    out = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    q100 = np.loadtxt(os.path.join(ROOT, "tests", "golden", "quality.csv"), delimiter=",", dtype=np.int32)
    d = synth.uniform_state(N, W, W, seed=51, speed=10.0)
    S = DistributedSchool(dist, rank, world, local, n=N, w=W, h=W, seed=SEED)
    S.set_state(**d)
    S.set_quality(synth.tile_quality(q100, W, W))
    S.step(STEPS // 2)
    S.step(STEPS - STEPS // 2)
    mine = dict(state=S.get_state(), quality=S.get_quality(), zone=S.zone_counts())
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    if rank == 0:
        st = merge_states([p["state"] for p in parts])
        np.savez(out, n=N, w=W, steps=STEPS, seed=SEED, quality=merge_quality([p["quality"] for p in parts]),
                 zone=merge_owned([p["zone"] for p in parts], -1), **st)
    S.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
