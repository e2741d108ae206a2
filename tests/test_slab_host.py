"""Host-side logic of the slab decomposition, on CPU: the column split and ring the library uses, assembling per-slab
outputs, and the NCCL-id bootstrap over a world_size-2 torch.distributed group (gloo)."""
import os
import socket

import numpy as np
import pytest

from fish_abm_b200 import _lib
from fish_abm_b200.school import School
from fish_abm_b200.slab import merge_owned, merge_quality, merge_states, partition, ring


def test_partition_covers_every_column_once():
    for nx in (3, 10, 14, 224, 916):
        for k in (1, 2, 3, 4, 7, 8):
            if k > nx:
                continue
            p = partition(nx, k)
            assert p[0][0] == 0 and p[-1][1] == nx
            assert all(a[1] == b[0] for a, b in zip(p, p[1:]))
            widths = [b - a for a, b in p]
            assert min(widths) >= 1 and max(widths) - min(widths) <= 1


def test_c4_partition_matches_the_survey():
    # 916 cell columns -> 114/115 per GPU at 8 GPUs
    assert sorted({b - a for a, b in partition(916, 8)}) == [114, 115]


def test_ring_is_periodic():
    assert ring(0, 4) == (3, 1) and ring(3, 4) == (2, 0)
    assert ring(0, 2) == (1, 1) and ring(0, 1) == (0, 0)


def test_merge_owned_and_states():
    a = np.array([1, -1, -1, 4]); b = np.array([-1, 2, 3, -1])
    assert merge_owned([a, b], -1).tolist() == [1, 2, 3, 4]
    with pytest.raises(ValueError):
        merge_owned([a, a], -1)
    t0 = np.array([0.5, np.nan]); t1 = np.array([np.nan, -2.0])
    assert merge_owned([t0, t1], float("nan")).tolist() == [0.5, -2.0]
    NO = School.NOT_OWNED
    s0 = dict(lag=np.array([0, NO], np.int32), dir=np.array([1.0, np.nan]))
    s1 = dict(lag=np.array([NO, 7], np.int32), dir=np.array([np.nan, 2.0]))
    m = merge_states([s0, s1])
    assert m["lag"].tolist() == [0, 7] and m["dir"].tolist() == [1.0, 2.0]
    with pytest.raises(ValueError):
        merge_states([s0, s0])
    q0 = np.array([[1, -1], [2, -1]]); q1 = np.array([[-1, 0], [-1, 3]])
    assert merge_quality([q0, q1]).tolist() == [[1, 0], [2, 3]]
    with pytest.raises(ValueError):
        merge_quality([q0, q0])


def _bootstrap_worker(rank, world, port, q):
    import torch.distributed as dist
    from fish_abm_b200.slab import broadcast_unique_id
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    uid = broadcast_unique_id(dist, rank)
    from fish_abm_b200.slab import gather_endpoints
    eps = gather_endpoints(dist, bytes([rank + 1]) * 96)   # the 96-byte inbox endpoints travel the same way
    assert eps == [bytes([1]) * 96, bytes([2]) * 96]
    q.put((rank, uid))
    dist.barrier()
    dist.destroy_process_group()


def test_unique_id_bootstrap_over_gloo_world_size_2():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bootstrap_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(got[0]) == _lib.UNIQUE_ID_BYTES and got[0] == got[1] and any(got[0])
