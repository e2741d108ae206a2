"""Slab decomposition (one school split along x): any number of slabs must reproduce the whole-school
run bit for bit -- positions, headings, integer state and the food grid -- because pair sums are integers, feeding
ranks are by global id inside one food cell, and draws are keyed by global id."""
import os
import subprocess
import sys

import numpy as np
import pytest

from fish_abm_b200 import School, synth
from fish_abm_b200.slab import LocalSlabs, partition

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIELDS = ("coords", "dir", "speed", "speed_factor", "lag", "sample_rate", "food_intake")


def whole_school(n, w, d, q, seed, steps):
    with School(n=n, w=w, h=w, seed=seed) as S:
        S.set_state(**d)
        S.set_quality(q)
        zc0 = S.zone_counts()
        S.step(steps)
        return zc0, S.get_state(), S.get_quality(), S.zone_counts(), S.social()[0]


def assert_same(a, b):
    for k in FIELDS:
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("n,w,steps,speed", [(20_000, 6400, 12, 5.0), (3_000, 1000 * 2, 25, 20.0), (400, 600, 30, 30.0)])
def test_single_slab_with_halo_columns_equals_whole_school(q100, n, w, steps, speed):
    """force_slab: the slab is its own neighbour on both sides; every fish crossing x=0 migrates through a message"""
    d = synth.uniform_state(n, w, w, seed=41, speed=speed)
    q = synth.tile_quality(q100, w, w)
    zc0, g, qa, zc1, turn = whole_school(n, w, d, q, 77, steps)
    with School(n=n, w=w, h=w, seed=77, force_slab=True) as S:
        S.set_state(**d)
        S.set_quality(q)
        assert np.array_equal(S.zone_counts(), zc0)
        S.step(steps)
        assert_same(S.get_state(), g)
        assert np.array_equal(S.get_quality(), qa)
        assert np.array_equal(S.zone_counts(), zc1)
        assert np.array_equal(S.social()[0], turn)
        info = S.slab_info()
        assert info["owned"] == n and info["held"] > n


@pytest.mark.parametrize("nslabs", [2, 3, 4, 7])
def test_local_slabs_equal_whole_school(q100, nslabs):
    n, w, steps = 30_000, 7 * 200 * 2, 15  # 14 cell columns: uneven splits for 3 and 4 slabs, 2 columns each for 7
    d = synth.uniform_state(n, w, w, seed=43, speed=12.0)
    q = synth.tile_quality(q100, w, w)
    zc0, g, qa, zc1, turn = whole_school(n, w, d, q, 78, steps)
    with LocalSlabs(nslabs, n=n, w=w, h=w, seed=78) as P:
        P.set_state(**d)
        P.set_quality(q)
        cols = [(s.slab_info()["col0"], s.slab_info()["col1"]) for s in P.S]
        assert cols == partition(w // 200, nslabs)
        assert sum(s.slab_info()["owned"] for s in P.S) == n
        assert np.array_equal(P.zone_counts(), zc0)
        P.step(steps)
        assert sum(s.slab_info()["owned"] for s in P.S) == n  # migration conserves fish
        assert_same(P.get_state(), g)
        assert np.array_equal(P.get_quality(), qa)
        assert np.array_equal(P.zone_counts(), zc1)
        assert np.array_equal(P.social_turn(), turn)
        assert P.sum_array() == int(qa.sum())


def gpus_for(nslabs):
    """one GPU per slab, or skip: the peer-memory transport waits on the device for the neighbour's kernels, and kernels of
    ONE GPU are not guaranteed to run at the same time (B200_PROFILING.md; fishabm_slab_attach refuses shared GPUs)"""
    import torch
    if torch.cuda.device_count() < nslabs:
        pytest.skip(f"peer-memory transport needs one GPU per slab ({nslabs})")
    return list(range(nslabs))


@pytest.mark.parametrize("nslabs", [2, 3, 5])
def test_peer_memory_transport_in_one_process(q100, nslabs, monkeypatch):
    """the slabs pack straight into each other's inboxes and synchronise with device-side flags (one process, one GPU per slab)"""
    devices = gpus_for(nslabs)
    monkeypatch.setenv("FISHABM_SLAB_TIMEOUT_MS", "4000")
    n, w, steps = 30_000, 7 * 200 * 2, 21
    d = synth.uniform_state(n, w, w, seed=43, speed=12.0)
    q = synth.tile_quality(q100, w, w)
    zc0, g, qa, zc1, turn = whole_school(n, w, d, q, 78, steps)
    with LocalSlabs(nslabs, devices=devices, transport="p2p", n=n, w=w, h=w, seed=78) as P:
        assert P.transport == "p2p"
        P.set_state(**d)
        P.set_quality(q)
        P.step(steps)
        assert_same(P.get_state(), g)
        assert np.array_equal(P.get_quality(), qa)
        assert np.array_equal(P.zone_counts(), zc1)
        assert np.array_equal(P.social_turn(), turn)


def test_slab_local_state_round_trip_and_halo_refresh(q100, monkeypatch):
    """set_state_owned / get_state_owned: every slab is given only its own fish, the halos come from one exchange; the
    run must continue exactly like the whole school"""
    from fish_abm_b200 import FishAbmError
    devices = gpus_for(3)
    monkeypatch.setenv("FISHABM_SLAB_TIMEOUT_MS", "4000")
    n, w = 30_000, 7 * 200 * 2
    d = synth.uniform_state(n, w, w, seed=49, speed=12.0)
    q = synth.tile_quality(q100, w, w)
    _, g10, _, _, _ = whole_school(n, w, d, q, 80, 10)
    _, g20, qa20, zc20, turn20 = whole_school(n, w, d, q, 80, 20)
    with LocalSlabs(3, devices=devices, transport="p2p", n=n, w=w, h=w, seed=80) as P:
        P.set_state(**d)
        P.set_quality(q)
        P.step(10)
        parts = [s.get_state_owned() for s in P.S]
        assert sum(len(p["ids"]) for p in parts) == n
        allids = np.concatenate([p["ids"] for p in parts])
        assert np.array_equal(np.sort(allids), np.arange(n))
        for k in FIELDS:  # compact + ids == the whole-school state after 10 steps
            full = np.concatenate([p[k] for p in parts])
            assert np.array_equal(full[np.argsort(allids)], g10[k]), k
        # hand every slab its own fish back (shuffled), halos are rebuilt by the exchange, and go on
        rng = np.random.default_rng(0)
        for s, p in zip(P.S, parts):
            perm = rng.permutation(len(p["ids"]))
            s.set_state_owned(**{k: v[perm] for k, v in p.items()}, wait=False)
        for s in P.S:
            s.wait()
            s.step_index = 10
        P.step(10)
        assert_same(P.get_state(), g20)
        assert np.array_equal(P.get_quality(), qa20)
        assert np.array_equal(P.zone_counts(), zc20)
        assert np.array_equal(P.social_turn(), turn20)
        # a fish outside the slab's columns is refused
        wrong = {k: v.copy() for k, v in parts[1].items()}
        P.S[0].set_state_owned(**{k: v[:1] for k, v in wrong.items()}, wait=False)
        P.S[1].set_state_owned(**parts[1], wait=False)
        P.S[2].set_state_owned(**parts[2], wait=False)
        with pytest.raises(FishAbmError) as e:
            P.S[0].wait()
        assert "outside this slab's columns" in str(e.value)
        P.S[1].wait(); P.S[2].wait()


def test_peer_memory_timeout_is_reported_not_hung(q100, monkeypatch):
    """a neighbour that never steps: the wait kernel gives up after the timeout and the call reports it"""
    from fish_abm_b200 import FishAbmError
    devices = gpus_for(2)
    monkeypatch.setenv("FISHABM_SLAB_TIMEOUT_MS", "300")
    n, w = 5_000, 2000
    d = synth.uniform_state(n, w, w, seed=48)
    with LocalSlabs(2, devices=devices, transport="p2p", n=n, w=w, h=w, seed=3) as P:
        P.set_state(**d)
        P.set_quality(q100)
        P.S[0].step_enqueue(1)
        with pytest.raises(FishAbmError) as e:
            P.S[0].wait()
        assert e.value.code == -5 and "did not arrive" in str(e.value)


def test_slab_queries_and_feed_match_whole_school(q100):
    """in_zone with arbitrary fov / r, get_speed and feed() through slab contexts (each serves the ids it owns)"""
    n, w = 12_000, 2400
    d = synth.uniform_state(n, w, w, seed=52)
    q = synth.tile_quality(q100, w, w)
    ids = np.random.default_rng(7).choice(n, 300, replace=False).astype(np.int32)
    with School(n=n, w=w, h=w, seed=81) as S:
        S.set_state(**d); S.set_quality(q)
        want_zone, want_speed = S.in_zone(200, 150.0), S.get_speed()
        S.feed(ids)
        want_state, want_q = S.get_state(), S.get_quality()
    with LocalSlabs(3, n=n, w=w, h=w, seed=81) as P:
        P.set_state(**d); P.set_quality(q)
        assert np.array_equal(P.in_zone(200, 150.0), want_zone)
        from fish_abm_b200.slab import merge_owned
        assert np.array_equal(merge_owned([s.get_speed() for s in P.S], float("nan")), want_speed)
        for s in P.S:
            s.feed(ids)   # every slab is given the whole list and feeds the fish it owns, in list order
        got = P.get_state()
        for k in FIELDS:
            assert np.array_equal(got[k], want_state[k]), k
        assert np.array_equal(P.get_quality(), want_q)


def test_one_column_slabs_are_rejected(q100):
    """a slab needs two owned columns (a migrant into a one-column slab would also be a ghost for the slab beyond)"""
    from fish_abm_b200 import FishAbmError
    with pytest.raises(FishAbmError) as e:
        School(n=2000, w=1000, h=1000, rank=0, nranks=3)
    assert e.value.code == -1 and "2 of the 5 cell columns" in str(e.value)
    School(n=2000, w=1000, h=1000, rank=1, nranks=2).close()


def test_slab_capacity_overflow_is_reported(q100):
    from fish_abm_b200 import FishAbmError
    n, w = 20_000, 2000
    d = synth.uniform_state(n, w, w, seed=45)
    with School(n=n, w=w, h=w, seed=1, force_slab=True, halo_capacity=16) as S:
        S.set_state(**d)
        S.set_quality(synth.tile_quality(q100, w, w))
        with pytest.raises(FishAbmError) as e:
            S.step(1)
        assert e.value.code == -4 and "capacity" in str(e.value)
    with pytest.raises(FishAbmError):
        with School(n=n, w=w, h=w, seed=1, force_slab=True, capacity=1000) as S:
            S.set_state(**d)


def test_slab_rejects_fast_fish(q100):
    from fish_abm_b200 import FishAbmError
    d = synth.uniform_state(100, 2000, 2000, seed=46, speed=70.0)
    with School(n=100, w=2000, h=2000, force_slab=True) as S:
        with pytest.raises(FishAbmError):
            S.set_state(**d)


def test_unconnected_multi_rank_step_fails_loudly(q100):
    from fish_abm_b200 import FishAbmError
    d = synth.uniform_state(1000, 2000, 2000, seed=47)
    with School(n=1000, w=2000, h=2000, rank=0, nranks=2) as S:
        S.set_state(**d)
        S.set_quality(q100)
        with pytest.raises(FishAbmError) as e:
            S.step(1)
        assert e.value.code == -3


@pytest.mark.parametrize("transport", ["p2p", "nccl"])
@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_multi_gpu_slabs_equal_whole_school(q100, tmp_path, nranks, transport, monkeypatch):
    """one process per GPU; the exchange runs inside the library: peer-memory inboxes over NVLink (CUDA IPC), or NCCL
    send/recv.  Needs that many GPUs on the box."""
    import torch
    if torch.cuda.device_count() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    monkeypatch.setenv("FISHABM_SLAB_TRANSPORT", transport)
    monkeypatch.setenv("FISHABM_SLAB_TIMEOUT_MS", "20000")
    out = tmp_path / "res.npz"
    port = 29650 + nranks + (10 if transport == "nccl" else 0)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nranks}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "slab_worker.py"), str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = np.load(out)
    n, w, steps, seed = int(res["n"]), int(res["w"]), int(res["steps"]), int(res["seed"])
    d = synth.uniform_state(n, w, w, seed=51, speed=10.0)
    q = synth.tile_quality(q100, w, w)
    _, g, qa, zc1, _ = whole_school(n, w, d, q, seed, steps)
    for k in FIELDS:
        assert np.array_equal(res[k], g[k]), k
    assert np.array_equal(res["quality"], qa)
    assert np.array_equal(res["zone"], zc1)


def test_lone_slab_local_state_round_trip(q100):
    """set_state_owned / get_state_owned on ONE slab that is its own neighbour (force_slab): runs on a single GPU and keeps
    the slab-local state path covered there; the run continues exactly like the whole school"""
    n, w = 20_000, 2400
    d = synth.uniform_state(n, w, w, seed=61, speed=12.0)
    q = synth.tile_quality(q100, w, w)
    _, g20, qa20, zc20, _ = whole_school(n, w, d, q, 82, 20)
    with School(n=n, w=w, h=w, seed=82, force_slab=True) as S:
        S.set_state(**d); S.set_quality(q)
        S.step(10)
        own = S.get_state_owned()
        assert np.array_equal(np.sort(own["ids"]), np.arange(n))
        perm = np.random.default_rng(1).permutation(n)
        S.set_state_owned(**{k: v[perm] for k, v in own.items()})
        S.step_index = 10
        S.step(10)
        assert_same(S.get_state(), g20)
        assert np.array_equal(S.get_quality(), qa20)
        assert np.array_equal(S.zone_counts(), zc20)
        bad = {k: v[:4].copy() for k, v in own.items()}
        bad["ids"][2] = n + 5        # ids index device and host arrays: anything outside [0, n) is refused up front
        from fish_abm_b200 import FishAbmError
        with pytest.raises(FishAbmError) as e:
            S.set_state_owned(**bad)
        assert e.value.code == -1 and "outside [0" in str(e.value)


def test_peer_memory_attach_refuses_a_shared_gpu(q100):
    """two slabs on ONE GPU cannot wait for each other on the device: fishabm_slab_attach says so (and LocalSlabs falls back
    to the caller-driven exchange)"""
    from fish_abm_b200 import FishAbmError
    A = School(n=4000, w=2000, h=2000, rank=0, nranks=2, device=0)
    B = School(n=4000, w=2000, h=2000, rank=1, nranks=2, device=0)
    try:
        with pytest.raises(FishAbmError) as e:
            A.slab_attach(B.slab_endpoint(), B.slab_endpoint())
        assert e.value.code == -1 and "same GPU" in str(e.value)
    finally:
        A.close(); B.close()
    with LocalSlabs(2, transport="p2p", n=4000, w=2000, h=2000, seed=1) as P:
        assert P.transport == "hooks"
