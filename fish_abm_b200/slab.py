"""One school split into x-slabs across GPUs (BASELINE.json north_star).  No reference counterpart:
fish_mod.cpp is single-threaded.  The decomposition itself lives in libfishabm.so (include/fishabm.h, "slab
decomposition"); this module is the host plumbing around it:

  partition / ring   the column split and ring neighbours the library uses (restated for host-side checks)
  broadcast_unique_id, connect   NCCL bootstrap over an existing torch.distributed group (any backend)
  DistributedSchool  one rank's slab in a one-process-per-GPU job; the exchange runs inside the library (NCCL)
  LocalSlabs         several slabs in ONE process (one or more GPUs), the exchange driven from here through the
                     caller-driven hooks: used to test the decomposition on a single GPU
  merge_owned        assemble per-slab outputs into whole-school arrays
"""
from __future__ import annotations

import ctypes as C
import json
import os
import time

import numpy as np

from . import _lib
from .school import School, slab_unique_id


def partition(nx: int, nranks: int):
    """owned cell columns [col0, col1) of every rank: col_r = floor(r*nx/nranks)"""
    return [(r * nx // nranks, (r + 1) * nx // nranks) for r in range(nranks)]


def ring(rank: int, nranks: int):
    """(left, right) neighbours in the periodic world"""
    return (rank - 1) % nranks, (rank + 1) % nranks


def merge_owned(parts, not_owned):
    """parts: per-slab arrays of identical shape in which entries a slab does not own hold `not_owned` (NaN allowed).
    Every entry must be owned by exactly one slab."""
    out = np.array(parts[0], copy=True)
    seen = np.zeros(out.shape[0], np.int32)
    for p in parts:
        own = ~np.isnan(p) if (isinstance(not_owned, float) and np.isnan(not_owned)) else (p != not_owned)
        own = own.reshape(own.shape[0], -1).any(axis=1) if own.ndim > 1 else own
        out[own] = p[own]
        seen += own
    if not np.all(seen == 1):
        raise ValueError(f"ownership is not a partition: {np.count_nonzero(seen == 0)} unowned, {np.count_nonzero(seen > 1)} multiply owned")
    return out


def merge_states(states):
    """whole-school state dict from the per-slab get_state() dicts"""
    masks = [School.owned_mask(s) for s in states]
    seen = np.sum(masks, axis=0)
    if not np.all(seen == 1):
        raise ValueError(f"ownership is not a partition: {np.count_nonzero(seen == 0)} unowned, {np.count_nonzero(seen > 1)} multiply owned")
    out = {k: np.array(v, copy=True) for k, v in states[0].items()}
    for s, m in zip(states[1:], masks[1:]):
        for k in out:
            out[k][m] = s[k][m]
    return out


def merge_quality(grids):
    out = np.array(grids[0], copy=True)
    for g in grids[1:]:
        m = g >= 0
        if np.any(m & (out >= 0)):
            raise ValueError("food columns owned twice")
        out[m] = g[m]
    if np.any(out < 0):
        raise ValueError("food columns unowned")
    return out


# ---- NCCL bootstrap over torch.distributed ------------------------------------------------------------------------------
def broadcast_unique_id(dist, rank: int, device=None) -> bytes:
    """rank 0 creates the NCCL unique id (through the library) and every rank of the torch.distributed group gets it"""
    import torch
    t = torch.zeros(_lib.UNIQUE_ID_BYTES, dtype=torch.uint8)
    if rank == 0:
        t = torch.frombuffer(bytearray(slab_unique_id()), dtype=torch.uint8).clone()
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src=0)
    return bytes(t.cpu().numpy().tobytes())


def gather_endpoints(dist, endpoint: bytes, device=None):
    """every rank's inbox endpoint, in rank order"""
    import torch
    n = dist.get_world_size()
    mine = torch.frombuffer(bytearray(endpoint), dtype=torch.uint8).clone()
    if device is not None:
        mine = mine.to(device)
    parts = [torch.zeros_like(mine) for _ in range(n)]
    dist.all_gather(parts, mine)
    return [bytes(p.cpu().numpy().tobytes()) for p in parts]


class DistributedSchool(School):
    """This rank's slab.  `dist` is an initialised torch.distributed module/group used only for the bootstrap.
    transport: "p2p" = peer-memory inboxes over NVLink (default), "nccl" = ncclSend/ncclRecv inside the library
    (FISHABM_SLAB_TRANSPORT overrides)."""

    def __init__(self, dist, rank: int, nranks: int, device: int, transport: str | None = None, **kw):
        super().__init__(rank=rank, nranks=nranks, device=device, **kw)
        self.transport = transport or os.environ.get("FISHABM_SLAB_TRANSPORT", "p2p")
        if nranks > 1:
            import torch
            dev = torch.device("cuda", device) if dist.get_backend() == "nccl" else None
            if self.transport != "nccl":
                # peer memory needs CUDA IPC + peer access between ring neighbours; if ANY rank cannot attach, every
                # rank falls back to NCCL together
                ok = 1
                try:
                    eps = gather_endpoints(dist, self.slab_endpoint(), dev)
                    left, right = ring(rank, nranks)
                    self.slab_attach(eps[left], eps[right])
                except Exception as e:  # noqa: BLE001
                    ok = 0
                    self.attach_error = str(e)
                flag = torch.tensor([ok], dtype=torch.int32, device=dev if dev is not None else "cpu")
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)  # also the barrier: nobody stores into an inbox before all attached
                if int(flag.item()) == 0:
                    if ok:  # attached here but not everywhere: a fresh context for the NCCL path
                        self.close()
                        School.__init__(self, rank=rank, nranks=nranks, device=device, **kw)
                    self.transport = "nccl"
            if self.transport == "nccl":
                self.slab_connect(broadcast_unique_id(dist, rank, dev))


# ---- several slabs in one process -----------------------------------------------------------------------------------------
class LocalSlabs:
    """nslabs contexts in this process; step() drives the exchange through the caller-driven hooks
    (fishabm_slab_begin_step / fishabm_slab_buffers / fishabm_slab_end_step)."""

    def __init__(self, nslabs: int, devices=None, transport: str = "hooks", **kw):
        """transport "hooks": this class copies the messages (caller-driven exchange); "p2p": the slabs are attached
        to each other's inboxes and exchange by themselves (peer-memory transport, here inside one process)"""
        self.k = nslabs
        devices = devices or [0] * nslabs
        if transport == "p2p" and len(set(devices)) < nslabs:
            # kernels of one GPU cannot wait for each other (include/fishabm.h: fishabm_slab_attach): slabs that share a
            # GPU are driven through the hooks
            transport = "hooks"
        self.transport = transport
        self.S = [School(rank=r, nranks=nslabs, device=devices[r], force_slab=True, **kw) for r in range(nslabs)]
        self.n, self.w, self.h = self.S[0].n, self.S[0].w, self.S[0].h
        if transport == "p2p":
            assert nslabs > 1
            eps = [s.slab_endpoint() for s in self.S]
            for r, s in enumerate(self.S):
                left, right = ring(r, nslabs)
                s.slab_attach(eps[left], eps[right])

    def close(self):
        for s in self.S:
            s.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_state(self, **d):
        for s in self.S:
            s.set_state(**d)

    def set_quality(self, q):
        for s in self.S:
            s.set_quality(q)

    def set_step_index(self, v: int):
        for s in self.S:
            s.step_index = v

    def step(self, nsteps: int = 1):
        L = self.S[0].L
        if self.transport == "p2p":
            # every slab waits (on the device) for its neighbours' messages: enqueue all of them before waiting for
            # any, in chunks short enough for the launch queues
            done = 0
            while done < nsteps:
                k = min(16, nsteps - done)
                for s in self.S:
                    s.step_enqueue(k)
                for s in self.S:
                    s.wait()
                done += k
            return
        for _ in range(nsteps):
            for s in self.S:
                s.slab_begin_step()
            bufs = [s.slab_buffers() for s in self.S]
            for r in range(self.k):
                (out_l, out_r, _, _), nbytes = bufs[r]
                left, right = ring(r, self.k)
                # what leaves through my left edge arrives at my left neighbour's right side, and vice versa
                assert L.fishabm_slab_copy(bufs[left][0][3], out_l, nbytes) == 0
                assert L.fishabm_slab_copy(bufs[right][0][2], out_r, nbytes) == 0
            for s in self.S:
                s.slab_end_step()

    def get_state(self):
        return merge_states([s.get_state() for s in self.S])

    def get_quality(self):
        return merge_quality([s.get_quality() for s in self.S])

    def sum_array(self) -> int:
        return sum(s.sum_array() for s in self.S)

    def zone_counts(self):
        return merge_owned([s.zone_counts() for s in self.S], -1)

    def social_turn(self):
        return merge_owned([s.social()[0] for s in self.S], float("nan"))

    def in_zone(self, fov, r):
        return merge_owned([s.in_zone(fov, r) for s in self.S], -1)


# ---- the multi-GPU benchmark leg (bench.py --gpus N, N > 1) ----------------------------------------------------------------
def compare_with_whole_school(S, g: dict, gq: np.ndarray) -> dict:
    """this rank's slab against a whole-school run of the same steps: every field of every fish the slab owns and every
    food cell of its columns must be bit-identical (id-ordered feed, fish_mod.cpp:541,563-574, and migration are what can
    break across ranks).  g / gq: get_state() / get_quality() of the whole school."""
    own = S.get_state_owned()
    ids = own["ids"]
    bad = np.zeros(len(ids), bool)
    per_field = {}
    for k in School.FIELDS:
        a, b = own[k], g[k][ids]
        neq = (a != b).reshape(len(ids), -1).any(axis=1)
        per_field[k] = int(neq.sum())
        bad |= neq
    sq = S.get_quality()
    m = sq >= 0
    return {"owned": int(len(ids)), "fish_mismatches": int(bad.sum()), "unique_ids": int(len(np.unique(ids)) == len(ids)),
            "food_cells": int(m.sum()), "food_mismatches": int((sq[m] != gq[m]).sum()), "per_field": per_field}


def run_slab_bench(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist
    from . import synth
    from bench import (N_C4, ClockSampler, measured_peaks, flops_alg, FP32_LANES_PER_SM, ROOT, workload_string, whole_school_leg,
                       build_info, ensemble_key)

    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n = args.fish or N_C4
    w = synth.world_size_for(n)
    q100 = np.loadtxt(os.path.join(ROOT, "tests", "golden", "quality.csv"), delimiter=",", dtype=np.int32)
    d = synth.uniform_state(n, w, w, seed=0x5EED)   # every rank composes the same whole-school state ...
    q = synth.tile_quality(q100, w, w)
    P = 5                                            # steps of the phase-timed run
    T = args.warmup + args.steps + 1 + P             # steps every arm ends at (the parity check compares states at T)

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    def allmax(*v):
        t = torch.tensor(v, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def allsum(*v):
        t = torch.tensor(v, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(x) for x in t]

    def timed_arm(transport: str, detailed: bool):
        """warm-up, K timed steps (device events, max over ranks), then up to step T; returns the open slab + numbers"""
        S = DistributedSchool(dist, rank, world, local_rank, transport=transport, n=n, w=w, h=w, seed=0x5EED)
        S.set_state(**d)                             # ... and keeps its own columns
        S.set_quality(q)
        info = S.slab_info()
        for _ in range(args.warmup):
            S.step(1)
        S.flush()
        barrier()
        launches0 = S.launch_count()
        # K steps in one library call per rank: neighbour pass, update+pack, exchange, ingest, sort, all on the library's
        # stream with no host synchronisation in between; the trailing sample() is part of the K steps
        t0 = time.perf_counter()
        S.step(args.steps)
        ms = S.last_step_ms()
        nb_ms, nb_l = S.last_neighbour_ms()
        S.flush()
        ms += S.last_step_ms()
        a, b = S.last_neighbour_ms(); nb_ms += a; nb_l += b
        barrier()
        wall = time.perf_counter() - t0
        launches = S.launch_count() - launches0
        ms_max, wall_max = allmax(ms, wall * 1e3)
        res = {"transport": S.transport, "value": n * args.steps / (ms_max * 1e-3), "ms_per_step": ms_max / args.steps,
               "wall_ms_per_step_max": wall_max / args.steps, "launches": launches, "nb_ms": nb_ms, "nb_l": nb_l, "info": info, "ms_max": ms_max}
        if detailed:
            # pair counters of one pass for the roofline line (rank-local work, summed over ranks)
            S.enable_stats(True)
            S.step(1)
            res["stats"] = S.last_pass_stats()
            S.flush()
            S.enable_stats(False)
            # where a step goes on every rank: device time by phase over P steps
            S.enable_phase_timing(True)
            barrier()
            S.step(P)
            ph = S.last_phase_ms()
            S.enable_phase_timing(False)
            S.flush()
            mine = torch.tensor([ph["neighbour"] / P, ph["update"] / P, ph["exchange"] / P, ph["sort"] / P, float(S.slab_info()["owned"])],
                                dtype=torch.float64, device="cuda")
            parts = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
            tab = np.array([p.cpu().numpy() for p in parts])
            res["per_rank"] = {"neighbour_ms": tab[:, 0].round(4).tolist(), "update_ms": tab[:, 1].round(4).tolist(),
                               "exchange_wait_ms": tab[:, 2].round(4).tolist(), "sort_ms": tab[:, 3].round(4).tolist(),
                               "owned_fish": tab[:, 4].astype(int).tolist(),
                               "max": dict(zip(("neighbour_ms", "update_ms", "exchange_wait_ms", "sort_ms"), tab[:, :4].max(axis=0).round(4).tolist())),
                               "mean": dict(zip(("neighbour_ms", "update_ms", "exchange_wait_ms", "sort_ms"), tab[:, :4].mean(axis=0).round(4).tolist())),
                               "how": f"fishabm_last_phase_ms over {P} steps after the timed run (CUDA events on each rank's stream)"}
        else:
            S.step(1 + P)
            S.flush()
        assert S.step_index == T, (S.step_index, T)
        return S, res

    clocks = ClockSampler(local_rank)   # sampled from the warm-up through the end of the timed steps
    clocks.start()
    S, main = timed_arm(None, detailed=True)
    clk = clocks.stop()
    info, st = main["info"], main["stats"]
    c = torch.tensor([st["candidates"], st["interacting"], st["focal"], main["launches"], main["nb_ms"], main["nb_l"], info["owned"]],
                     dtype=torch.float64, device="cuda")
    cs = c.clone(); dist.all_reduce(cs, op=dist.ReduceOp.SUM)
    cm = c.clone(); dist.all_reduce(cm, op=dist.ReduceOp.MAX)

    # ---- parity on real hardware: every rank runs the WHOLE school for the same T steps on its own GPU (2.2 GB) and
    # compares the fish and food columns its slab owns bit for bit; rank 0's timing of that run is the 1-GPU base of the
    # scaling curve
    one_gpu, G = whole_school_leg(n, local_rank, args.warmup, args.steps, d=d, q=q, keep=True)
    G.step(1 + P)
    G.flush()
    assert G.step_index == T
    g, gq = G.get_state(), G.get_quality()
    G.close()

    def parity_of(Sx):
        r = compare_with_whole_school(Sx, g, gq)
        tot = allsum(r["owned"], r["fish_mismatches"], r["food_cells"], r["food_mismatches"], 1 - r["unique_ids"])
        return {"against": f"a whole-school run of the same {T} steps on one GPU (every rank, its own copy)", "steps": T,
                "fish_compared": int(tot[0]), "fish_mismatches": int(tot[1]), "fields": list(School.FIELDS),
                "food_cells_compared": int(tot[2]), "food_mismatches": int(tot[3]), "ranks_with_duplicate_ids": int(tot[4]),
                "all_fish_accounted_for": bool(int(tot[0]) == n)}

    parity = parity_of(S)

    # end to end through the C ABI, host buffers (pinned) in and out every step: each rank uploads ITS fish
    # (fishabm_set_state_owned: ids + state, then one halo exchange), steps, downloads its fish (which migration has
    # changed) and its partial food sum.  No rank ever holds the whole school.
    e2e_steps = max(3, min(args.steps, 10))
    cap = int(info["capacity"])
    shapes = dict(ids=((cap,), np.int32), coords=((cap, 2), np.float64), dir=((cap,), np.float64), speed=((cap,), np.float64),
                  speed_factor=((cap,), np.float64), lag=((cap,), np.int32), sample_rate=((cap,), np.int32), food_intake=((cap,), np.int32))
    bufs = [{k: torch.empty(shp, dtype=torch.from_numpy(np.empty(0, dt)).dtype).pin_memory().numpy() for k, (shp, dt) in shapes.items()} for _ in range(2)]
    own = S.get_state_owned(out=bufs[0])
    S.set_state_owned(**own)
    S.step(1)
    own = S.get_state_owned(out=bufs[1])
    moved_bytes = 0
    barrier()
    t0 = time.perf_counter()
    for it in range(e2e_steps):
        S.set_state_owned(**own)
        S.step(1)
        moved_bytes += 56 * len(own["ids"])
        own = S.get_state_owned(out=bufs[it & 1])
        S.sum_array()
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_max, = allmax(e2e_s)
    h2d = d2h = int(allsum(moved_bytes / e2e_steps)[0])
    transport_main = S.transport
    S.close()
    del bufs

    # ---- the other transport (north_star names NCCL send/recv; the default is the peer-memory inboxes): same steps, same check
    other = "nccl" if transport_main == "p2p" else "p2p"
    transports = {transport_main: {"value": main["value"], "ms_per_step": main["ms_per_step"], "parity": parity}}
    if not getattr(args, "one_transport", False):
        try:
            S2, r2 = timed_arm(other, detailed=False)
            transports[r2["transport"]] = {"value": r2["value"], "ms_per_step": r2["ms_per_step"], "parity": parity_of(S2)}
            S2.close()
        except Exception as e:  # noqa: BLE001
            transports[other] = {"error": str(e)[:300]}
    per_rank = max(1, args.schools // world)   # replicate ids rank * per_rank ...: part of the Philox key, no two ranks run the same school
    ens = ensemble_key(local_rank, per_rank, replicate0=rank * per_rank) if not getattr(args, "no_ensemble", False) else None
    if ens is not None:   # replicates are independent: every rank ran schools/world of them, slowest rank sets the time
        ems, = allmax(ens["ms_per_step"])
        eat, = allsum(ens["food_eaten"])
        tot = ens["schools"] * world
        ens = dict(ens, schools=tot, ms_per_step=ems, value=tot * ens["n_fish"] / (ems * 1e-3), school_steps_per_sec=tot / (ems * 1e-3),
                   food_eaten=int(eat), scaling="weak: replicates split across the GPUs, no communication")
    if rank == 0:
        peaks, how = measured_peaks()
        sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
        fp32_peak = sm_count * FP32_LANES_PER_SM * 2 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
        per_launch_ms = float(cm[4]) / max(float(cm[5]), 1.0)           # slowest rank's average neighbour launch
        flops = flops_alg(float(cs[0]), float(cs[1])) / world            # per rank, per launch
        achieved = flops / (per_launch_ms * 1e-3) / 1e12
        value = main["value"]
        line = {"metric": "agent_steps_per_sec", "value": value, "unit": "agent-steps/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "scaling_note": "strong scaling of C4 (16.8M fish) over the GPUs; its 1-GPU base is c4_one_gpu, measured in this same run",
                "config": {"workload": workload_string(n, w), "n_fish": n, "world": w, "parallelism": f"slab{world}",
                           "decomposition": f"{world} x-slabs with 200-unit halo columns + migration", "transport": transport_main,
                           "l2": "per-rank state larger than L2" if n // world * 76 > (126 << 20) else "state fits L2, not flushed",
                           "timing": "CUDA events on each rank's library stream around fishabm_step(K) + the trailing sample pass, max over ranks",
                           "message_bytes_per_neighbour_per_step": info["message_bytes"], "wall_ms_per_step_max": main["wall_ms_per_step_max"]},
                "clocks": clk, "gpu_launches": int(cs[3]),
                "e2e": {"value": n * e2e_steps / e2e_max, "unit": "agent-steps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h) + 8 * world,
                        "steps": e2e_steps, "api": "per rank: fishabm_set_state_owned(its fish, pinned host arrays) + halo exchange + fishabm_step(1) + fishabm_get_state_owned + fishabm_sum_quality"},
                "roofline": {"bound": "fp32", "kernel": "neighbour_kernel_v4", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                             "frac": achieved / fp32_peak, "traffic": None, "per": "rank (slowest rank's launch time, mean rank's pairs)",
                             "peak_source": f"SMs({sm_count}) x 128 lanes x 2 x sm_max_mhz ({how} clock)", "flops_per_launch": flops,
                             "launch_ms": per_launch_ms, "share_of_step": float(cm[4]) / main["ms_max"]},
                "per_rank": main["per_rank"],
                "c4_one_gpu": one_gpu, "efficiency_vs_c4_one_gpu": value / (world * one_gpu["value"]),
                "parity_spot_check": parity, "transports": transports, "ensemble": ens,
                "cpu_baseline": None, "build": build_info()}
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()
