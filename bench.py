#!/usr/bin/env python
"""Benchmark of the per-step agent update (BASELINE.json: agent-steps/s at 1M and 16M fish on 1/2/4/8 B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--fish F]

One "step" = one reference step (move() then sample(), fish_mod.cpp:130-131) of the whole school.
  N=1  : workload C3 = 1,000,000 fish, W=H=44,800, quality.csv tiled.
  N>1  : workload C4 = 16,777,216 fish, W=H=183,200, split into N spatial slabs (one rank per GPU).
Prints ONE JSON line (rank 0).  `value` is device-timed with the state resident in HBM; `e2e` goes through the
C ABI with host buffers in the reference's layout, host<->device copies inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FP32_LANES_PER_SM = 128
N_C3 = 1_000_000
N_C4 = 16_777_216


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md's clocks line), in-process
    through NVML (nvidia_ml_py) every 2 ms -- the timed region is tens of milliseconds, too short for an `nvidia-smi -lms`
    child process to even start.  Falls back to one nvidia-smi query if NVML is unavailable."""
    BAD = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}

    def __init__(self, device_index: int):
        self.idx = device_index
        self.samples, self.reasons = [], set()
        self.stop_flag = threading.Event()
        self.thread = None
        self.h = None
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[device_index]) if visible and visible.split(",")[device_index].strip().isdigit() else device_index
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def _loop(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for name, bit in self.BAD.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                break
            time.sleep(0.002)

    def start(self):
        if self.h is not None:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()

    def stop(self) -> dict:
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=1)
            if self.samples:
                return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                        "samples": len(self.samples), "how": "NVML, 2 ms, from the warm-up to the end of the timed steps"}
        try:  # fallback: one query after the fact
            q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
            out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.idx)],
                                 capture_output=True, text=True, timeout=20).stdout.strip().split(",")
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": [n for n, v in zip(names, out[2:6]) if v.strip().lower().startswith("active")],
                    "samples": 1, "how": "nvidia-smi, one query after the timed region (NVML unavailable)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"], "samples": 0}


def ncu_capture(kernel_prefix: str):
    """the committed `ncu --set full` capture of the dominant kernel (profiles/neighbour_ncu.json, written by
    tools/ncu_to_json.py from the capture named inside): DRAM traffic per launch and pipe / issue utilisation"""
    p = os.path.join(ROOT, "profiles", "neighbour_ncu.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        d = json.load(f)
    for k in d.get("launches", []):
        if k.get("kernel", "").startswith(kernel_prefix):
            return dict(k, source="profiles/neighbour_ncu.json <- " + d.get("source", "?"))
    return None


def workload_string(n: int, w: int) -> str:
    """config.workload, identical in both arms (ours / reference) for the same school"""
    tag = {N_C3: "C3: ", N_C4: "C4: "}.get(n, "")
    return f"{tag}N={n} school, W=H={w} periodic, cell-list r=200, rho=5e-4, quality.csv tiled, 60% fish lagged at t=0"


def stratified_ids(lag: np.ndarray, k: int, seed: int) -> np.ndarray:
    """k focal fish whose moving (lag == 0, 8 all-pairs passes) : lagged (3 passes) mix equals the school's: a plain random
    sample of a few dozen fish makes the extrapolated agent-steps/s swing by +-50 % with the mix alone"""
    rng = np.random.default_rng(seed)
    moving, lagged = np.nonzero(lag == 0)[0], np.nonzero(lag != 0)[0]
    km = int(round(k * len(moving) / len(lag)))
    km = min(max(km, 1 if len(moving) else 0), len(moving))
    ids = np.concatenate([rng.choice(moving, km, replace=False), rng.choice(lagged, min(k - km, len(lagged)), replace=False)])
    return rng.permutation(ids).astype(np.int32)


def build_info() -> dict:
    from fish_abm_b200 import build as b
    return {"libfishabm": b.LAST["decision"], "nvcc": b.LAST["nvcc"] or b.nvcc_version()}


def flops_alg(cand: float, inter: float) -> float:
    """13 flops per candidate pair + 20 per interacting pair"""
    return 13.0 * cand + 20.0 * inter


# ------------------------------------------------------------------------------------------------------------------
def reference_arm(args, rank: int, world: int):
    """The reference's own CPU implementation of the path on this box's host cores: the compiled, unmodified
    fish_mod.cpp (oracle/_ref) when present, else the C restatement.  Each step = a bounded sample of the
    workload: the full per-fish work of move()+sample() for `k` focal fish on the full frozen state."""
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to its workers; this arm is the CPU implementation on ALL host cores, and the
    # oracle port is an OpenMP loop: set the thread count before libgomp initialises (nothing has loaded it yet)
    os.environ["OMP_NUM_THREADS"] = str(min(os.cpu_count() or 1, 32))
    from fish_abm_b200 import synth
    from oracle.pyoracle import Oracle, RefLib, State
    # the same workload as our arm: C3 (1M fish) for one GPU, C4 (16.8M fish) for several.  The compiled reference has a
    # compile-time capacity (#define num) and passes the whole school by value on the stack, so it exists here for 1M
    # fish; at C4 size the C restatement of the same loops (pinned against it) stands in (kind "port").
    n = args.fish or (N_C3 if args.gpus <= 1 else N_C4)
    w = synth.world_size_for(n)
    d = synth.uniform_state(n, w, w, seed=0x5EED)
    cores = min(os.cpu_count() or 1, 32)
    kind = "reference" if (n == N_C3 and RefLib.available(N_C3)) else "port"
    # >= 16 focal fish per thread and step at C3 (about 2.5 s of all-core work per step); a C4-sized fish costs 17x more
    k = 16 * cores if n <= N_C3 else 4 * cores
    ids = stratified_ids(d["lag"], k, seed=1)
    if kind == "reference":
        R = RefLib(N_C3)
        R.set_world(w, w)
        R.set_state(State(**d))
        run = lambda: R.time_focal(ids, cores)
    else:
        O = Oracle(State(**d), w, w, None, seed=1)
        used = O.time_focal(ids[:1])[1]   # threads OpenMP really gives us
        cores = used
        run = lambda: O.time_focal(ids)[0]
    if n > N_C3:   # a C4-sized step is ~15-20 s of all-core work: keep the whole arm to about a minute and a half
        args.steps, args.warmup = min(args.steps, 4), min(args.warmup, 1)
    else:
        args.steps, args.warmup = min(args.steps, 10), min(args.warmup, 1)
    for _ in range(min(args.warmup, 2)):
        run()
    secs = [run() for _ in range(args.steps)]
    tot = float(np.sum(secs))
    value = k * args.steps / tot
    cfg = {"workload": workload_string(n, w), "n_fish": n, "world": w, "l2": "n/a (CPU)"}
    sample = (f"{k} focal fish per step (moving : lagged mix as in the school) on the full {n}-fish frozen state, each doing the reference's own per-fish work of "
              f"move()+sample() (8 all-pairs passes if lag==0 else 3); agent-steps/s = focal fish / seconds (extrapolation, "
              f"the O(N^2) step itself is ~days)")
    line = {"impl": "reference", "metric": "agent_steps_per_sec", "value": value, "unit": "agent-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": "agent-steps/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(n, w, d, budget_s=12.0):
    """bounded CPU timing on rank 0 (N=1 only): see reference_arm"""
    from oracle.pyoracle import Oracle, RefLib, State
    cores = min(os.cpu_count() or 1, 32)
    kind = "reference" if (n == N_C3 and RefLib.available(N_C3)) else "port"
    if kind == "reference":
        R = RefLib(N_C3); R.set_world(w, w); R.set_state(State(**d))
        run = lambda ids: R.time_focal(ids, cores)
    else:
        O = Oracle(State(**{k: v.copy() for k, v in d.items()}), w, w, None, seed=1)
        run = lambda ids: O.time_focal(ids)[0]
    t1 = run(stratified_ids(d["lag"], 2 * cores, seed=2))
    k = int(max(2 * cores, min(64 * cores, 2 * cores * budget_s / max(t1, 1e-3))))
    secs = run(stratified_ids(d["lag"], k, seed=3))
    return {"value": k / secs, "unit": "agent-steps/s", "cores": cores, "kind": kind,
            "sample": f"{k} focal fish (moving : lagged mix as in the school) of the {n}-fish state in {secs:.1f} s on {cores} host threads (per-fish work of move()+sample(): "
                      f"8 all-pairs passes if lag==0 else 3), extrapolated to agent-steps/s = focal fish / seconds"}


def whole_school_leg(n: int, device: int, warmup: int, steps: int, d=None, q=None, keep: bool = False):
    """one whole school of n fish on ONE GPU: device-timed throughput of `steps` steps (+ the trailing sample pass).  Used for
    the 1-GPU base of the C4 scaling curve (N=1 line: c4_one_gpu; N>1 lines: every rank runs it as the parity reference of
    its slab).  keep=True returns the live School as well."""
    from fish_abm_b200 import School, synth
    w = synth.world_size_for(n)
    if d is None:
        d = synth.uniform_state(n, w, w, seed=0x5EED)
    if q is None:
        q100 = np.loadtxt(os.path.join(ROOT, "tests", "golden", "quality.csv"), delimiter=",", dtype=np.int32)
        q = synth.tile_quality(q100, w, w)
    S = School(n=n, w=w, h=w, seed=0x5EED, device=device)
    S.set_state(**d)
    S.set_quality(q)
    for _ in range(warmup):
        S.step(1)
    S.flush()
    S.step(steps)
    ms = S.last_step_ms()
    nb_ms, nb_l = S.last_neighbour_ms()
    S.flush()
    ms += S.last_step_ms()
    a, b = S.last_neighbour_ms(); nb_ms += a; nb_l += b
    out = {"value": n * steps / (ms * 1e-3), "unit": "agent-steps/s", "ms_per_step": ms / steps, "steps": steps, "warmup": warmup,
           "neighbour_ms_per_launch": nb_ms / max(nb_l, 1), "workload": workload_string(n, w),
           "l2": "state larger than L2" if n * 76 > (126 << 20) else "state fits L2, not flushed"}
    if keep:
        return out, S
    S.close()
    return out


def small_configs_leg(device: int, with_cpu: bool):
    """BASELINE.json configs[0] / configs[1] next to the reference itself: the verbatim,
    single-threaded fish_mod.cpp (move()+sample(), full sequential steps; -O2 and, as shipped by fishmod_compile.sh:1, -O0)
    on this box's host, beside the GPU's ms per step for the same school."""
    from fish_abm_b200 import Ensemble, School, synth
    from oracle.pyoracle import RefLib, State
    qpath = os.path.join(ROOT, "tests", "golden", "quality.csv")
    q100 = np.loadtxt(qpath, delimiter=",", dtype=np.int32)
    out = {}
    # ---- GPU
    with School(n=60, w=2000, h=2000, seed=1, device=device, speed_norm=60.0) as S:      # C1: the reference's default school
        S.initialize(5.0); S.set_quality(q100)
        S.step(50); S.flush()
        S.step(500); ms = S.last_step_ms(); S.flush(); ms += S.last_step_ms()
        out["C1_n60"] = {"gpu_ms_per_step_cell_list": ms / 500}
    with Ensemble(1, n=60, seed=1, device=device, speed_norm=60.0) as E:
        E.initialize(5.0); E.set_quality(q100)
        E.step(50)
        E.step(500)
        out["C1_n60"]["gpu_ms_per_step_ensemble_kernel"] = E.last_step_ms() / 500
    for n in (1000, 10000):                                                                # C2 and the point in between
        d = synth.uniform_state(n, 2000, 2000, seed=3)
        with School(n=n, w=2000, h=2000, seed=1, device=device) as S:
            S.set_state(**d); S.set_quality(q100)
            S.step(20); S.flush()
            S.step(100); ms = S.last_step_ms(); S.flush(); ms += S.last_step_ms()
            out[("C2_n10000" if n == 10000 else f"n{n}")] = {"gpu_ms_per_step_cell_list": ms / 100}
    if not with_cpu:
        return out
    # ---- CPU: the unmodified reference, one thread, whole sequential steps
    def ref_ms(n, suffix, steps):
        if not RefLib.available(n, suffix):
            return None
        R = RefLib(n, suffix)
        R.set_world(2000, 2000)
        if n == 60:
            R.initialize(60, 5.0, 1)
        else:
            R.set_state(State(**synth.uniform_state(n, 2000, 2000, seed=3)))
        R.get_environment(os.path.dirname(qpath))
        R.rng(1, 0, 0)
        secs, _, _ = R.run(steps, 0, log_sumq=False)
        return 1e3 * secs / steps
    out["C1_n60"].update(cpu_ref_ms_per_step_O2=ref_ms(60, "", 500), cpu_ref_ms_per_step_O0=ref_ms(60, "_O0", 200))
    out["n1000"].update(cpu_ref_ms_per_step_O2=ref_ms(1000, "", 4), cpu_ref_ms_per_step_O0=ref_ms(1000, "_O0", 2))
    out["C2_n10000"].update(cpu_ref_ms_per_step_O2=ref_ms(10000, "", 1), cpu_ref_ms_per_step_O0=None,
                            note="-O0 at N=10k is ~3.2x the -O2 figure (the measured N=60 / N=1000 ratio); not run to keep the bench short")
    out["cpu"] = {"threads": 1, "what": "unmodified fish_mod.cpp (oracle/_ref), move()+sample() as in its main loop, whole sequential steps",
                  "model": next((l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")), "unknown"),
                  "host_cores": os.cpu_count()}
    return out


def ensemble_key(device: int, schools: int, steps: int = 20, replicate0: int = 0):
    """BASELINE.json configs[4] on this rank's GPU: `schools` replicate schools of 200 fish sweeping speed (2.5..20),
    sample-rate reset (15/30/45) and fed speed factor (0.25/0.5/0.75) per school; K steps in ONE launch"""
    from fish_abm_b200 import Ensemble
    q100 = np.loadtxt(os.path.join(ROOT, "tests", "golden", "quality.csv"), delimiter=",", dtype=np.int32)
    n = 200
    E = Ensemble(schools, n=n, seed=0x5EED, replicate0=replicate0, device=device, speed_norm=float(n))
    E.initialize(5.0)
    g = E.get_state()
    g["speed"][:] = np.array([2.5, 5, 7.5, 10, 12.5, 15, 17.5, 20])[((np.arange(schools) + replicate0) % 8)][:, None]
    E.set_state(**g)
    E.set_quality(q100)
    swept = E.sweep_default() if hasattr(E, "sweep_default") else None
    E.step(3)
    E.step(steps)
    ms = E.last_step_ms()
    eaten = int((3499 - np.asarray(E.sum_array())).sum())
    E.close()
    return {"value": schools * n * steps / (ms * 1e-3), "unit": "agent-steps/s", "ms_per_step": ms / steps, "schools": schools, "n_fish": n,
            "steps": steps, "school_steps_per_sec": schools * steps / (ms * 1e-3), "food_eaten": eaten, "swept": swept or "speed 2.5..20"}


# ------------------------------------------------------------------------------------------------------------------
def ours(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist
    from fish_abm_b200 import School, synth

    torch.cuda.set_device(local_rank)
    if world > 1:
        from fish_abm_b200.slab import run_slab_bench
        return run_slab_bench(args, rank, world, local_rank)

    n = args.fish or N_C3
    w = synth.world_size_for(n)
    q100 = np.loadtxt(os.path.join(ROOT, "tests", "golden", "quality.csv"), delimiter=",", dtype=np.int32)
    d = synth.uniform_state(n, w, w, seed=0x5EED)
    q = synth.tile_quality(q100, w, w)
    S = School(n=n, w=w, h=w, seed=0x5EED, device=local_rank)
    S.set_state(**d)
    S.set_quality(q)

    state_bytes = n * (48 + 4 + 24)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if state_bytes < (200 << 20) else None

    def l2_flush():
        if flush is not None:
            flush.fill_(1)
            torch.cuda.synchronize()

    # ---- device-resident throughput: K steps, each timed on the library's stream with CUDA events
    clocks = ClockSampler(local_rank)   # sampled from the warm-up through the end of the timed steps
    clocks.start()
    for _ in range(args.warmup):
        S.step(1)
    S.flush()
    launches0 = S.launch_count()
    torch.cuda.synchronize()
    step_ms, nb_ms, nb_launches = [], 0.0, 0
    for _ in range(args.steps):
        l2_flush()
        S.step(1)
        step_ms.append(S.last_step_ms())
        a, b = S.last_neighbour_ms(); nb_ms += a; nb_launches += b
    S.flush()  # the trailing sample() of the last step belongs to the K steps
    step_ms.append(S.last_step_ms())
    a, b = S.last_neighbour_ms(); nb_ms += a; nb_launches += b
    torch.cuda.synchronize()
    launches = S.launch_count() - launches0
    clk = clocks.stop()
    total_ms = float(np.sum(step_ms))
    value = n * args.steps / (total_ms * 1e-3)

    # ---- roofline of the dominant kernel (neighbour_kernel, FP32 pipe): algorithmic flops from exact pair counters
    S.enable_stats(True)
    cand = inter = focal = 0.0
    reps = 4
    for _ in range(reps):
        S.step(1)
        st = S.last_pass_stats()
        cand += st["candidates"]; inter += st["interacting"]; focal += st["focal"]
    S.flush()
    S.enable_stats(False)
    cand /= reps; inter /= reps; focal /= reps
    peaks, how = measured_peaks()
    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
    fp32_peak = sm_count * FP32_LANES_PER_SM * 2 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
    per_launch_ms = nb_ms / max(nb_launches, 1)
    achieved = flops_alg(cand, inter) / (per_launch_ms * 1e-3) / 1e12
    kname = {"v1": "neighbour_kernel<", "v3": "neighbour_kernel_v3"}.get(os.environ.get("FISHABM_NEIGHBOUR", ""), "neighbour_kernel_v4")
    cap = ncu_capture(kname)
    # issue-slot fraction: the kernel is bound by instruction issue, not by the FP32 pipe, so the fraction of the SMs' issue
    # slots (SMs x 4 schedulers x clock x launch time) it fills says more than flops / peak.  ncu measures it directly for the
    # captured launch (smsp__issue_active); for the timed launches it is estimated from the capture's warp instructions per
    # focal fish x the live focal count.
    issue = None
    if cap and cap.get("warp_instructions") and cap.get("focal_per_launch"):
        slots = sm_count * 4 * peaks.get("sm_max_mhz", 1965.0) * 1e6 * per_launch_ms * 1e-3
        issue = {"ncu_issue_active_pct": cap.get("issue_active_pct"), "warp_instructions_per_focal_fish": cap["warp_instructions"] / cap["focal_per_launch"],
                 "estimated_frac_timed_launches": cap["warp_instructions"] / cap["focal_per_launch"] * focal / slots,
                 "how": "ncu: smsp__issue_active of the captured launch (the measured figure); estimate: (capture's warp instructions / its focal fish) x live "
                        "focal fish / (SMs x 4 x sm_max_mhz x launch time) -- an UPPER estimate: the per-cell work is spread over more focal fish in the "
                        "timed launches than in the captured one"}
    roofline = {"bound": "fp32", "kernel": kname.rstrip("<"), "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": achieved / fp32_peak, "traffic": cap["dram_bytes"] if cap else None,
                "traffic_note": "ncu dram__bytes_read+write of one launch (the 1M-fish state is L2-resident: 16 B x 1M hot records)" if cap else None,
                "issue_slots": issue,
                "ncu": {k: cap[k] for k in ("duration", "issue_active_pct", "pipe_fma_pct", "pipe_alu_pct", "pipe_xu_pct", "pipe_lsu_pct", "warp_instructions", "threads_per_instruction", "registers", "source") if k in cap} if cap else None,
                "binding_limit": "instruction issue (85 % of the issue slots are used): a round of 32 pairs costs ~100 issued instructions "
                                 "(geometry, rounding bands, polynomial atan2, weight, MUFU sin/cos, fixed-point sums) of which SURVEY 8(d)'s convention counts "
                                 "20 flops per interacting pair, and a candidate ~0.4 (packed-fp16 disc filter) of which it counts 13; see DESIGN.md section 3",
                "peak_source": f"SMs({sm_count}) x 128 lanes x 2 x sm_max_mhz ({how} clock); FP32 is not in MEASURED_PEAKS.json",
                "flops_per_launch": flops_alg(cand, inter), "candidates_per_launch": cand, "interacting_per_launch": inter,
                "focal_per_launch": focal, "launch_ms": per_launch_ms, "share_of_step": nb_ms / total_ms}
    # the HBM-bound rest of the step (update + counting sort): algorithmic bytes per fish-step, SURVEY 8(d)
    rest_ms = (total_ms - nb_ms) / args.steps
    rest_bytes = n * (48 + 28 + 48 + 8 + 8 + 48 + 48 + 4)  # update r/w, pass outputs, rank/cell, scatter r/w
    roofline_hbm = {"bound": "hbm", "kernels": "update+scan+scatter", "achieved": rest_bytes / (rest_ms * 1e-3) / 1e9,
                    "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": rest_bytes / (rest_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "ms_per_step": rest_ms}

    # ---- end to end through the C ABI with host buffers (pinned), H2D of the state and D2H of the result every step
    pinned = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in S.get_state().items()}
    host = {k: v.numpy() for k, v in pinned.items()}
    h2d = sum(v.nbytes for v in host.values())
    e2e_steps = max(3, min(args.steps, 10))
    out = {k: torch.empty_like(v).pin_memory().numpy() for k, v in pinned.items()}
    from fish_abm_b200.school import _p
    import ctypes as C
    order = ("coords", "dir", "speed", "speed_factor", "lag", "sample_rate", "food_intake")
    sumq = np.zeros(1, np.int64)

    bufs = [host, out]

    def e2e_step():
        src, dst = bufs  # the state a step returns is the next step's input: the two pinned sets swap roles
        S._ck(S.L.fishabm_set_state(S.ctx, *(_p(src[k]) for k in order)))
        S._ck(S.L.fishabm_step(S.ctx, 1))
        S._ck(S.L.fishabm_get_state(S.ctx, *(_p(dst[k]) for k in order)))
        S._ck(S.L.fishabm_sum_quality(S.ctx, _p(sumq)))
        bufs.reverse()

    e2e_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e = {"value": n * e2e_steps / e2e_s, "unit": "agent-steps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(h2d + 8),
           "steps": e2e_steps, "api": "fishabm_set_state + fishabm_step(1) + fishabm_get_state + fishabm_sum_quality, pinned host buffers"}

    # ---- live parity figure at the benchmark size: the state the timed run ended in, 1024 fish against ALL others in the
    # library's fp64 check mode (brute force, double precision, the reference's formulas; pinned against the CPU oracle by
    # tests/test_check_fp64_gpu.py)
    cid = np.random.default_rng(9).choice(n, 1024, replace=False).astype(np.int32)
    cc, ct, cf = S.check_fp64(cid)
    zc, st = S.zone_counts()[cid], S.social()[0][cid]
    okc = (cf & 6) == 0
    parity = {"checked_fish": int(len(cid)), "against": "fishabm_check_fp64 (fp64 brute force over all fish)",
              "count_mismatches": int((zc != cc).any(axis=1).sum()),
              "max_turn_err_rad": float(np.abs(np.deg2rad(st - ct))[okc].max()), "knife_edge_or_tie_skipped": int((~okc).sum())}

    cpu = cpu_baseline_leg(n, w, d) if not args.no_cpu_baseline else None
    S.close()
    l2_note = "flushed between steps (256 MiB write)" if flush is not None else "state larger than L2"
    flush = None
    torch.cuda.empty_cache()
    # the 1-GPU base of the C4 scaling curve (the N>1 lines run C4), the small configurations, and configs[4]
    extra = {}
    if n == N_C3 and not args.fish:
        extra["c4_one_gpu"] = whole_school_leg(N_C4, local_rank, warmup=3, steps=10)
        extra["small_configs"] = small_configs_leg(local_rank, with_cpu=not args.no_cpu_baseline)
        if not args.no_ensemble:
            extra["ensemble"] = ensemble_key(local_rank, args.schools)
    line = {"metric": "agent_steps_per_sec", "value": value, "unit": "agent-steps/s", "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "scaling_note": "the N=1 line is C3 (1M fish); the N>1 lines are C4 (16.8M fish, strong scaling over 2/4/8 GPUs) whose 1-GPU base "
                            "is c4_one_gpu below (and is re-measured inside every N>1 line)",
            "config": {"workload": workload_string(n, w), "n_fish": n, "world": w,
                       "l2": l2_note,
                       "timing": "CUDA events on the library's stream around each fishabm_step(1); K steps + the trailing sample pass"},
            "clocks": clk, "gpu_launches": int(launches), "e2e": e2e, "roofline": roofline, "roofline_hbm": roofline_hbm,
            "cpu_baseline": cpu, "parity_spot_check": parity, "build": build_info(), **extra}
    print(json.dumps(line), flush=True)


def ensemble_leg(args, rank: int, world: int, local_rank: int):
    """BASELINE.json configs[4]: 65,536 replicate schools of N=200 (speed sweep 2.5..20 as in data/), replicates split
    across the GPUs with no communication (weak scaling in the contract's sense: nothing is exchanged).  Not the
    headline line; run with --workload ensemble."""
    import torch
    from fish_abm_b200 import Ensemble
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    total, n = args.schools, 200
    r0, r1 = rank * total // world, (rank + 1) * total // world
    R = r1 - r0
    q100 = np.loadtxt(os.path.join(ROOT, "tests", "golden", "quality.csv"), delimiter=",", dtype=np.int32)
    E = Ensemble(R, n=n, seed=0x5EED, replicate0=r0, device=local_rank, speed_norm=float(n))
    E.initialize(5.0)
    g = E.get_state()
    speeds = np.array([2.5, 5, 7.5, 10, 12.5, 15, 17.5, 20])[(np.arange(r0, r1) % 8)]
    g["speed"][:] = speeds[:, None]
    E.set_state(**g)
    E.set_quality(q100)
    for _ in range(args.warmup):
        E.step(1)
    torch.cuda.synchronize()
    launches0 = E.launch_count()
    E.step(args.steps)   # K steps of every school in ONE launch: state and food grid stay in shared memory
    ms = E.last_step_ms()
    launches = E.launch_count() - launches0
    t0 = time.perf_counter()
    sumq, _ = E.run_logged(args.steps)   # e2e flavour: the same steps with the depletion log copied back to the host
    e2e_s = time.perf_counter() - t0
    eaten = int((3499 - np.asarray(E.sum_array())).sum())
    E.close()
    vals = np.array([ms, e2e_s, launches, eaten], np.float64)
    if world > 1:
        t = torch.tensor(vals, device="cuda")
        mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, e2e_s, launches, eaten = float(mx[0]), float(mx[1]), int(sm[2]), int(sm[3])
    if rank == 0:
        line = {"metric": "agent_steps_per_sec", "value": total * n * args.steps / (ms * 1e-3), "unit": "agent-steps/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"C5: ensemble of {total} replicate schools of N={n}, 2000x2000 window each, quality.csv, speed sweep 2.5..20, "
                                       f"{world} GPU(s), no communication", "schools": total, "n_fish": n, "l2": "state in shared memory for the whole launch"},
                "gpu_launches": int(launches),
                "e2e": {"value": total * n * args.steps / e2e_s, "unit": "agent-steps/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": int(8 * total), "api": "fishabm_run_logged: K steps + sum_array per step per school copied to the host"},
                "school_steps_per_sec": total * args.steps / (ms * 1e-3), "food_eaten": eaten, "roofline": None, "cpu_baseline": None}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--fish", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="school", choices=["school", "ensemble"])
    ap.add_argument("--schools", type=int, default=65536)
    ap.add_argument("--one-transport", action="store_true", help="N>1: skip the second slab transport")
    ap.add_argument("--no-ensemble", action="store_true", help="skip the configs[4] ensemble key")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import __graft_entry__ as g
    if rank == 0:
        g.build()
    if args.impl == "reference":
        if args.steps > 20:
            args.steps = 20
        return reference_arm(args, rank, world)
    if args.workload == "ensemble":
        return ensemble_leg(args, rank, world, local_rank)
    return ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
