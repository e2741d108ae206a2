"""Host-side mirror of the reference's operator interface for the per-step agent update
(fish_mod.cpp:64-82: initialize, move, sample, feed, social, in_zone, get_speed, averageturn, ...), bound to
libfishabm.so through its C ABI.  Arrays use the reference's `individuals` layout (float64 / int32, id order)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class FishAbmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"fishabm error {code}: {msg}")
        self.code = code


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class School:
    """One school resident on one GPU.  Method names follow the reference's functions."""

    def __init__(self, n: int = 60, w: int = 2000, h: int = 2000, seed: int = 0, device: int = 0,
                 rolling: bool = True, speed_norm: float = 0.0, mode: int = _lib.MODE_CELL_FP32,
                 rank: int = 0, nranks: int = 1, force_slab: bool = False, capacity: int = 0, halo_capacity: int = 0,
                 migrant_capacity: int = 0, replicates: int = 1, replicate0: int = 0, model: dict | None = None):
        """rank / nranks > 1: this context is one x-slab of the school (see include/fishabm.h and slab.py); n, w, h and
        every array passed in or out still describe the WHOLE school.
        replicates > 1 (or mode=MODE_ENSEMBLE): an ensemble of independent schools; arrays are flat, replicates*n
        entries, replicate-major (class Ensemble reshapes them)."""
        self.L = _lib.load()
        p = _lib.Params()
        self.L.fishabm_params_default(C.byref(p))
        p.n, p.w, p.h, p.seed, p.device, p.rolling, p.speed_norm, p.mode = n, w, h, seed, device, int(rolling), speed_norm, mode
        p.rank, p.nranks, p.force_slab = rank, nranks, int(force_slab)
        p.capacity, p.halo_capacity, p.migrant_capacity = capacity, halo_capacity, migrant_capacity
        p.replicates, p.replicate0 = replicates, replicate0
        for k, v in (model or {}).items():   # fishabm_model_params: the reference's literals as run-time parameters
            if not hasattr(p.model, k):
                raise KeyError(f"unknown model parameter {k!r}")
            setattr(p.model, k, v)
        self.replicates = replicates
        self.is_slab = nranks > 1 or force_slab
        self.params = p
        self.ctx = C.c_void_p()
        rc = self.L.fishabm_create(C.byref(p), C.byref(self.ctx))
        if rc != 0:
            raise FishAbmError(rc, self.L.fishabm_last_error(None).decode())
        self.n, self.w, self.h = n, w, h
        self.total = n * replicates
        self.rows, self.cols = h // 20, w // 20

    def close(self):
        if getattr(self, "ctx", None) is not None and self.ctx:
            self.L.fishabm_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc: int):
        if rc != 0:
            raise FishAbmError(rc, self.L.fishabm_last_error(self.ctx).decode())

    # ---- state -------------------------------------------------------------------------------------------
    def set_state(self, coords, dir, speed, speed_factor, lag, sample_rate, food_intake):
        f = lambda a: np.ascontiguousarray(a, np.float64)
        i = lambda a: np.ascontiguousarray(a, np.int32)
        coords, dir, speed, speed_factor = f(coords).reshape(-1, 2), f(dir), f(speed), f(speed_factor)
        lag, sample_rate, food_intake = i(lag), i(sample_rate), i(food_intake)
        assert coords.shape[0] == dir.shape[0] == self.total
        self._ck(self.L.fishabm_set_state(self.ctx, _p(coords), _p(dir), _p(speed), _p(speed_factor), _p(lag),
                                          _p(sample_rate), _p(food_intake)))

    NOT_OWNED = np.iinfo(np.int32).min
    FIELDS = ("coords", "dir", "speed", "speed_factor", "lag", "sample_rate", "food_intake")

    def set_state_owned(self, ids, coords, dir, speed, speed_factor, lag, sample_rate, food_intake, wait: bool = True):
        """slab-local set_state: only this slab's own fish (global ids `ids`); collective over the slabs.  wait=False only
        enqueues (several slabs driven from one process: enqueue all, then wait() each)."""
        f = lambda a: np.ascontiguousarray(a, np.float64)
        i = lambda a: np.ascontiguousarray(a, np.int32)
        ids = i(ids)
        args = (ids, f(coords).reshape(-1, 2), f(dir), f(speed), f(speed_factor), i(lag), i(sample_rate), i(food_intake))
        self._keep = args  # the copies are asynchronous: keep the host arrays alive until wait()
        self._ck(self.L.fishabm_set_state_owned(self.ctx, len(ids), *(_p(a) for a in args)))
        if wait:
            self.wait()

    def get_state_owned(self, out: dict | None = None) -> dict:
        """the fish this slab (or school) owns now: compact arrays plus their global ids"""
        n = self.n
        if out is None:
            out = dict(ids=np.empty(n, np.int32), coords=np.empty((n, 2)), dir=np.empty(n), speed=np.empty(n), speed_factor=np.empty(n),
                       lag=np.empty(n, np.int32), sample_rate=np.empty(n, np.int32), food_intake=np.empty(n, np.int32))
        cnt = C.c_int32(len(out["ids"]))
        self._ck(self.L.fishabm_get_state_owned(self.ctx, C.byref(cnt), _p(out["ids"]), *(_p(out[k]) for k in self.FIELDS)))
        return {k: v[:cnt.value] for k, v in out.items()}

    def get_state(self) -> dict:
        """Whole school: every fish.  Slab: only the entries of the fish this slab owns are written; the others keep
        NaN / NOT_OWNED (see owned_mask)."""
        n = self.total
        out = dict(coords=np.empty((n, 2)), dir=np.empty(n), speed=np.empty(n), speed_factor=np.empty(n),
                   lag=np.empty(n, np.int32), sample_rate=np.empty(n, np.int32), food_intake=np.empty(n, np.int32))
        if self.is_slab:
            for v in out.values():
                v.fill(np.nan if v.dtype == np.float64 else self.NOT_OWNED)
        self._ck(self.L.fishabm_get_state(self.ctx, *(_p(out[k]) for k in
                                                      ("coords", "dir", "speed", "speed_factor", "lag", "sample_rate", "food_intake"))))
        return out

    @staticmethod
    def owned_mask(state: dict) -> np.ndarray:
        return state["lag"] != School.NOT_OWNED

    def initialize(self, speed: float = 5.0):
        """initialize(fish, num, 2, speed, struc), fish_mod.cpp:116,183-206"""
        self._ck(self.L.fishabm_initialize(self.ctx, float(speed)))

    def set_models(self, models):
        """one dict of model parameters per replicate (missing keys keep the reference's values); a single school takes a
        list of one.  include/fishabm.h: fishabm_set_replicate_models"""
        arr = (_lib.ModelParams * len(models))()
        for m, over in zip(arr, models):
            self.L.fishabm_model_params_default(C.byref(m))
            for k, v in over.items():
                if not hasattr(m, k):
                    raise KeyError(f"unknown model parameter {k!r}")
                setattr(m, k, v)
        self._ck(self.L.fishabm_set_replicate_models(self.ctx, arr, len(models)))

    # ---- food grid -----------------------------------------------------------------------------------------
    def set_quality(self, q):
        q = np.ascontiguousarray(q, np.int32)
        self._ck(self.L.fishabm_set_quality(self.ctx, _p(q), q.shape[0], q.shape[1], q.shape[1]))

    def get_quality(self, replicate: int = 0) -> np.ndarray:
        """Whole school: the grid.  Slab: the owned columns, -1 elsewhere."""
        q = np.full((self.rows, self.cols), -1, np.int32)
        self._ck(self.L.fishabm_get_quality(self.ctx, replicate, _p(q), self.rows, self.cols, self.cols))
        return q

    def sum_array(self):
        """sum_array(quality, struc), fish_mod.cpp:593-601 (one value per replicate for an ensemble)"""
        s = np.zeros(self.replicates, np.int64)
        self._ck(self.L.fishabm_sum_quality(self.ctx, _p(s)))
        return int(s[0]) if self.replicates == 1 else s

    def get_environment(self, path: str):
        """get_environment(environment, quality, struc), fish_mod.cpp:603-629 (the CSV parse; painting is viewer-only)"""
        q = np.zeros((self.rows, self.cols), np.int32)
        r, c = C.c_int32(0), C.c_int32(0)
        rc = self.L.fishabm_load_quality_csv(path.encode(), _p(q), self.rows, self.cols, self.cols, C.byref(r), C.byref(c))
        if rc != 0:
            raise FishAbmError(rc, f"cannot read {path} as a {self.rows}x{self.cols} grid")
        self.set_quality(q)
        return q

    def create_environment(self, n_seeds: int, seed: int | None = None, replicate: int = 0):
        """create_environment(environment, quality, struc), fish_mod.cpp:482-531, with a live seed count: generates the
        food map on the host and uploads it"""
        q = create_environment(n_seeds, self.rows, self.cols, int(self.params.seed) if seed is None else seed, replicate)
        self.set_quality(q)
        return q

    # ---- frozen-state queries --------------------------------------------------------------------------------
    def in_zone(self, fov: int, r: float) -> np.ndarray:
        out = np.full(self.total, -1, np.int32)
        self._ck(self.L.fishabm_in_zone(self.ctx, int(fov), float(r), _p(out)))
        return out

    def zone_counts(self) -> np.ndarray:
        out = np.full((self.total, 4), -1, np.int32)
        self._ck(self.L.fishabm_zone_counts(self.ctx, _p(out)))
        return out

    def social(self):
        turn = np.full(self.total, np.nan)
        na, nl, nt = (np.full(self.total, -1, np.int32) for _ in range(3))
        self._ck(self.L.fishabm_social(self.ctx, _p(turn), _p(na), _p(nl), _p(nt)))
        return turn, na, nl, nt

    def check_fp64(self, ids):
        """fp64 check mode: (counts [k,4], turn [k], flags [k]) of the listed fish by brute force over the whole school in
        double precision (include/fishabm.h: fishabm_check_fp64)"""
        ids = np.ascontiguousarray(ids, np.int32)
        k = len(ids)
        counts, turn, flags = np.empty((k, 4), np.int32), np.empty(k), np.empty(k, np.int32)
        self._ck(self.L.fishabm_check_fp64(self.ctx, _p(ids), k, _p(counts), _p(turn), _p(flags)))
        return counts, turn, flags

    def get_speed(self) -> np.ndarray:
        out = np.full(self.total, np.nan)
        self._ck(self.L.fishabm_get_speed(self.ctx, _p(out)))
        return out

    # ---- stepping ------------------------------------------------------------------------------------------------
    def move(self):
        self._ck(self.L.fishabm_move(self.ctx))

    def sample(self):
        self._ck(self.L.fishabm_sample(self.ctx))

    def step(self, nsteps: int = 1):
        self._ck(self.L.fishabm_step(self.ctx, int(nsteps)))

    def flush(self):
        """complete the pending sample() of the last step (see include/fishabm.h)"""
        self._ck(self.L.fishabm_flush(self.ctx))

    def feed(self, ids):
        ids = np.ascontiguousarray(ids, np.int32)
        self._ck(self.L.fishabm_feed(self.ctx, _p(ids), len(ids)))

    def run_logged(self, nsteps: int, log_sumq: bool = True, intake_every: int = 0):
        """the reference's two commented-out loggers (fish_mod.cpp:132-141,149-154): sum_array after every step
        ([nsteps] or [nsteps, replicates]) and food_intake every `intake_every` steps ([rows, replicates*n])"""
        sumq = (np.zeros((nsteps, self.replicates), np.int64) if self.replicates > 1 else np.zeros(nsteps, np.int64)) if log_sumq else None
        rows = (nsteps + intake_every - 1) // intake_every if intake_every > 0 else 0
        intake = np.zeros((rows, self.total), np.int32) if rows else None
        self._ck(self.L.fishabm_run_logged(self.ctx, nsteps, _p(sumq), _p(intake), intake_every))
        return sumq, intake

    @property
    def step_index(self) -> int:
        s = C.c_uint32(0)
        self.L.fishabm_get_step(self.ctx, C.byref(s))
        return int(s.value)

    @step_index.setter
    def step_index(self, v: int):
        self.L.fishabm_set_step(self.ctx, int(v))

    # ---- introspection ---------------------------------------------------------------------------------------------
    def draw(self, stream: int, step: int, fish_id: int, a: int, b: int, k: int = 0, replicate: int = 0) -> int:
        out = C.c_int32(0)
        self._ck(self.L.fishabm_draw(self.ctx, stream, replicate, step, fish_id, k, a, b, C.byref(out)))
        return int(out.value)

    def last_step_ms(self) -> float:
        ms = C.c_float(0)
        self.L.fishabm_last_step_ms(self.ctx, C.byref(ms))
        return float(ms.value)

    def last_neighbour_ms(self):
        ms, k = C.c_float(0), C.c_int32(0)
        self.L.fishabm_last_neighbour_ms(self.ctx, C.byref(ms), C.byref(k))
        return float(ms.value), int(k.value)

    def enable_stats(self, on: bool = True):
        self.L.fishabm_enable_stats(self.ctx, int(on))

    PHASES = ("neighbour", "update", "exchange", "sort")

    def enable_phase_timing(self, on: bool = True):
        self.L.fishabm_enable_phase_timing(self.ctx, int(on))

    def last_phase_ms(self) -> dict:
        """device ms by phase, summed over the steps of the last step() call (include/fishabm.h)"""
        ms, k = (C.c_float * 4)(), C.c_int32(0)
        self.L.fishabm_last_phase_ms(self.ctx, C.byref(ms), C.byref(k))
        return dict(zip(self.PHASES, (float(v) for v in ms)), steps=int(k.value))

    def last_pass_stats(self) -> dict:
        s = _lib.PassStats()
        self._ck(self.L.fishabm_last_pass_stats(self.ctx, C.byref(s)))
        return dict(focal=s.focal, candidates=s.candidates, interacting=s.interacting, rechecked=s.rechecked)

    def launch_count(self) -> int:
        v = C.c_uint64(0)
        self.L.fishabm_launch_count(self.ctx, C.byref(v))
        return int(v.value)

    # ---- slab decomposition ------------------------------------------------------------------------------------------
    def slab_info(self) -> dict:
        s = _lib.SlabInfo()
        self._ck(self.L.fishabm_slab_get_info(self.ctx, C.byref(s)))
        return {k: getattr(s, k) for k, _ in s._fields_}

    def slab_connect(self, unique_id: bytes):
        buf = (C.c_char * _lib.UNIQUE_ID_BYTES).from_buffer_copy(unique_id)
        self._ck(self.L.fishabm_slab_connect(self.ctx, C.cast(buf, C.c_void_p), _lib.UNIQUE_ID_BYTES))

    def slab_endpoint(self) -> bytes:
        """this slab's inbox (peer-memory transport) as opaque bytes to hand to its neighbours"""
        e = _lib.SlabEndpoint()
        self._ck(self.L.fishabm_slab_get_endpoint(self.ctx, C.byref(e)))
        return bytes(e)

    def slab_attach(self, left: bytes, right: bytes):
        """attach to the two neighbours' inboxes: from now on move() packs straight into them over NVLink"""
        l, r = _lib.SlabEndpoint.from_buffer_copy(left), _lib.SlabEndpoint.from_buffer_copy(right)
        self._ck(self.L.fishabm_slab_attach(self.ctx, C.byref(l), C.byref(r)))

    def step_enqueue(self, nsteps: int = 1):
        self._ck(self.L.fishabm_step_enqueue(self.ctx, int(nsteps)))

    def wait(self):
        self._ck(self.L.fishabm_wait(self.ctx))

    def slab_begin_step(self):
        self._ck(self.L.fishabm_slab_begin_step(self.ctx))

    def slab_end_step(self):
        self._ck(self.L.fishabm_slab_end_step(self.ctx))

    def slab_buffers(self):
        """(out_left, out_right, in_left, in_right) device pointers and the message size in bytes"""
        ptrs = [C.c_void_p() for _ in range(4)]
        nbytes = C.c_int64(0)
        self._ck(self.L.fishabm_slab_buffers(self.ctx, *(C.byref(p) for p in ptrs), C.byref(nbytes)))
        return tuple(p.value for p in ptrs), int(nbytes.value)


class Ensemble(School):
    """replicates independent schools of n fish (BASELINE.json configs[4]); per-fish arrays are [replicates, n]"""

    def __init__(self, replicates: int, n: int = 200, **kw):
        super().__init__(n=n, replicates=replicates, mode=_lib.MODE_ENSEMBLE, **kw)

    def set_state(self, coords, dir, speed, speed_factor, lag, sample_rate, food_intake):
        r = lambda a: np.asarray(a).reshape(self.total, *np.asarray(a).shape[2:])
        super().set_state(r(coords), r(dir), r(speed), r(speed_factor), r(lag), r(sample_rate), r(food_intake))

    def get_state(self) -> dict:
        g = super().get_state()
        return {k: v.reshape(self.replicates, self.n, *v.shape[1:]) for k, v in g.items()}

    SWEEP = {"sample_rate_fed": (15, 30, 45), "fed_speed_factor": (0.25, 0.5, 0.75)}

    def sweep_default(self) -> str:
        """BASELINE.json configs[4]: replicate r gets sample_rate_fed = SWEEP[..][(r // 8) % 3] and fed_speed_factor =
        SWEEP[..][(r // 24) % 3] (the speed sweep itself is per-fish state: speed = 2.5..20 by r % 8)"""
        r = np.arange(self.replicates) + int(self.params.replicate0)
        self.set_models([{"sample_rate_fed": self.SWEEP["sample_rate_fed"][(k // 8) % 3],
                          "fed_speed_factor": self.SWEEP["fed_speed_factor"][(k // 24) % 3]} for k in r])
        return "speed 2.5..20 (8) x sample_rate_fed 15/30/45 x fed_speed_factor 0.25/0.5/0.75, one combination per replicate"

    def zone_counts(self):
        return super().zone_counts().reshape(self.replicates, self.n, 4)

    def social(self):
        return tuple(a.reshape(self.replicates, self.n) for a in super().social())

    def get_speed(self):
        return super().get_speed().reshape(self.replicates, self.n)

    def run_logged(self, nsteps: int, log_sumq: bool = True, intake_every: int = 0):
        sumq, intake = super().run_logged(nsteps, log_sumq, intake_every)
        if sumq is not None:
            sumq = sumq.reshape(nsteps, self.replicates)
        if intake is not None:
            intake = intake.reshape(-1, self.replicates, self.n)
        return sumq, intake


def slab_unique_id() -> bytes:
    """ncclGetUniqueId through the library (rank 0 calls it and hands the bytes to the other ranks)"""
    L = _lib.load()
    buf = (C.c_char * _lib.UNIQUE_ID_BYTES)()
    rc = L.fishabm_slab_unique_id(C.cast(buf, C.c_void_p), _lib.UNIQUE_ID_BYTES)
    if rc != 0:
        raise FishAbmError(rc, L.fishabm_last_error(None).decode())
    return bytes(buf)


def create_environment(n_seeds: int, rows: int = 100, cols: int = 100, seed: int = 0, replicate: int = 0) -> np.ndarray:
    """the reference's procedural food map (fish_mod.cpp:482-531); host-side, needs no GPU"""
    q = np.zeros((rows, cols), np.int32)
    rc = _lib.load().fishabm_create_environment(seed, replicate, int(n_seeds), _p(q), rows, cols, cols)
    if rc != 0:
        raise FishAbmError(rc, "fishabm_create_environment: bad argument")
    return q


def averageturn(v, comb_weight: bool = False) -> float:
    """averageturn(vec, comb_weight), fish_mod.cpp:379-408"""
    v = np.ascontiguousarray(v, np.float64)
    return float(_lib.load().fishabm_averageturn(_p(v), len(v), int(comb_weight)))


def correctangle(d: float) -> float:
    return float(_lib.load().fishabm_correctangle(float(d)))


def correct_coords(xy, w: int = 2000, h: int = 2000):
    a = np.array(xy, np.float64)
    _lib.load().fishabm_correct_coords(_p(a), w, h)
    return a
