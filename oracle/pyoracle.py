"""oracle/pyoracle.py -- TEST INFRASTRUCTURE (parity oracle), not product code.

ctypes bindings for
  * ``Oracle``  : oracle/libfishoracle.so, the plain-C fp64 restatement (oracle/fish_oracle.c);
  * ``RefLib``  : oracle/_ref/libfishref_<num>.so, the UNMODIFIED reference fish_mod.cpp compiled with
                  the replay harness (oracle/ref_harness.cpp, oracle/build_ref.sh).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

STREAM_CHOICE, STREAM_ERROR, STREAM_RANDOMWALK, STREAM_QUALITY, STREAM_SEED, STREAM_SAMPLE, STREAM_INIT = range(7)


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


@dataclass
class State:
    """Host-side school state in the reference's layout (class individuals, fish_mod.cpp:35-49)."""

    coords: np.ndarray  # [n,2] float64
    dir: np.ndarray  # [n] float64, degrees, unwrapped
    speed: np.ndarray
    speed_factor: np.ndarray
    lag: np.ndarray  # int32
    sample_rate: np.ndarray
    food_intake: np.ndarray

    @property
    def n(self) -> int:
        return int(self.dir.shape[0])

    def copy(self) -> "State":
        return State(*(np.array(getattr(self, f), copy=True) for f in
                       ("coords", "dir", "speed", "speed_factor", "lag", "sample_rate", "food_intake")))

    @staticmethod
    def empty(n: int) -> "State":
        return State(np.zeros((n, 2)), np.zeros(n), np.zeros(n), np.ones(n), np.zeros(n, np.int32),
                     np.zeros(n, np.int32), np.zeros(n, np.int32))

    def normalise(self) -> "State":
        self.coords = np.ascontiguousarray(self.coords, np.float64).reshape(-1, 2)
        for f in ("dir", "speed", "speed_factor"):
            setattr(self, f, np.ascontiguousarray(getattr(self, f), np.float64))
        for f in ("lag", "sample_rate", "food_intake"):
            setattr(self, f, np.ascontiguousarray(getattr(self, f), np.int32))
        return self


class OrcModel(C.Structure):
    """orc_model: the reference's literals as parameters (defaults = the reference's values)"""
    _fields_ = [("lag_after_feed", C.c_int), ("rate_fed", C.c_int), ("rate_unfed", C.c_int), ("sample_draw_max", C.c_int),
                ("error_range", C.c_int), ("randomwalk_range", C.c_int), ("alone_align_min", C.c_int), ("pad", C.c_int),
                ("fed_sf", C.c_double), ("w_align", C.c_double), ("w_attract", C.c_double)]

    @staticmethod
    def default(**over) -> "OrcModel":
        m = OrcModel(10, 30, 0, 35, 10, 45, 5, 0, 0.5, 1.7, 1.7 - 1)
        names = {"sample_rate_fed": "rate_fed", "sample_rate_unfed": "rate_unfed", "fed_speed_factor": "fed_sf",
                 "weight_align": "w_align", "weight_attract": "w_attract"}   # the product's field names are accepted too
        for k, v in over.items():
            setattr(m, names.get(k, k), v)
        return m


class _OrcSchool(C.Structure):
    _fields_ = [("n", C.c_int), ("w", C.c_int), ("h", C.c_int), ("rows", C.c_int), ("cols", C.c_int),
                ("q_stride", C.c_int), ("speed_norm", C.c_double), ("seed", C.c_uint64),
                ("replicate", C.c_uint32), ("pad_", C.c_uint32),
                ("coords", C.c_void_p), ("dir", C.c_void_p), ("speed", C.c_void_p),
                ("speed_factor", C.c_void_p), ("lag", C.c_void_p), ("sample_rate", C.c_void_p),
                ("food_intake", C.c_void_p), ("quality", C.c_void_p), ("m", OrcModel)]


class _SocialOut(C.Structure):
    _fields_ = [("turn", C.c_double), ("n_avoid", C.c_int), ("n_align", C.c_int), ("n_attract", C.c_int),
                ("n_ties", C.c_int), ("randomwalk", C.c_int), ("knife", C.c_int)]


SOCIAL_DTYPE = np.dtype([("turn", "f8"), ("n_avoid", "i4"), ("n_align", "i4"), ("n_attract", "i4"),
                         ("n_ties", "i4"), ("randomwalk", "i4"), ("knife", "i4")])
assert SOCIAL_DTYPE.itemsize == C.sizeof(_SocialOut)


def load_quality_csv(path: str) -> np.ndarray:
    """quality.csv as the reference parses it (get_environment, fish_mod.cpp:603-629): rows of ints."""
    return np.loadtxt(path, delimiter=",", dtype=np.int32, ndmin=2)


class Oracle:
    """The C restatement bound to one school (state + world + quality grid + RNG address)."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            path = os.path.join(HERE, "libfishoracle.so")
            if not os.path.exists(path):
                raise RuntimeError(f"{path} missing: run `make -C oracle libfishoracle.so` (or __graft_entry__.build())")
            L = C.CDLL(path)
            L.orc_correctangle.restype = C.c_double
            L.orc_correctangle.argtypes = [C.c_double]
            L.orc_averageturn.restype = C.c_double
            L.orc_averageturn.argtypes = [C.c_void_p, C.c_int, C.c_int]
            L.orc_getangle.restype = C.c_double
            L.orc_getangle.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_int, C.c_int, C.c_uint64,
                                       C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
            L.orc_in_zone.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
            L.orc_in_zone_all.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p]
            L.orc_sum_quality.restype = C.c_long
            L.orc_create_environment.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint32]
            L.orc_create_environment.restype = None
            L.orc_time_focal.restype = C.c_double
            L.orc_time_focal.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
            L.orc_run.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
            cls._lib = L
        return cls._lib

    def __init__(self, state: State, w: int = 2000, h: int = 2000, quality: np.ndarray | None = None,
                 food_cell: int = 20, seed: int = 0, replicate: int = 0, speed_norm: float | None = None, model: dict | None = None):
        self.L = self.lib()
        self.state = state.normalise()
        self.w, self.h = int(w), int(h)
        rows, cols = self.h // food_cell, self.w // food_cell
        if quality is None:
            quality = np.zeros((rows, cols), np.int32)
        self.quality = np.ascontiguousarray(quality, np.int32)
        assert self.quality.shape == (rows, cols), (self.quality.shape, rows, cols)
        self.seed, self.replicate = int(seed), int(replicate)
        self.speed_norm = float(state.n if speed_norm is None else speed_norm)
        self.model = OrcModel.default(**(model or {}))

    def _s(self) -> _OrcSchool:
        st = self.state
        rows, cols = self.quality.shape
        return _OrcSchool(st.n, self.w, self.h, rows, cols, cols, self.speed_norm, self.seed, self.replicate, 0,
                          st.coords.ctypes.data, st.dir.ctypes.data, st.speed.ctypes.data,
                          st.speed_factor.ctypes.data, st.lag.ctypes.data, st.sample_rate.ctypes.data,
                          st.food_intake.ctypes.data, self.quality.ctypes.data, self.model)

    # ---- frozen-state queries
    def in_zone(self, fov: int, r: float) -> np.ndarray:
        out = np.zeros(self.state.n, np.int32)
        s = self._s()
        self.L.orc_in_zone_all(C.byref(s), fov, float(r), _p(out))
        return out

    def zone_counts(self, ids: np.ndarray | None = None) -> np.ndarray:
        """[k,4] = in_zone(330,15), (330,100), (330,200), (180,200) per fish (all fish if ids is None)."""
        s = self._s()
        if ids is None:
            out = np.zeros((self.state.n, 4), np.int32)
            self.L.orc_zone_counts_all(C.byref(s), _p(out))
        else:
            ids = np.ascontiguousarray(ids, np.int32)
            out = np.zeros((len(ids), 4), np.int32)
            self.L.orc_zone_counts_ids(C.byref(s), _p(ids), len(ids), _p(out))
        return out

    def social(self, step: int = 0, ids: np.ndarray | None = None) -> np.ndarray:
        s = self._s()
        if ids is None:
            out = np.zeros(self.state.n, SOCIAL_DTYPE)
            self.L.orc_social_all(C.byref(s), step, _p(out))
        else:
            ids = np.ascontiguousarray(ids, np.int32)
            out = np.zeros(len(ids), SOCIAL_DTYPE)
            self.L.orc_social_ids(C.byref(s), step, _p(ids), len(ids), _p(out))
        return out

    # ---- state updates
    def move_seq(self, step: int):
        s = self._s(); self.L.orc_move_seq(C.byref(s), step)

    def move_sync(self, step: int):
        s = self._s(); self.L.orc_move_sync(C.byref(s), step, None, None)

    def sample(self, step: int):
        s = self._s(); self.L.orc_sample(C.byref(s), step, None)

    def feed(self, fish_id: int):
        s = self._s(); self.L.orc_feed(C.byref(s), int(fish_id))

    def step_sync(self, step: int):
        s = self._s(); self.L.orc_step_sync(C.byref(s), step)

    def step_seq(self, step: int):
        s = self._s(); self.L.orc_step_seq(C.byref(s), step)

    def run(self, steps: int, first_step: int = 0, sync: bool = True, log_sumq: bool = True, intake_every: int = 0):
        s = self._s()
        sumq = np.zeros(steps, np.int64) if log_sumq else None
        rows = (steps + intake_every - 1) // intake_every if intake_every > 0 else 0
        intake = np.zeros((rows, self.state.n), np.int32) if rows else None
        self.L.orc_run(C.byref(s), steps, first_step, 1 if sync else 0, _p(sumq) if log_sumq else None,
                       _p(intake) if rows else None, intake_every)
        return sumq, intake

    def sum_quality(self) -> int:
        s = self._s(); return int(self.L.orc_sum_quality(C.byref(s)))

    @classmethod
    def create_environment(cls, n_seeds: int, rows: int = 100, cols: int = 100, seed: int = 0, replicate: int = 0) -> np.ndarray:
        q = np.zeros((rows, cols), np.int32)
        cls.lib().orc_create_environment(_p(q), rows, cols, cols, int(n_seeds), C.c_uint64(seed), C.c_uint32(replicate))
        return q

    def time_focal(self, ids: np.ndarray):
        ids = np.ascontiguousarray(ids, np.int32)
        s = self._s()
        nt = C.c_int(0)
        secs = self.L.orc_time_focal(C.byref(s), _p(ids), len(ids), C.byref(nt))
        return float(secs), int(nt.value)

    # ---- scalar helpers
    @classmethod
    def correctangle(cls, a: float) -> float:
        return float(cls.lib().orc_correctangle(float(a)))

    @classmethod
    def averageturn(cls, v, comb: bool) -> float:
        v = np.ascontiguousarray(v, np.float64)
        return float(cls.lib().orc_averageturn(_p(v), len(v), 1 if comb else 0))

    @classmethod
    def getangle(cls, focal_xy, focal_dir, pos_xy, w=2000, h=2000, seed=0, replicate=0, step=0, fish_id=0, k=0) -> float:
        f = np.ascontiguousarray(focal_xy, np.float64); p = np.ascontiguousarray(pos_xy, np.float64)
        return float(cls.lib().orc_getangle(_p(f), float(focal_dir), _p(p), w, h, seed, replicate, step, fish_id, k))

    @classmethod
    def correct_coords(cls, xy, w=2000, h=2000):
        a = np.array(xy, np.float64)
        cls.lib().orc_correct_coords(_p(a), w, h)
        return a


def draw(seed: int, replicate: int, step: int, fish_id: int, stream: int, a: int, b: int, k0: int = 0) -> int:
    """Pure-Python restatement of orc_rng.h (Philox4x32-10 + libstdc++'s Lemire mapping); small cases only."""
    rng = (b - a) + 1
    k = k0
    while True:
        word = philox4x32_10((fish_id & 0xFFFFFFFF, step & 0xFFFFFFFF, (stream & 0xFF) | ((k << 8) & 0xFFFFFF00),
                              (replicate ^ (k & 0xFF000000)) & 0xFFFFFFFF), (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))[0]
        prod = word * rng
        low = prod & 0xFFFFFFFF
        thr = ((1 << 32) - rng) % rng
        if low >= thr:
            return a + (prod >> 32)
        k += 1


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c3 ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF
        k0 = (k0 + 0x9E3779B9) & 0xFFFFFFFF
        k1 = (k1 + 0xBB67AE85) & 0xFFFFFFFF
    return c0, c1, c2, c3


class RefLib:
    """The unmodified reference compiled for capacity `num` (process-global state: one school at a time)."""

    _cache: dict = {}

    @staticmethod
    def available(num: int, suffix: str = "") -> bool:
        return os.path.exists(os.path.join(REF_DIR, f"libfishref_{num}{suffix}.so"))

    def __new__(cls, num: int, suffix: str = ""):
        """suffix "_O0": the same translation unit compiled without optimisation, as the reference ships it
        (fishmod_compile.sh:1); timing only"""
        if (num, suffix) in cls._cache:
            return cls._cache[(num, suffix)]
        self = super().__new__(cls)
        path = os.path.join(REF_DIR, f"libfishref_{num}{suffix}.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run oracle/build_ref.sh {num} where /root/reference exists")
        L = C.CDLL(path)
        L.ref_rng.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32]
        if hasattr(L, "ref_create_environment"):
            L.ref_create_environment.argtypes = [C.c_int, C.c_uint64, C.c_uint32]
        L.ref_run.restype = C.c_double
        L.ref_run.argtypes = [C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
        L.ref_time_focal.restype = C.c_double
        L.ref_time_focal.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_correctangle.restype = C.c_double
        L.ref_correctangle.argtypes = [C.c_double]
        L.ref_getangle.restype = C.c_double
        L.ref_getangle.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_int]
        L.ref_averageturn.restype = C.c_double
        L.ref_averageturn.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_in_zone_all.argtypes = [C.c_int, C.c_double, C.c_void_p, C.c_int]
        L.ref_in_zone_ids.argtypes = [C.c_int, C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.ref_social_all.argtypes = [C.c_int, C.c_void_p, C.c_int]
        L.ref_social_ids.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.ref_initialize.argtypes = [C.c_int, C.c_double, C.c_uint]
        L.ref_sizeof_individuals.restype = C.c_long
        self.L = L
        self.num = num
        assert L.ref_num() == num
        cls._cache[(num, suffix)] = self
        return self

    def set_world(self, w: int, h: int, food_cell: int = 20):
        self.L.ref_set_world(w, h, h // food_cell, w // food_cell)

    def rng(self, seed: int, replicate: int, step: int):
        self.L.ref_rng(seed, replicate, step)

    def set_state(self, st: State):
        st = st.normalise()
        rc = self.L.ref_set_state(st.n, _p(st.coords), _p(st.dir), _p(st.speed), _p(st.speed_factor), _p(st.lag),
                                  _p(st.sample_rate), _p(st.food_intake))
        assert rc == 0, "state larger than the compiled capacity"

    def get_state(self) -> State:
        st = State.empty(self.L.ref_n())
        self.L.ref_get_state(_p(st.coords), _p(st.dir), _p(st.speed), _p(st.speed_factor), _p(st.lag),
                             _p(st.sample_rate), _p(st.food_intake))
        return st

    def set_quality(self, q: np.ndarray):
        q = np.ascontiguousarray(q, np.int32)
        assert self.L.ref_set_quality(_p(q), q.shape[0], q.shape[1]) == 0

    def get_quality(self, rows: int, cols: int) -> np.ndarray:
        q = np.zeros((rows, cols), np.int32)
        self.L.ref_get_quality(_p(q), rows, cols)
        return q

    def get_environment(self, directory: str) -> int:
        return self.L.ref_get_environment(directory.encode())

    def create_environment(self, n_seeds: int, seed: int, replicate: int = 0, rows: int = 100, cols: int = 100) -> np.ndarray:
        self.L.ref_create_environment(int(n_seeds), C.c_uint64(seed), C.c_uint32(replicate))
        return self.get_quality(rows, cols)

    def sum_array(self) -> int:
        return int(self.L.ref_sum_array())

    def initialize(self, n: int, speed: float, srand_seed: int):
        assert self.L.ref_initialize(n, float(speed), srand_seed) == 0

    def in_zone(self, fov: int, r: float, ids=None, threads: int = 8) -> np.ndarray:
        if ids is None:
            out = np.zeros(self.L.ref_n(), np.int32)
            self.L.ref_in_zone_all(fov, float(r), _p(out), threads)
        else:
            ids = np.ascontiguousarray(ids, np.int32)
            out = np.zeros(len(ids), np.int32)
            self.L.ref_in_zone_ids(fov, float(r), _p(ids), len(ids), _p(out), threads)
        return out

    def social(self, fov: int = 330, ids=None, threads: int = 8) -> np.ndarray:
        if ids is None:
            out = np.zeros(self.L.ref_n(), np.float64)
            self.L.ref_social_all(fov, _p(out), threads)
        else:
            ids = np.ascontiguousarray(ids, np.int32)
            out = np.zeros(len(ids), np.float64)
            self.L.ref_social_ids(fov, _p(ids), len(ids), _p(out), threads)
        return out

    def get_speed_all(self, threads: int = 8) -> np.ndarray:
        out = np.zeros(self.L.ref_n(), np.float64)
        self.L.ref_get_speed_all(_p(out), threads)
        return out

    def correctangle(self, a): return float(self.L.ref_correctangle(float(a)))

    def getangle(self, focal_xy, focal_dir, pos_xy, fish_id=0) -> float:
        f = np.ascontiguousarray(focal_xy, np.float64); p = np.ascontiguousarray(pos_xy, np.float64)
        return float(self.L.ref_getangle(_p(f), float(focal_dir), _p(p), fish_id))

    def averageturn(self, v, comb: bool) -> float:
        v = np.ascontiguousarray(v, np.float64)
        return float(self.L.ref_averageturn(_p(v), len(v), 1 if comb else 0))

    def correct_coords(self, xy):
        a = np.array(xy, np.float64); self.L.ref_correct_coords(_p(a)); return a

    def move(self): self.L.ref_move()
    def sample(self): self.L.ref_sample()
    def feed(self, fish_id: int): self.L.ref_feed(int(fish_id))

    def move_focal(self, fish_id: int) -> np.ndarray:
        out = np.zeros(5, np.float64)
        self.L.ref_move_focal(int(fish_id), _p(out))
        return out  # dir, x, y, speed_factor, lag

    def run(self, steps: int, first_step: int = 0, log_sumq: bool = True, intake_every: int = 0):
        n = self.L.ref_n()
        sumq = np.zeros(steps, np.int32) if log_sumq else None
        rows = (steps + intake_every - 1) // intake_every if intake_every > 0 else 0
        intake = np.zeros((rows, n), np.int32) if rows else None
        secs = self.L.ref_run(steps, first_step, _p(sumq) if log_sumq else None, _p(intake) if rows else None, intake_every)
        return float(secs), sumq, intake

    def time_focal(self, ids, threads: int) -> float:
        ids = np.ascontiguousarray(ids, np.int32)
        return float(self.L.ref_time_focal(_p(ids), len(ids), threads))

    def census(self) -> list:
        c = (C.c_long * 8)(); self.L.ref_census(c); return list(c)
