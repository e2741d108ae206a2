"""ctypes loader for libfishabm.so (include/fishabm.h).  Fails loudly when the CUDA library is missing:
there is no CPU path behind this package."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FISHABM_LIB") or os.path.join(HERE, "libfishabm.so")   # FISHABM_LIB: a development build to A/B against

OK, E_ARG, E_CUDA, E_STATE, E_NOMEM, E_COMM = 0, -1, -2, -3, -4, -5
MODE_CELL_FP32, MODE_BRUTE_FP64, MODE_ENSEMBLE = 0, 1, 2
STREAM_CHOICE, STREAM_ERROR, STREAM_RANDOMWALK, STREAM_QUALITY, STREAM_SEED, STREAM_SAMPLE, STREAM_INIT = range(7)


class ModelParams(C.Structure):
    """fishabm_model_params: the reference's literals as run-time parameters"""
    _fields_ = [("lag_after_feed", C.c_int32), ("sample_rate_fed", C.c_int32), ("sample_rate_unfed", C.c_int32),
                ("sample_draw_max", C.c_int32), ("error_range", C.c_int32), ("randomwalk_range", C.c_int32),
                ("alone_align_min", C.c_int32), ("reserved", C.c_int32), ("fed_speed_factor", C.c_double),
                ("weight_align", C.c_double), ("weight_attract", C.c_double)]


class Params(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("n", C.c_int32), ("w", C.c_int32), ("h", C.c_int32),
                ("mode", C.c_int32), ("device", C.c_int32), ("replicates", C.c_int32), ("rolling", C.c_int32),
                ("seed", C.c_uint64), ("speed_norm", C.c_double), ("rank", C.c_int32), ("nranks", C.c_int32),
                ("n_global", C.c_int32), ("capacity", C.c_int32), ("halo_capacity", C.c_int32),
                ("migrant_capacity", C.c_int32), ("replicate0", C.c_uint32), ("force_slab", C.c_int32), ("model", ModelParams)]


class SlabInfo(C.Structure):
    _fields_ = [("rank", C.c_int32), ("nranks", C.c_int32), ("left", C.c_int32), ("right", C.c_int32),
                ("col0", C.c_int32), ("col1", C.c_int32), ("owned", C.c_int32), ("held", C.c_int32),
                ("capacity", C.c_int32), ("halo_capacity", C.c_int32), ("migrant_capacity", C.c_int32),
                ("message_bytes", C.c_int64)]


class PassStats(C.Structure):
    _fields_ = [("focal", C.c_uint64), ("candidates", C.c_uint64), ("interacting", C.c_uint64),
                ("rechecked", C.c_uint64)]


class SlabEndpoint(C.Structure):
    _fields_ = [("ipc_handle", C.c_ubyte * 64), ("raw_pointer", C.c_uint64), ("bytes", C.c_int64),
                ("pid", C.c_int32), ("device", C.c_int32), ("rank", C.c_int32), ("pad", C.c_int32)]


# every symbol include/fishabm.h declares (tests check the library exports each of them)
SYMBOLS = [
    "fishabm_params_default", "fishabm_create", "fishabm_destroy", "fishabm_last_error", "fishabm_set_state",
    "fishabm_get_state", "fishabm_initialize", "fishabm_set_quality", "fishabm_get_quality", "fishabm_sum_quality",
    "fishabm_load_quality_csv", "fishabm_in_zone", "fishabm_zone_counts", "fishabm_social", "fishabm_get_speed",
    "fishabm_move", "fishabm_sample", "fishabm_step", "fishabm_flush", "fishabm_feed", "fishabm_run_logged", "fishabm_get_step",
    "fishabm_set_step", "fishabm_draw", "fishabm_last_step_ms", "fishabm_last_neighbour_ms",
    "fishabm_last_pass_stats", "fishabm_launch_count", "fishabm_enable_stats", "fishabm_averageturn",
    "fishabm_correctangle", "fishabm_correct_coords",
    "fishabm_slab_get_info", "fishabm_slab_unique_id", "fishabm_slab_connect", "fishabm_slab_begin_step",
    "fishabm_slab_buffers", "fishabm_slab_end_step", "fishabm_slab_copy", "fishabm_slab_get_endpoint",
    "fishabm_slab_attach", "fishabm_step_enqueue", "fishabm_wait", "fishabm_set_state_owned", "fishabm_get_state_owned",
    "fishabm_create_environment", "fishabm_check_fp64", "fishabm_enable_phase_timing", "fishabm_last_phase_ms",
    "fishabm_model_params_default", "fishabm_set_replicate_models",
]
UNIQUE_ID_BYTES = 128

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a).  fish_abm_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, u32, dbl = C.c_void_p, C.c_int32, C.c_uint32, C.c_double
    L.fishabm_params_default.argtypes = [C.POINTER(Params)]
    L.fishabm_model_params_default.argtypes = [C.POINTER(ModelParams)]
    L.fishabm_set_replicate_models.argtypes = [vp, C.POINTER(ModelParams), i32]
    L.fishabm_create.argtypes = [C.POINTER(Params), C.POINTER(vp)]
    L.fishabm_destroy.argtypes = [vp]
    L.fishabm_destroy.restype = None
    L.fishabm_last_error.argtypes = [vp]
    L.fishabm_last_error.restype = C.c_char_p
    L.fishabm_set_state.argtypes = [vp] + [vp] * 7
    L.fishabm_get_state.argtypes = [vp] + [vp] * 7
    L.fishabm_initialize.argtypes = [vp, dbl]
    L.fishabm_set_quality.argtypes = [vp, vp, i32, i32, i32]
    L.fishabm_get_quality.argtypes = [vp, i32, vp, i32, i32, i32]
    L.fishabm_sum_quality.argtypes = [vp, vp]
    L.fishabm_load_quality_csv.argtypes = [C.c_char_p, vp, i32, i32, i32, C.POINTER(i32), C.POINTER(i32)]
    L.fishabm_in_zone.argtypes = [vp, i32, dbl, vp]
    L.fishabm_zone_counts.argtypes = [vp, vp]
    L.fishabm_social.argtypes = [vp, vp, vp, vp, vp]
    L.fishabm_get_speed.argtypes = [vp, vp]
    L.fishabm_move.argtypes = [vp]
    L.fishabm_sample.argtypes = [vp]
    L.fishabm_step.argtypes = [vp, i32]
    L.fishabm_flush.argtypes = [vp]
    L.fishabm_feed.argtypes = [vp, vp, i32]
    L.fishabm_run_logged.argtypes = [vp, i32, vp, vp, i32]
    L.fishabm_get_step.argtypes = [vp, C.POINTER(u32)]
    L.fishabm_set_step.argtypes = [vp, u32]
    L.fishabm_draw.argtypes = [vp, i32, u32, u32, u32, u32, i32, i32, C.POINTER(i32)]
    L.fishabm_last_step_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.fishabm_last_neighbour_ms.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(i32)]
    L.fishabm_last_pass_stats.argtypes = [vp, C.POINTER(PassStats)]
    L.fishabm_launch_count.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.fishabm_enable_stats.argtypes = [vp, i32]
    L.fishabm_enable_phase_timing.argtypes = [vp, i32]
    L.fishabm_last_phase_ms.argtypes = [vp, C.POINTER(C.c_float * 4), C.POINTER(i32)]
    L.fishabm_averageturn.argtypes = [vp, i32, i32]
    L.fishabm_averageturn.restype = dbl
    L.fishabm_correctangle.argtypes = [dbl]
    L.fishabm_correctangle.restype = dbl
    L.fishabm_correct_coords.argtypes = [vp, i32, i32]
    L.fishabm_correct_coords.restype = None
    L.fishabm_slab_get_info.argtypes = [vp, C.POINTER(SlabInfo)]
    L.fishabm_slab_unique_id.argtypes = [vp, i32]
    L.fishabm_slab_connect.argtypes = [vp, vp, i32]
    L.fishabm_slab_begin_step.argtypes = [vp]
    L.fishabm_slab_buffers.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_int64)]
    L.fishabm_slab_end_step.argtypes = [vp]
    L.fishabm_slab_copy.argtypes = [vp, vp, C.c_int64]
    L.fishabm_slab_get_endpoint.argtypes = [vp, C.POINTER(SlabEndpoint)]
    L.fishabm_slab_attach.argtypes = [vp, C.POINTER(SlabEndpoint), C.POINTER(SlabEndpoint)]
    L.fishabm_step_enqueue.argtypes = [vp, i32]
    L.fishabm_wait.argtypes = [vp]
    L.fishabm_check_fp64.argtypes = [vp, vp, i32, vp, vp, vp]
    L.fishabm_create_environment.argtypes = [C.c_uint64, u32, i32, vp, i32, i32, i32]
    L.fishabm_set_state_owned.argtypes = [vp, i32] + [vp] * 8
    L.fishabm_get_state_owned.argtypes = [vp, C.POINTER(i32)] + [vp] * 8
    _lib = L
    return L
