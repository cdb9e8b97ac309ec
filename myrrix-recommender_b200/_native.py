"""ctypes binding of include/myrrix_als.h (libmyrrix_als.so).

There is no CPU fallback: if the CUDA library is missing or no B200 is visible the
product path raises.  Nothing here imports oracle/.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# MYRRIX_ALS_LIB: development override (A/B builds of the same C ABI), never a fallback
LIB_PATH = os.environ.get("MYRRIX_ALS_LIB") or os.path.join(HERE, "libmyrrix_als.so")

ALS_OK, ALS_E_SINGULAR, ALS_E_NONFINITE, ALS_E_CUDA, ALS_E_NCCL, ALS_E_OOM, ALS_E_ARG, \
    ALS_E_UNSUPPORTED, ALS_E_STATE = range(9)
ALS_KERNEL_AUTO, ALS_KERNEL_SIMT, ALS_KERNEL_TCGEN05 = 0, 1, 2
STATUS_NAMES = ["ALS_OK", "ALS_E_SINGULAR", "ALS_E_NONFINITE", "ALS_E_CUDA", "ALS_E_NCCL",
                "ALS_E_OOM", "ALS_E_ARG", "ALS_E_UNSUPPORTED", "ALS_E_STATE"]


class AlsConfig(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("features", C.c_int32), ("alpha", C.c_double),
                ("lambda_", C.c_double), ("reconstruct_r", C.c_int32),
                ("loss_ignores_unspecified", C.c_int32), ("singularity_threshold", C.c_double),
                ("device", C.c_int32), ("kernel", C.c_int32)]


class AlsInfo(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("features", C.c_int32),
                ("padded_features", C.c_int32), ("kernel", C.c_int32), ("n_users", C.c_int64),
                ("n_items", C.c_int64), ("nnz", C.c_int64), ("device_bytes", C.c_int64),
                ("sm_count", C.c_int32), ("world_size", C.c_int32), ("rank", C.c_int32)]


class AlsTimings(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("launches", C.c_int32), ("gramian_ms", C.c_double),
                ("update_x_ms", C.c_double), ("update_y_ms", C.c_double),
                ("exchange_ms", C.c_double), ("n_half_x", C.c_int32), ("n_half_y", C.c_int32),
                ("fp64_retry_rows", C.c_int64), ("fp64_resolve_rows", C.c_int64)]


# Every symbol include/myrrix_als.h declares: (name, restype, argtypes)
_H = C.c_void_p
_i64p, _i32p, _f32p, _f64p = (C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_float),
                              C.POINTER(C.c_double))
SYMBOLS = [
    ("als_abi_version", C.c_int, []),
    ("als_config_default", C.c_int, [C.POINTER(AlsConfig)]),
    ("als_create", C.c_int, [C.POINTER(AlsConfig), C.POINTER(_H)]),
    ("als_destroy", C.c_int, [_H]),
    ("als_set_stream", C.c_int, [_H, C.c_void_p]),
    ("als_set_interactions", C.c_int, [_H, C.c_int64, C.c_int64, _i64p, _i32p, _f32p]),
    ("als_set_interactions_by_column", C.c_int, [_H, _i64p, _i32p, _f32p]),
    ("als_set_interactions_device", C.c_int, [_H, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                              C.c_void_p]),
    ("als_set_present_empty_rows", C.c_int, [_H, C.c_int32, _i32p, C.c_int64]),
    ("als_set_y", C.c_int, [_H, _f32p]),
    ("als_set_x", C.c_int, [_H, _f32p]),
    ("als_half_x", C.c_int, [_H]),
    ("als_half_y", C.c_int, [_H]),
    ("als_iterate", C.c_int, [_H, C.c_int32]),
    ("als_probe", C.c_int, [_H, _i32p, C.c_int32, _i32p, C.c_int32, _f64p]),
    ("als_get_x", C.c_int, [_H, _f32p]),
    ("als_get_y", C.c_int, [_H, _f32p]),
    ("als_get_factor_block", C.c_int, [_H, C.c_int32, C.c_int64, C.c_int64, _f32p]),
    ("als_get_rows", C.c_int, [_H, C.c_int32, _i32p, C.c_int32, _f32p]),
    ("als_gramian", C.c_int, [_H, C.c_int32, _f64p]),
    ("als_sync", C.c_int, [_H]),
    ("als_last_error", C.c_char_p, [_H]),
    ("als_singular_rank", C.c_int, [_H]),
    ("als_get_info", C.c_int, [_H, C.POINTER(AlsInfo)]),
    ("als_profile_enable", C.c_int, [_H, C.c_int32]),
    ("als_get_timings", C.c_int, [_H, C.POINTER(AlsTimings), C.c_int32]),
    ("als_synth_interactions", C.c_int, [_H, C.c_int64, C.c_int64, C.c_int32, C.c_uint64,
                                         C.c_double]),
    ("als_synth_interactions_powerlaw", C.c_int, [_H, C.c_int64, C.c_int64, C.c_double, C.c_int32, C.c_uint64,
                                                  C.c_double]),
    ("als_synth_y0", C.c_int, [_H, C.c_uint64]),
    ("als_get_interactions", C.c_int, [_H, _i64p, _i32p, _f32p]),
    ("als_get_interactions_by_column", C.c_int, [_H, _i64p, _i32p, _f32p]),
    ("als_get_interaction_rows", C.c_int, [_H, C.c_int32, C.c_int64, C.c_int64, _i64p, _i32p, _f32p,
                                           C.c_int64]),
    ("als_call", C.c_int, [_H, _i32p, C.c_int32, _i32p, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_int32,
                           _i32p, _f64p]),
    ("als_set_fold_in_state", C.c_int, [_H, C.c_int32, _f64p, _f64p, _i32p, C.c_double]),
    ("als_fold_in", C.c_int, [_H, _i32p, _i32p, _f32p, C.c_int64]),
    ("als_recommend", C.c_int, [_H, _i32p, C.c_int32, C.c_int32, C.c_int32, _i32p, C.c_int32, _i32p, _f32p,
                                _i32p]),
    ("als_recommend_batch", C.c_int, [_H, _i32p, C.c_int64, C.c_int32, C.c_int32, _i32p, _f32p, _i32p]),
    ("als_top_n", C.c_int, [_H, C.c_int32, _f32p, C.c_int32, _i32p, C.c_int32, C.c_int32, _i32p, _f32p,
                            _i32p]),
    ("als_comm_unique_id_size", C.c_int, []),
    ("als_comm_get_unique_id", C.c_int, [C.c_void_p]),
    ("als_comm_init", C.c_int, [_H, C.c_int32, C.c_int32, C.c_void_p]),
]

_lib = None


def load():
    """Load libmyrrix_als.so or raise -- never falls back to a CPU implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "CUDA library %s is missing: run `python myrrix-recommender_b200/build.py` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.als_abi_version() != 1:
        raise RuntimeError("ABI version mismatch")
    _lib = lib
    return lib
