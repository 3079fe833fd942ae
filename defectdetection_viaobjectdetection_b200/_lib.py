"""ctypes binding of libpaut.so (include/paut.h).  No CPU fallback: a missing library is an error."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpaut.so")

PAUT_MAX_OUTPUTS = 12
F32, BF16, I64 = 0, 1, 2
PRECISION = {"fp32": 0, "bf16": 1}
KINDS = {"msc": 0, "msc_n": 1, "conv1d_msc": 2, "ssd": 3, "enhanced": 4, "two_stage": 5,
         "msc_legacy": 6, "improved": 7, "hybrid": 8, "complex": 9}

# every symbol include/paut.h declares (tests/test_abi.py checks the header against this list)
SYMBOLS = (
    "paut_abi_version", "paut_ctx_create", "paut_ctx_destroy", "paut_last_error",
    "paut_ctx_set_workspace_limit", "paut_model_create", "paut_model_destroy", "paut_model_set_tensor",
    "paut_model_finalize", "paut_model_num_keys", "paut_model_key", "paut_forward", "paut_postprocess",
    "paut_window_gather", "paut_window_table_host", "paut_ctx_launch_count", "paut_ctx_profile_begin",
    "paut_ctx_profile_end", "paut_op_linear", "paut_debug_mma", "paut_debug_stage",
    "paut_difference_matrix", "paut_metrics_match", "paut_metrics_confusion",
    "paut_json_load_host", "paut_json_free", "paut_json_last_error", "paut_json_num_beams", "paut_json_beam_info",
    "paut_json_beam_copy_host", "paut_json_scan_key", "paut_json_scan_copy_host", "paut_json_beam_status", "paut_group_nonzero",
)


class ModelCfg(C.Structure):
    _fields_ = [
        ("signal_length", C.c_int32), ("hidden_sizes", C.c_int32 * 3), ("num_heads", C.c_int32),
        ("d_model", C.c_int32), ("num_classes", C.c_int32), ("num_layers", C.c_int32),
        ("dim_feedforward", C.c_int32), ("precision", C.c_int32), ("reserved", C.c_int32 * 6),
    ]


class Metrics(C.Structure):
    """paut_metrics"""
    _fields_ = [("tp", C.c_int64), ("fp", C.c_int64), ("fn", C.c_int64), ("tn", C.c_int64),
                ("sum_iou", C.c_double), ("sum_position_error", C.c_double)]


class Outputs(C.Structure):
    _fields_ = [("slot", C.c_void_p * PAUT_MAX_OUTPUTS)]


# numpy view of paut_detection (48 bytes)
DETECTION = np.dtype([
    ("set_index", "<i4"), ("position", "<i4"), ("cls", "<i4"), ("start_index", "<i4"), ("end_index", "<i4"),
    ("start", "<f4"), ("end", "<f4"), ("score", "<f4"), ("uncertainty", "<f4"), ("anomaly", "<f4"),
    ("confidence", "<f8"),
], align=True)
assert DETECTION.itemsize == 48


class PautError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libpaut error {code}: {message}")
        self.code = code


_lib = None


def load():
    """Load libpaut.so once.  Raises if the extension has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m defectdetection_viaobjectdetection_b200.build` "
            "(the package has no CPU or PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64
    sig = {
        "paut_abi_version": (i32, []),
        "paut_ctx_create": (i32, [i32, vp, C.POINTER(vp)]),
        "paut_ctx_destroy": (None, [vp]),
        "paut_last_error": (C.c_char_p, [vp]),
        "paut_ctx_set_workspace_limit": (i32, [vp, u64]),
        "paut_model_create": (i32, [vp, i32, C.POINTER(ModelCfg), C.POINTER(vp)]),
        "paut_model_destroy": (None, [vp]),
        "paut_model_set_tensor": (i32, [vp, C.c_char_p, vp, i32, C.POINTER(i64), i32]),
        "paut_model_finalize": (i32, [vp]),
        "paut_model_num_keys": (i32, [vp]),
        "paut_model_key": (i32, [vp, i32, C.POINTER(C.c_char_p), C.POINTER(i64), C.POINTER(i32)]),
        "paut_forward": (i32, [vp, vp, i32, i64, i64, i64, C.POINTER(Outputs)]),
        "paut_postprocess": (i32, [vp, C.POINTER(Outputs), i64, i64, i64, C.c_double, vp, vp]),
        "paut_window_gather": (i32, [vp, vp, i32, i64, i64, i64, vp, i64, i64, vp, i32]),
        "paut_window_table_host": (i32, [i32, i64, i64, C.POINTER(C.c_int32), i32]),
        "paut_ctx_launch_count": (i64, [vp]),
        "paut_ctx_profile_begin": (i32, [vp]),
        "paut_ctx_profile_end": (i32, [vp, C.c_char_p, i64]),
        "paut_op_linear": (i32, [vp, vp, i64, i32, vp, vp, i32, vp, i32, i32]),
        "paut_debug_mma": (i32, [vp, i32, i32, i32, i32, i32, vp]),
        "paut_debug_stage": (i32, [vp, i32, vp, i32, i64, i64, i64, vp]),
        "paut_difference_matrix": (i32, [vp, vp, i32, vp, i64, i64, i64, C.c_double, vp, vp, vp]),
        "paut_metrics_match": (i32, [vp, i32, vp, vp, i64, i64, vp, vp, C.c_double, vp]),
        "paut_metrics_confusion": (i32, [vp, vp, vp, i64, C.c_double, i32, vp]),
        "paut_group_nonzero": (i32, [vp, vp, i32, i64, i64, i64, vp]),
        "paut_json_load_host": (i32, [C.c_char_p, C.POINTER(vp)]),
        "paut_json_free": (None, [vp]),
        "paut_json_last_error": (C.c_char_p, []),
        "paut_json_num_beams": (i32, [vp]),
        "paut_json_beam_info": (i32, [vp, i32, C.POINTER(C.c_char_p), C.POINTER(i64), C.POINTER(i64)]),
        "paut_json_beam_copy_host": (i32, [vp, i32, vp, vp, vp, vp]),
        "paut_json_scan_key": (C.c_char_p, [vp, i32, i64]),
        "paut_json_scan_copy_host": (i64, [vp, i32, i64, vp, i64]),
        "paut_json_beam_status": (i32, [vp, i32, C.POINTER(C.c_char_p)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.paut_abi_version() != 1:
        raise ImportError("libpaut.so ABI version mismatch; rebuild the extension")
    _lib = lib
    return lib


def check(code, ctx_handle=None):
    if code == 0:
        return
    msg = load().paut_last_error(ctx_handle)
    raise PautError(code, msg.decode() if msg else "unknown error")
