"""ctypes binding of libdsb200.so (include/dsb200.h).  There is no fallback: if the CUDA library is
missing or a call fails, the caller gets an exception."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdsb200.so")

DSB_MAX_OPS = 8
ABI_VERSION = 3
NOTE_NEGZERO = 1

# dsb_dtype
NONE, F32, F64, I8, U8, I16, U16, I32, U32, I64, U64 = range(11)
_NP_TO_DSB = {
    np.dtype("float32"): F32, np.dtype("float64"): F64, np.dtype("int8"): I8, np.dtype("uint8"): U8,
    np.dtype("int16"): I16, np.dtype("uint16"): U16, np.dtype("int32"): I32, np.dtype("uint32"): U32,
    np.dtype("int64"): I64, np.dtype("uint64"): U64, np.dtype("bool"): U8,
}

# dsb_op
OP_COUNT, OP_ANY, OP_SUM, OP_MAX32, OP_MIN32, OP_MAX64, OP_MIN64 = 1, 2, 3, 4, 5, 6, 7
OP_MAXROW, OP_MINROW, OP_ARGMAX32, OP_ARGMIN32, OP_MATCHROW64 = 8, 9, 10, 11, 12

LINE_ANY, LINE_COUNT, LINE_SUM, LINE_MAX, LINE_MIN, LINE_MEAN, LINE_MEAN_2STAGE = 1, 2, 3, 4, 5, 6, 7
AA2_SUM, AA2_COUNT, AA2_MIN, AA2_FIRST, AA2_LAST, AA2_ARGMIN, AA2_ARGMAX = 1, 2, 3, 4, 5, 6, 7


class View(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("x_log", C.c_int32), ("y_log", C.c_int32),
                ("sx", C.c_double), ("tx", C.c_double), ("sy", C.c_double), ("ty", C.c_double),
                ("xmin", C.c_double), ("xmax", C.c_double), ("ymin", C.c_double), ("ymax", C.c_double)]


class Base(C.Structure):
    _fields_ = [("op", C.c_int32), ("val_dtype", C.c_int32), ("val", C.c_void_p),
                ("chk_dtype", C.c_int32), ("chk", C.c_void_p), ("agg", C.c_void_p), ("aux", C.c_void_p)]


class Plan(C.Structure):
    _fields_ = [("nops", C.c_int32), ("ops", Base * DSB_MAX_OPS), ("cat", C.c_void_p),
                ("cat_dtype", C.c_int32), ("ncat", C.c_int32), ("notes", C.c_void_p)]


class LineLayout(C.Structure):
    _fields_ = [("x_line_stride", C.c_int64), ("y_line_stride", C.c_int64), ("value_per_vertex", C.c_int32),
                ("plot_start", C.c_int32),
                # ragged layouts: start index per row of the flat vertex arrays (NULL x_starts = dense)
                ("x_starts", C.c_void_p), ("y_starts", C.c_void_p), ("y1_starts", C.c_void_p),
                ("x_flat_len", C.c_int64), ("y_flat_len", C.c_int64), ("y1_flat_len", C.c_int64)]


class Dsb200Error(RuntimeError):
    pass


def dsb_dtype(np_dtype) -> int:
    try:
        return _NP_TO_DSB[np.dtype(np_dtype)]
    except KeyError:
        raise TypeError(f"unsupported column dtype {np_dtype!r}") from None


_p, _i32, _i64, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double

_SIGNATURES = {
    "dsb_abi_version": ([], C.c_int),
    "dsb_last_error": ([], C.c_char_p),
    "dsb_launch_count": ([], C.c_int64),
    "dsb_last_kernel": ([], C.c_char_p),
    "dsb_configure": ([C.c_char_p, _i64], C.c_int),
    "dsb_init_canvas": ([_i32, _p, _i64, _p], C.c_int),
    "dsb_points": ([C.POINTER(View), _p, _p, _i32, _i64, _i64, C.POINTER(Plan), _p], C.c_int),
    "dsb_points_priv": ([C.POINTER(View), _p, _p, _i32, _i64, _i64, C.POINTER(Plan), _i32, _p, _p, _p], C.c_int),
    "dsb_points_routed_scratch_bytes": ([C.POINTER(View), _i64], C.c_int64),
    "dsb_points_routed": ([C.POINTER(View), _p, _p, _i32, _i64, _i64, C.POINTER(Plan), _p, _i64, _p], C.c_int),
    "dsb_routed_configure": ([_i64], C.c_int),
    "dsb_points_match32_scratch_bytes": ([C.POINTER(View)], C.c_int64),
    "dsb_points_minmax_rest": ([C.POINTER(View), _p, _p, _i32, _i64, _i64, _p, _i32, _p, _i32, _p, _p, _i64, _p], C.c_int),
    "dsb_points_argminmax_rest": ([C.POINTER(View), _p, _p, _i32, _i64, _i64, _p, _i32, _p, _i32, _p, _i64, _p], C.c_int),
    "dsb_points_match32": ([C.POINTER(View), _p, _p, _i32, _i64, _i64, _p, _i32, _p, _i32, _p, _p, _i64, _p], C.c_int),
    "dsb_lines_configure": ([C.c_int], C.c_int),
    "dsb_points_views": ([_p, _i32, _i32, _i32, _f64, _f64, _f64, _f64, _p, _p, _i32, _i64, _i64, C.POINTER(Plan), _i64, _p],
                         C.c_int),
    "dsb_bounds": ([_p, _i32, _i64, _p, _p], C.c_int),
    "dsb_decode_minmax": ([_p, _i32, _i32, _p, _i64, _p], C.c_int),
    "dsb_decode_arg": ([_p, _i32, _i32, _i64, _p, _p, _i64, _p], C.c_int),
    "dsb_gather_rows": ([_p, _i64, _i64, _p, _i32, _p, _i64, _p], C.c_int),
    "dsb_finish_minrow": ([_p, _i64, _p], C.c_int),
    "dsb_finalize_mean": ([_p, _p, _p, _i64, _p], C.c_int),
    "dsb_finalize_sum": ([_p, _p, _p, _i64, _p], C.c_int),
    "dsb_finalize_sum_counted": ([_p, _p, _p, _i64, _p], C.c_int),
    "dsb_points_count16": ([C.POINTER(View), _p, _p, _i32, _i64, _i64, C.POINTER(Plan), _p, _i64, _p], C.c_int),
    "dsb_composite": ([_p, _p, C.c_uint32, _i64, _i32, _p, _p], C.c_int),
    "dsb_spread_image": ([_p, _i32, _i32, _p, _i32, _i32, _p, _p], C.c_int),
    "dsb_spread_array": ([_p, _i32, _i32, _i32, _i32, _p, _i32, _i32, _p, _p], C.c_int),
    "dsb_density": ([_p, _i32, _i32, _i32, _i32, _i32, _p, _p], C.c_int),
    "dsb_lines_aa2": ([_p, _p, _p, _i32, _i64, _i64, _p, _i64, _p, _i32, _i32, _i32, C.c_double, _p, _p, _p, _i64, _p], C.c_int),
    "dsb_lines_axis1_plan": ([C.POINTER(View), _p, _p, _i32, _i64, _i64, C.POINTER(LineLayout), _i64, C.POINTER(Plan), _p],
                             C.c_int),
    "dsb_areas_plan": ([C.POINTER(View), _p, _p, _p, _i32, _i64, _i64, C.POINTER(LineLayout), _i64, C.POINTER(Plan), _p],
                       C.c_int),
    "dsb_lines_axis1": ([C.POINTER(View), _p, _p, _i32, _i64, _i64, C.POINTER(LineLayout), _p, _i32, _i32, _f64, _p, _p,
                         _p], C.c_int),
    "dsb_lines_axis1_cat": ([C.POINTER(View), _p, _p, _i32, _i64, _i64, C.POINTER(LineLayout), _p, _i32, _i32, _f64, _p, _p,
                             _p, _i32, _i32, _p], C.c_int),
}

_lib = None


def declared_symbols():
    """Entry points declared in include/dsb200.h (parsed, so the header stays the source of truth)."""
    import re
    hdr = os.path.join(os.path.dirname(HERE), "include", "dsb200.h")
    with open(hdr) as f:
        text = f.read()
    return sorted(set(re.findall(r"\b(dsb_[a-z0-9_]+)\s*\(", text)) - {"dsb_status"})


def lib():
    """Load libdsb200.so; raises if it has not been built (no CPU or PyTorch fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Dsb200Error(
                f"{LIB_PATH} is missing: build it with `python -m datashader_b200._build` "
                "(datashader_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (argtypes, restype) in _SIGNATURES.items():
            fn = getattr(L, name, None)
            if fn is None:
                continue
            fn.argtypes = argtypes
            fn.restype = restype
        if L.dsb_abi_version() != ABI_VERSION:
            raise Dsb200Error("libdsb200.so ABI version mismatch")
        _lib = L
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().dsb_last_error()
        raise Dsb200Error(f"{what or 'dsb200'} failed ({rc}): {msg.decode() if msg else ''}")
