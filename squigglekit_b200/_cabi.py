"""ctypes binding of ``libsqk.so`` (include/sqk.h).  This is the only door into the CUDA code.

There is no CPU fallback: if the library has not been built, or no B200 is visible, the
functions here raise -- they never route to ``oracle/`` or to numpy.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsqk.so")

SQK_MEM_HOST, SQK_MEM_DEVICE = 0, 1
SQK_ERR_ARG, SQK_ERR_CUDA, SQK_ERR_NOMEM, SQK_ERR_UNSUPPORTED = -1, -2, -3, -4
SCALE = {"zscale": 0, "medmad": 1, "none": 2}
PRECISION = {"fp64": 0, "fp32": 1}
K_STATS, K_DTW, K_SEG_FSM, K_DTW_LB, K_DTW_WIN, K_COUNT = 0, 1, 2, 3, 4, 5
DTW_PLAN = {"auto": 0, "single_pass": 1, "two_pass": 2}

# sqk_hit: {int32 start; int32 end; double dist}
HIT_DTYPE = np.dtype([("start", "<i4"), ("end", "<i4"), ("dist", "<f8")], align=True)
assert HIT_DTYPE.itemsize == 16


class SqkError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libsqk error {code}: {msg}")
        self.code = code


class MotifParams(C.Structure):
    _fields_ = [("scale_mode", C.c_int32), ("lo", C.c_int32), ("hi", C.c_int32), ("precision", C.c_int32)]


class SegParams(C.Structure):
    _fields_ = [("error", C.c_int32), ("corrector", C.c_int32), ("window", C.c_int32), ("seg_dist", C.c_int32),
                ("std_scale", C.c_double), ("stall_len", C.c_double), ("lim_lo", C.c_int32), ("lim_hi", C.c_int32),
                ("num", C.c_int32), ("max_segs", C.c_int32)]


class AdapterParams(C.Structure):
    _fields_ = [("error", C.c_int32), ("no_err_thresh", C.c_int32), ("corrector", C.c_int32), ("window", C.c_int32),
                ("seg_dist", C.c_int32), ("t_start", C.c_int32), ("t_end", C.c_int32), ("std_scale", C.c_double),
                ("lim_lo", C.c_int32), ("lim_hi", C.c_int32)]


class RollmeanParams(C.Structure):
    _fields_ = [("w", C.c_int32), ("seg_dist", C.c_int32), ("lo_thresh", C.c_int32), ("hi_thresh", C.c_int32),
                ("shift", C.c_int32), ("lim_lo", C.c_int32), ("lim_hi", C.c_int32), ("reserved", C.c_int32),
                ("std_factor", C.c_double)]


class Timing(C.Structure):
    _fields_ = [("launches", C.c_int64 * K_COUNT), ("ms", C.c_double * K_COUNT)]


EXPORTS = [
    "sqk_version", "sqk_last_error", "sqk_ctx_create", "sqk_ctx_destroy", "sqk_ctx_set_stream", "sqk_ctx_reset_stream", "sqk_ctx_sync",
    "sqk_device_count", "sqk_ctx_device_props", "sqk_host_alloc", "sqk_host_free", "sqk_motifseq",
    "sqk_motifseq_trace", "sqk_segmenter", "sqk_segmenter_pa", "sqk_adapter", "sqk_motifseq_f64", "sqk_segmenter_f64", "sqk_ctx_enable_timing", "sqk_ctx_get_timing", "sqk_ctx_set_dtw_lanes", "sqk_ctx_set_chunk_samples", "sqk_ctx_set_dtw_plan", "sqk_ctx_get_plan_counters", "sqk_ctx_get_plan_counters_ex",
    "sqk_ctx_get_launches", "sqk_ctx_set_stats_generation", "sqk_device_alloc", "sqk_device_free", "sqk_ipc_export", "sqk_ipc_open",
    "sqk_ipc_close", "sqk_rollmean", "sqk_tsv_parse", "sqk_tsv_format", "sqk_tsv_heads", "sqk_tsv_format_rows", "sqk_tsv_format_segs", "sqk_score_hits", "sqk_ndtr", "sqk_ctx_set_hit_peers", "sqk_ctx_set_hit_peers_ex", "sqk_ctx_set_flag_peers", "sqk_peer_signal", "sqk_peer_wait",
]

_lib = None


def lib() -> C.CDLL:
    """Load libsqk.so (once).  Raises if it has not been built -- no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `make -C squigglekit_b200/csrc` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). squigglekit_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.sqk_tsv_format.restype = i64
    L.sqk_version.restype = C.c_int
    L.sqk_last_error.restype = C.c_char_p
    L.sqk_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.sqk_ctx_destroy.argtypes = [vp]
    L.sqk_ctx_set_stream.argtypes = [vp, vp]
    L.sqk_ctx_sync.argtypes = [vp]
    L.sqk_ctx_reset_stream.argtypes = [vp]
    L.sqk_device_count.argtypes = [C.POINTER(C.c_int)]
    L.sqk_ctx_device_props.argtypes = [vp, C.POINTER(i64)]
    L.sqk_host_alloc.argtypes = [C.c_uint64, C.POINTER(vp)]
    L.sqk_host_free.argtypes = [vp]
    L.sqk_motifseq.argtypes = [vp, vp, vp, i64, i64, vp, vp, i32, C.POINTER(MotifParams), C.c_int, vp, vp]
    L.sqk_motifseq_trace.argtypes = [vp, vp, i64, vp, i32, C.POINTER(MotifParams), vp, vp, i64, C.POINTER(i64), vp]
    L.sqk_segmenter.argtypes = [vp, vp, vp, i64, i64, C.POINTER(SegParams), C.c_int, vp, vp]
    L.sqk_adapter.argtypes = [vp, vp, vp, i64, i64, C.POINTER(AdapterParams), C.c_int, vp, vp]
    L.sqk_rollmean.argtypes = [vp, vp, vp, i64, i64, C.POINTER(RollmeanParams), C.c_int, vp, vp]
    L.sqk_segmenter_pa.argtypes = [vp, vp, vp, i64, i64, vp, vp, C.POINTER(SegParams), C.c_int, vp, vp]
    L.sqk_motifseq_f64.argtypes = [vp, vp, vp, i64, vp, vp, i32, C.POINTER(MotifParams), C.c_int, vp, vp]
    L.sqk_segmenter_f64.argtypes = [vp, vp, vp, i64, C.POINTER(SegParams), C.c_int, vp, vp]
    L.sqk_ctx_enable_timing.argtypes = [vp, C.c_int]
    L.sqk_ctx_get_timing.argtypes = [vp, C.POINTER(Timing), C.c_int]
    L.sqk_ctx_set_dtw_lanes.argtypes = [vp, C.c_int]
    L.sqk_ctx_set_chunk_samples.argtypes = [vp, i64]
    L.sqk_ctx_set_dtw_plan.argtypes = [vp, C.c_int]
    L.sqk_ctx_get_plan_counters.argtypes = [vp, C.POINTER(i64)]
    L.sqk_ctx_get_plan_counters_ex.argtypes = [vp, C.POINTER(i64)]
    L.sqk_ctx_get_launches.argtypes = [vp, C.POINTER(i64), C.c_int]
    L.sqk_ctx_set_stats_generation.argtypes = [vp, C.c_int]
    L.sqk_device_alloc.argtypes = [vp, C.c_uint64, C.POINTER(vp)]
    L.sqk_device_free.argtypes = [vp, vp]
    L.sqk_ipc_export.argtypes = [vp, vp, C.c_char_p]
    L.sqk_ipc_open.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    L.sqk_ipc_close.argtypes = [vp, vp]
    L.sqk_ctx_set_hit_peers.argtypes = [vp, C.POINTER(vp), C.c_int, i64]
    L.sqk_ctx_set_hit_peers_ex.argtypes = [vp, C.POINTER(vp), C.c_int, i64, i64]
    L.sqk_ctx_set_flag_peers.argtypes = [vp, C.POINTER(vp), C.c_int, C.c_int]
    L.sqk_peer_signal.argtypes = [vp, C.c_uint64]
    L.sqk_peer_wait.argtypes = [vp, C.c_uint64]
    L.sqk_tsv_parse.argtypes = [vp, i64, C.c_int, C.c_int, i64, i64, C.c_int, vp, vp, vp, vp, vp, C.POINTER(i64), C.POINTER(i64)]
    L.sqk_tsv_format.argtypes = [vp, vp, i64, vp, vp, C.c_int, vp, i64]
    L.sqk_tsv_heads.argtypes = [vp, vp, vp, i64, C.c_int, vp, i64]
    L.sqk_tsv_heads.restype = i64
    L.sqk_tsv_format_rows.argtypes = [vp, i64, vp, C.c_int, vp, vp, vp, vp, vp, C.c_int, vp, i64]
    L.sqk_tsv_format_rows.restype = i64
    L.sqk_tsv_format_segs.argtypes = [vp, i64, vp, vp, C.c_int, vp, C.c_int, vp, i64]
    L.sqk_tsv_format_segs.restype = i64
    L.sqk_score_hits.argtypes = [vp, i64, C.c_int, vp, vp, C.c_int, vp, vp, vp]
    L.sqk_score_hits.restype = None
    L.sqk_ndtr.argtypes = [vp, i64, vp]
    L.sqk_ndtr.restype = None
    for name in EXPORTS:
        if name not in ("sqk_version", "sqk_last_error", "sqk_tsv_format", "sqk_tsv_heads", "sqk_tsv_format_rows", "sqk_tsv_format_segs",
                        "sqk_score_hits", "sqk_ndtr"):
            getattr(L, name).restype = C.c_int
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise SqkError(rc, lib().sqk_last_error().decode("utf-8", "replace"))
