"""ctypes front-end of ``libsqk_oracle.so`` (oracle/sqk_oracle.c).  TEST INFRASTRUCTURE ONLY.

Function names mirror the reference call sites they stand in for:
``dtw_subsequence`` -> mlpy.dtw_subsequence as called at MotifSeq.py:437;
``zscale`` / ``medmad`` -> MotifSeq.py:186-200; ``get_segs`` -> segmenter.py:399-470.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsqk_oracle.so")
_lib = None


class _Hit(C.Structure):
    _fields_ = [("start", C.c_int32), ("end", C.c_int32), ("dist", C.c_double)]


HIT_DTYPE = np.dtype([("start", "<i4"), ("end", "<i4"), ("dist", "<f8")], align=True)


class _SegCfg(C.Structure):
    _fields_ = [
        ("error", C.c_int32), ("corrector", C.c_int32), ("window", C.c_int32), ("seg_dist", C.c_int32),
        ("std_scale", C.c_double), ("stall_len", C.c_double),
    ]


@dataclass
class SegCfg:
    """get_segs parameters with the reference's argparse defaults (segmenter.py:67-90)."""
    error: int = 5
    corrector: int = 50
    window: int = 150
    seg_dist: int = 50
    std_scale: float = 0.75
    stall_len: float = 0.25

    def c(self) -> _SegCfg:
        return _SegCfg(self.error, self.corrector, self.window, self.seg_dist, self.std_scale, self.stall_len)


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (make -C oracle).  Building the checker is not using it."""
    src = os.path.join(_HERE, "sqk_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class _AdapterCfg(C.Structure):
    _fields_ = [("error", C.c_int32), ("no_err_thresh", C.c_int32), ("corrector", C.c_int32), ("window", C.c_int32),
                ("seg_dist", C.c_int32), ("t_start", C.c_int32), ("t_end", C.c_int32), ("std_scale", C.c_double)]


@dataclass
class AdapterCfg:
    """The constants hard-coded in the slow5 branch of dRNA_segmenter.py (:81-106)."""
    error: int = 5
    no_err_thresh: int = 2500
    corrector: int = 1200
    window: int = 100
    seg_dist: int = 1200
    t_start: int = 1000
    t_end: int = 5000
    std_scale: float = 0.8

    def c(self) -> _AdapterCfg:
        return _AdapterCfg(self.error, self.no_err_thresh, self.corrector, self.window, self.seg_dist, self.t_start,
                           self.t_end, self.std_scale)


class _RollmeanCfg(C.Structure):
    _fields_ = [("w", C.c_int32), ("seg_dist", C.c_int32), ("lo_thresh", C.c_int32), ("hi_thresh", C.c_int32),
                ("shift", C.c_int32), ("std_factor", C.c_double)]


@dataclass
class RollmeanCfg:
    """The constants of the TSV branch of dRNA_segmenter.py (:81 `# w = 2000`, :287, :292-294, :320)."""
    w: int = 2000
    seg_dist: int = 1500
    lo_thresh: int = 2000
    hi_thresh: int = 200000
    shift: int = 1000
    std_factor: float = 0.5

    def c(self) -> _RollmeanCfg:
        return _RollmeanCfg(self.w, self.seg_dist, self.lo_thresh, self.hi_thresh, self.shift, self.std_factor)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        dp, ip, lp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
        L.orc_dtw_subsequence.argtypes = [dp, C.c_int, dp, C.c_int, dp, C.POINTER(_Hit), ip, ip, lp]
        L.orc_dtw_subsequence_rolling.argtypes = [dp, C.c_int, dp, C.c_int, C.POINTER(_Hit), dp]
        L.orc_np_sum.argtypes = [dp, C.c_int64]
        L.orc_np_sum.restype = C.c_double
        L.orc_np_median.argtypes = [dp, C.c_int64]
        L.orc_np_median.restype = C.c_double
        L.orc_zscale.argtypes = [dp, C.c_int64, dp, dp]
        L.orc_medmad.argtypes = [dp, C.c_int64, dp, dp]
        L.orc_get_segs.argtypes = [dp, C.c_int64, C.POINTER(_SegCfg), ip, C.c_int, dp]
        L.orc_motifseq_batch.argtypes = [C.c_void_p, lp, C.c_int64, dp, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_int, C.c_void_p, ip]
        L.orc_segmenter_batch.argtypes = [C.c_void_p, lp, C.c_int64, C.POINTER(_SegCfg), C.c_int, C.c_int,
                                          C.c_int, C.c_int, C.c_int, ip, ip]
        L.orc_convert_to_pa.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_double, dp]
        L.orc_convert_to_pa.restype = None
        L.orc_segmenter_batch_pa.argtypes = [C.c_void_p, lp, C.c_int64, dp, dp, C.POINTER(_SegCfg), C.c_int, C.c_int,
                                             C.c_int, C.c_int, C.c_int, ip, ip]
        L.orc_motifseq_batch_f64.argtypes = [dp, lp, C.c_int64, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.c_void_p, ip]
        L.orc_segmenter_batch_f64.argtypes = [dp, lp, C.c_int64, C.POINTER(_SegCfg), C.c_int, C.c_int, C.c_int, C.c_int,
                                              C.c_int, ip, ip]
        L.orc_adapter_seg.argtypes = [dp, C.c_int64, C.POINTER(_AdapterCfg), ip, dp]
        L.orc_adapter_batch.argtypes = [C.c_void_p, lp, C.c_int64, C.POINTER(_AdapterCfg), C.c_int, C.c_int, C.c_int, ip, ip]
        L.orc_rollmean_seg.argtypes = [dp, C.c_int64, C.POINTER(_RollmeanCfg), ip, dp]
        L.orc_rollmean_batch.argtypes = [C.c_void_p, lp, C.c_int64, C.POINTER(_RollmeanCfg), C.c_int, C.c_int, C.c_int, ip, ip]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _lp(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def max_threads() -> int:
    return int(lib().orc_max_threads())


def dtw_subsequence(x, y, want_cost: bool = True):
    """mlpy.dtw_subsequence(x, y) -> (dist, cost[n,m], (px, py)) -- restated, see sqk_oracle.c."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    n, m = x.size, y.size
    cost = np.empty((n, m), dtype=np.float64)
    px = np.empty(n + m, dtype=np.int32)
    py = np.empty(n + m, dtype=np.int32)
    k = C.c_int64(0)
    h = _Hit()
    rc = lib().orc_dtw_subsequence(_dp(x), n, _dp(y), m, _dp(cost), C.byref(h), _ip(px), _ip(py), C.byref(k))
    if rc:
        raise ValueError("orc_dtw_subsequence failed (empty input?)")
    assert py[0] == h.start and py[k.value - 1] == h.end
    return h.dist, (cost if want_cost else None), (px[:k.value].astype(np.int64), py[:k.value].astype(np.int64))


def dtw_subsequence_rolling(x, y, want_last_row: bool = False):
    """Independent rolling-column implementation -> (dist, start, end[, last_row])."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    h = _Hit()
    last = np.empty(y.size, dtype=np.float64) if want_last_row else None
    rc = lib().orc_dtw_subsequence_rolling(_dp(x), x.size, _dp(y), y.size, C.byref(h),
                                           _dp(last) if want_last_row else None)
    if rc:
        raise ValueError("orc_dtw_subsequence_rolling failed")
    return (h.dist, h.start, h.end, last) if want_last_row else (h.dist, h.start, h.end)


def np_sum(a) -> float:
    a = np.ascontiguousarray(a, dtype=np.float64)
    return float(lib().orc_np_sum(_dp(a), a.size))


def np_median(a) -> float:
    a = np.ascontiguousarray(a, dtype=np.float64)
    return float(lib().orc_np_median(_dp(a), a.size))


def zscale(sig):
    """sklearn.preprocessing.scale(sig) restated in C -> (scaled float64, mean, std)."""
    v = np.array(sig, dtype=np.float64)
    mean, sd = C.c_double(), C.c_double()
    if lib().orc_zscale(_dp(v), v.size, C.byref(mean), C.byref(sd)):
        raise ValueError("empty signal")
    return v, mean.value, sd.value


def medmad(sig):
    """MotifSeq.py:192-200 restated in C -> (scaled float64, med, mad)."""
    v = np.array(sig, dtype=np.float64)
    med, mad = C.c_double(), C.c_double()
    if lib().orc_medmad(_dp(v), v.size, C.byref(med), C.byref(mad)):
        raise ValueError("empty signal")
    return v, med.value, mad.value


def get_segs(sig, cfg: SegCfg = SegCfg(), max_segs: int = 64, want_thresholds: bool = False):
    """segmenter.get_segs(sig, args) restated -> list of [start, end] or False."""
    v = np.ascontiguousarray(sig, dtype=np.float64)
    out = np.zeros(2 * max_segs, dtype=np.int32)
    thr = np.zeros(4, dtype=np.float64)
    c = cfg.c()
    n = lib().orc_get_segs(_dp(v), v.size, C.byref(c), _ip(out), max_segs, _dp(thr))
    if n > max_segs:
        raise OverflowError(f"{n} segments > max_segs={max_segs}")
    segs = [[int(out[2 * i]), int(out[2 * i + 1])] for i in range(n)] if n > 0 else False
    return (segs, thr) if want_thresholds else segs


SCALE_MODES = {"zscale": 0, "medmad": 1, "none": 2}


def motifseq_batch(signals, offsets, model, lo=0, hi=1200, scale="zscale", full_matrix=True, n_threads=0):
    """Per read: scale_outliers -> normalise -> dtw_subsequence.  -> (hits[HIT_DTYPE], n_kept)."""
    signals = np.ascontiguousarray(signals, dtype=np.int16)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    model = np.ascontiguousarray(model, dtype=np.float64)
    n = offsets.size - 1
    hits = np.zeros(n, dtype=HIT_DTYPE)
    kept = np.zeros(n, dtype=np.int32)
    rc = lib().orc_motifseq_batch(signals.ctypes.data, _lp(offsets), n, _dp(model), model.size, lo, hi,
                                  SCALE_MODES[scale], int(full_matrix), n_threads, hits.ctypes.data, _ip(kept))
    if rc:
        raise RuntimeError("orc_motifseq_batch failed")
    return hits, kept


def segmenter_batch(signals, offsets, cfg: SegCfg = SegCfg(), lim_lo=0, lim_hi=900, num=0, max_segs=16,
                    n_threads=0):
    """Per read: sig[:Num] -> scale_outliers -> get_segs.  -> (segs[n,max_segs,2], n_segs[n])."""
    signals = np.ascontiguousarray(signals, dtype=np.int16)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n = offsets.size - 1
    segs = np.zeros((n, max_segs, 2), dtype=np.int32)
    nsegs = np.zeros(n, dtype=np.int32)
    c = cfg.c()
    rc = lib().orc_segmenter_batch(signals.ctypes.data, _lp(offsets), n, C.byref(c), lim_lo, lim_hi, num,
                                   max_segs, n_threads, _ip(segs), _ip(nsegs))
    if rc:
        raise RuntimeError("orc_segmenter_batch failed")
    return segs, nsegs


def convert_to_pa(sig, offset: float, raw_unit: float):
    """np.round(convert_to_pA_numpy(sig, digitisation, range, offset), 2) with raw_unit = range / digitisation."""
    sig = np.ascontiguousarray(sig, dtype=np.int16)
    out = np.empty(sig.size, dtype=np.float64)
    lib().orc_convert_to_pa(sig.ctypes.data, sig.size, float(offset), float(raw_unit), _dp(out))
    return out


def segmenter_batch_pa(signals, offsets, pa_offset, pa_scale, cfg: SegCfg = SegCfg(), lim_lo=0, lim_hi=900, num=0,
                       max_segs=16, n_threads=0):
    """fast5 default path: per read pA conversion -> sig[:Num] -> scale_outliers -> get_segs."""
    signals = np.ascontiguousarray(signals, dtype=np.int16)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    pa_offset = np.ascontiguousarray(pa_offset, dtype=np.float64)
    pa_scale = np.ascontiguousarray(pa_scale, dtype=np.float64)
    n = offsets.size - 1
    segs = np.zeros((n, max_segs, 2), dtype=np.int32)
    nsegs = np.zeros(n, dtype=np.int32)
    c = cfg.c()
    rc = lib().orc_segmenter_batch_pa(signals.ctypes.data, _lp(offsets), n, _dp(pa_offset), _dp(pa_scale), C.byref(c),
                                      lim_lo, lim_hi, num, max_segs, n_threads, _ip(segs), _ip(nsegs))
    if rc:
        raise RuntimeError("orc_segmenter_batch_pa failed")
    return segs, nsegs


def motifseq_batch_f64(signals, offsets, model, lo=0, hi=1200, scale="zscale", full_matrix=False, n_threads=0):
    """float64 signals (the reference's `-s` path on a TSV of floats): scale_outliers -> normalise -> DTW."""
    signals = np.ascontiguousarray(signals, dtype=np.float64)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    model = np.ascontiguousarray(model, dtype=np.float64)
    n = offsets.size - 1
    hits = np.zeros(n, dtype=HIT_DTYPE)
    kept = np.zeros(n, dtype=np.int32)
    rc = lib().orc_motifseq_batch_f64(_dp(signals), _lp(offsets), n, _dp(model), model.size, lo, hi, SCALE_MODES[scale],
                                      int(full_matrix), n_threads, hits.ctypes.data, _ip(kept))
    if rc:
        raise RuntimeError("orc_motifseq_batch_f64 failed")
    return hits, kept


def segmenter_batch_f64(signals, offsets, cfg: SegCfg = SegCfg(), lim_lo=0, lim_hi=900, num=0, max_segs=16, n_threads=0):
    signals = np.ascontiguousarray(signals, dtype=np.float64)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n = offsets.size - 1
    segs = np.zeros((n, max_segs, 2), dtype=np.int32)
    nsegs = np.zeros(n, dtype=np.int32)
    c = cfg.c()
    rc = lib().orc_segmenter_batch_f64(_dp(signals), _lp(offsets), n, C.byref(c), lim_lo, lim_hi, num, max_segs, n_threads,
                                       _ip(segs), _ip(nsegs))
    if rc:
        raise RuntimeError("orc_segmenter_batch_f64 failed")
    return segs, nsegs


def adapter_seg(sig, cfg: AdapterCfg = AdapterCfg(), want_thresholds: bool = False):
    """dRNA_segmenter.py slow5 branch on one post-outlier signal -> [start, end] of the first segment or None."""
    v = np.ascontiguousarray(sig, dtype=np.float64)
    out = np.zeros(2, dtype=np.int32)
    thr = np.zeros(3, dtype=np.float64)
    c = cfg.c()
    got = lib().orc_adapter_seg(_dp(v), v.size, C.byref(c), _ip(out), _dp(thr))
    if got < 0:
        raise RuntimeError("orc_adapter_seg failed")
    seg = [int(out[0]), int(out[1])] if got else None
    return (seg, thr) if want_thresholds else seg


def adapter_batch(signals, offsets, cfg: AdapterCfg = AdapterCfg(), lim_lo=0, lim_hi=1200, n_threads=0):
    """Per read: scale_outliers -> adapter_seg.  -> (segs[n,2], found[n])."""
    signals = np.ascontiguousarray(signals, dtype=np.int16)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n = offsets.size - 1
    segs = np.zeros((n, 2), dtype=np.int32)
    found = np.zeros(n, dtype=np.int32)
    c = cfg.c()
    rc = lib().orc_adapter_batch(signals.ctypes.data, _lp(offsets), n, C.byref(c), lim_lo, lim_hi, n_threads, _ip(segs), _ip(found))
    if rc:
        raise RuntimeError("orc_adapter_batch failed")
    return segs, found


def rollmean_seg(sig, cfg: RollmeanCfg = RollmeanCfg(), want_thresholds: bool = False):
    """dRNA_segmenter.py TSV branch (:272-326) on one post-outlier signal -> [x, y] of the first qualifying segment
    (already shifted by -1000 as the reference prints it) or None."""
    v = np.ascontiguousarray(sig, dtype=np.float64)
    out = np.zeros(2, dtype=np.int32)
    thr = np.zeros(3, dtype=np.float64)
    c = cfg.c()
    got = lib().orc_rollmean_seg(_dp(v), v.size, C.byref(c), _ip(out), _dp(thr))
    if got < 0:
        raise RuntimeError("orc_rollmean_seg failed")
    seg = [int(out[0]), int(out[1])] if got else None
    return (seg, thr) if want_thresholds else seg


def rollmean_batch(signals, offsets, cfg: RollmeanCfg = RollmeanCfg(), lim_lo=0, lim_hi=1200, n_threads=0):
    """Per read: scale_outliers -> rollmean_seg.  -> (segs[n,2], found[n])."""
    signals = np.ascontiguousarray(signals, dtype=np.int16)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n = offsets.size - 1
    segs = np.zeros((n, 2), dtype=np.int32)
    found = np.zeros(n, dtype=np.int32)
    c = cfg.c()
    rc = lib().orc_rollmean_batch(signals.ctypes.data, _lp(offsets), n, C.byref(c), lim_lo, lim_hi, n_threads, _ip(segs), _ip(found))
    if rc:
        raise RuntimeError("orc_rollmean_batch failed")
    return segs, found
