"""Literal drop-in for the one mlpy symbol the reference imports: ``from mlpy import dtw_subsequence``
(MotifSeq.py:12, called at :437 as ``dist, cost, path = dtw_subsequence(model[name], sig)``).

``install(ctx)`` registers a module named ``mlpy`` whose ``dtw_subsequence(x, y)`` runs on the GPU through
libsqk's float64 entry point (the caller has already normalised ``y``, so no scaling is applied), and returns
the three things MotifSeq uses: ``dist``, ``path[1][0]`` / ``path[1][-1]`` (start / end) and -- for the plotting
code -- an object whose ``cost[-1,]`` raises a clear error (the N x M matrix is never built).  With it the
reference's own ``get_region_multi`` runs unmodified on top of the CUDA kernel; tests/test_dropin_gpu.py does
exactly that.  Per-call overhead makes this a compatibility path: batch with ``Context.motifseq`` for speed.
"""
from __future__ import annotations

import sys
import types

import numpy as np

_WIDE = 40000          # libsqk's outlier window is clamped to +-40000; normalised signal lives well inside


class _NoCostMatrix:
    def __getitem__(self, key):
        raise NotImplementedError("libsqk keeps one DTW column in registers; use Context.motifseq_trace for cost[-1, :]")


def make_dtw_subsequence(ctx):
    def dtw_subsequence(x, y):
        y = np.ascontiguousarray(y, dtype=np.float64)
        if y.size and (np.abs(y).max() >= _WIDE or not np.isfinite(y).all()):
            raise ValueError("dtw_subsequence shim: signal values must be finite and below 40000 in magnitude")
        offsets = np.array([0, y.size], dtype=np.int64)
        hits, _ = ctx.motifseq(y, offsets, np.asarray(x, dtype=np.float64), scale="none", scale_low=-_WIDE, scale_hi=_WIDE,
                               want_kept=False)
        h = hits[0, 0]
        if h["start"] < 0:
            raise ValueError("dtw_subsequence shim: empty signal")
        path = (None, np.array([int(h["start"]), int(h["end"])], dtype=np.int64))
        return h["dist"], _NoCostMatrix(), path
    return dtw_subsequence


def install(ctx):
    """Register the shim as ``sys.modules['mlpy']`` -> the previous entry (or None)."""
    prev = sys.modules.get("mlpy")
    mod = types.ModuleType("mlpy")
    mod.dtw_subsequence = make_dtw_subsequence(ctx)
    sys.modules["mlpy"] = mod
    return prev
