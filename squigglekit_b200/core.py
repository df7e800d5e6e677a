"""Host-side mirror of the reference's per-read functions, batched over reads and executed by
libsqk's sm_100a kernels through the C ABI (include/sqk.h).

Reference call sites replaced:
  * ``dtw_subsequence(model[name], sig)`` + the normalisation block in front of it
    (MotifSeq.py:180-209, 437-439)  ->  :meth:`Context.motifseq`
  * ``get_segs(sig, args)`` + ``sig[:Num]`` + ``scale_outliers`` (segmenter.py:124-128, 399-470)
    ->  :meth:`Context.segmenter`;  ``test_segs`` (segmenter.py:473-494) -> :func:`test_segs`

Inputs are either numpy arrays (host mode: the library streams them to the GPU in chunks) or
torch CUDA tensors (device mode: kernels are enqueued on torch's current stream, nothing is
copied).  torch is plumbing here -- device memory and streams -- not the compute path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _cabi
from ._cabi import HIT_DTYPE, SqkError  # noqa: F401


@dataclass
class SegConfig:
    """get_segs / test_segs parameters; defaults are the reference's argparse defaults
    (segmenter.py:65-94)."""
    error: int = 5
    corrector: int = 50
    window: int = 150
    seg_dist: int = 50
    std_scale: float = 0.75
    stall_len: float = 0.25
    lim_low: int = 0
    lim_hi: int = 900
    Num: int = 0
    stall: bool = False
    stall_start: int = 300
    gap: bool = False
    gap_dist: int = 3000
    max_segs: int = 16


@dataclass
class AdapterConfig:
    """The constants the slow5 branch of dRNA_segmenter.py hard-codes (dRNA_segmenter.py:82-106, :333)."""
    error: int = 5
    no_err_thresh: int = 2500
    corrector: int = 1200
    window: int = 100
    seg_dist: int = 1200
    t_start: int = 1000
    t_end: int = 5000
    std_scale: float = 0.8
    lim_low: int = 0
    lim_hi: int = 1200


@dataclass
class RollmeanConfig:
    """The constants of the TSV branch of dRNA_segmenter.py (:81 ``# w = 2000``, :287, :292-294, :320, :333)."""
    w: int = 2000
    seg_dist: int = 1500
    lo_thresh: int = 2000
    hi_thresh: int = 200000
    shift: int = 1000
    std_factor: float = 0.5
    lim_low: int = 0
    lim_hi: int = 1200


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def pinned_empty(shape, dtype) -> np.ndarray:
    """numpy array backed by page-locked host memory (sqk_host_alloc); host-mode calls overlap
    the H2D copy with compute only when the signal buffer is pinned."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = C.c_void_p()
    _cabi.check(_cabi.lib().sqk_host_alloc(max(n, 1), C.byref(p)))
    buf = (C.c_char * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED[arr.__array_interface__["data"][0]] = p.value
    return arr


_PINNED: dict[int, int] = {}


def pinned_free(arr: np.ndarray) -> None:
    addr = arr.__array_interface__["data"][0]
    p = _PINNED.pop(addr, None)
    if p is not None:
        _cabi.check(_cabi.lib().sqk_host_free(p))


class Context:
    """One libsqk context = one GPU.  Not thread-safe (one per host thread)."""

    def __init__(self, device: int = 0):
        self._lib = _cabi.lib()
        h = C.c_void_p()
        _cabi.check(self._lib.sqk_ctx_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)
        props = (C.c_int64 * 5)()
        _cabi.check(self._lib.sqk_ctx_device_props(self._h, props))
        self.n_sms, self.smem_optin, self.clock_khz, self.l2_bytes, self.cc = (int(v) for v in props)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.sqk_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- instrumentation -----------------------------------------------------------------
    def enable_timing(self, on: bool = True):
        _cabi.check(self._lib.sqk_ctx_enable_timing(self._h, int(on)))

    def timing(self, reset: bool = True) -> dict:
        t = _cabi.Timing()
        _cabi.check(self._lib.sqk_ctx_get_timing(self._h, C.byref(t), int(reset)))
        names = ["stats", "dtw", "seg_fsm", "dtw_lb", "dtw_win"]
        return {n: {"launches": int(t.launches[i]), "ms": float(t.ms[i])} for i, n in enumerate(names)}

    def set_dtw_lanes(self, lanes: int):
        _cabi.check(self._lib.sqk_ctx_set_dtw_lanes(self._h, int(lanes)))

    def set_dtw_plan(self, plan: str = "auto"):
        """How exact (fp64) MotifSeq requests run: "single_pass" (float64 recurrence over every column),
        "two_pass" (float32 lower-bound scan + float64 windows, same results bit for bit) or "auto"."""
        _cabi.check(self._lib.sqk_ctx_set_dtw_plan(self._h, _cabi.DTW_PLAN[plan]))

    def plan_counters(self) -> dict:
        """Two-pass diagnostics of the most recent launch (first model): exact windows run, reads re-run in full."""
        out = (C.c_int64 * 4)()
        _cabi.check(self._lib.sqk_ctx_get_plan_counters_ex(self._h, out))
        return {"windows": int(out[0]), "fallback_reads": int(out[1]), "second_attempt_windows": int(out[2])}

    def launches(self, reset: bool = True) -> int:
        """Kernels this context launched since the last reset (counted by the library at every launch site)."""
        out = C.c_int64(0)
        _cabi.check(self._lib.sqk_ctx_get_launches(self._h, C.byref(out), int(reset)))
        return int(out.value)

    def set_stats_generation(self, gen: int = 0):
        """0 = automatic, 1 = first-generation statistics kernel only (A/B measurements, tests); same results."""
        _cabi.check(self._lib.sqk_ctx_set_stats_generation(self._h, int(gen)))

    # ---- multi-GPU publication of hit records (see include/sqk.h; squigglekit_b200.dist.PeerGather drives these) ------
    def device_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        _cabi.check(self._lib.sqk_device_alloc(self._h, int(nbytes), C.byref(p)))
        return int(p.value)

    def device_free(self, ptr: int):
        _cabi.check(self._lib.sqk_device_free(self._h, C.c_void_p(ptr)))

    def ipc_export(self, ptr: int) -> bytes:
        buf = C.create_string_buffer(64)
        _cabi.check(self._lib.sqk_ipc_export(self._h, C.c_void_p(ptr), buf))
        return buf.raw

    def ipc_open(self, handle: bytes) -> int:
        p = C.c_void_p()
        _cabi.check(self._lib.sqk_ipc_open(self._h, C.create_string_buffer(handle, 64), C.byref(p)))
        return int(p.value)

    def ipc_close(self, ptr: int):
        _cabi.check(self._lib.sqk_ipc_close(self._h, C.c_void_p(ptr)))

    def set_hit_peers(self, peers, first_record: int = 0, capacity_records: int = 0):
        """Arm the NEXT device-mode motifseq call to store its records into the peers' gathered buffers as well (P2P); calls
        after it do not publish until this is called again.  capacity_records: size of a gathered buffer (bounds check)."""
        arr = (C.c_void_p * max(1, len(peers)))(*[C.c_void_p(p) for p in peers])
        _cabi.check(self._lib.sqk_ctx_set_hit_peers_ex(self._h, arr, len(peers), int(first_record), int(capacity_records)))

    def set_flag_peers(self, flag_arrays, my_rank: int):
        arr = (C.c_void_p * max(1, len(flag_arrays)))(*[C.c_void_p(p) for p in flag_arrays])
        _cabi.check(self._lib.sqk_ctx_set_flag_peers(self._h, arr, len(flag_arrays), int(my_rank)))

    def peer_signal(self, value: int):
        self._use_torch_stream()
        _cabi.check(self._lib.sqk_peer_signal(self._h, int(value)))

    def peer_wait(self, value: int):
        self._use_torch_stream()
        _cabi.check(self._lib.sqk_peer_wait(self._h, int(value)))

    def set_chunk_samples(self, samples: int):
        """Host mode: samples per in-flight chunk of the copy/compute pipeline (0 = default)."""
        _cabi.check(self._lib.sqk_ctx_set_chunk_samples(self._h, int(samples)))

    def sync(self):
        _cabi.check(self._lib.sqk_ctx_sync(self._h))

    def _use_torch_stream(self):
        import torch
        _cabi.check(self._lib.sqk_ctx_set_stream(self._h, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))

    # ---- MotifSeq ------------------------------------------------------------------------
    @staticmethod
    def _pack_models(models):
        if isinstance(models, np.ndarray) and models.ndim == 1:
            models = [models]
        vecs = [np.ascontiguousarray(m, dtype=np.float64).reshape(-1) for m in models]
        if not vecs:
            raise ValueError("no models")
        offs = np.zeros(len(vecs) + 1, dtype=np.int32)
        np.cumsum([v.size for v in vecs], out=offs[1:])
        return np.concatenate(vecs), offs

    def motifseq(self, signals, offsets, models, scale: str = "medmad", scale_low: int = 0, scale_hi: int = 1200,
                 precision: str = "fp64", max_read_len: int = 0, out=None, want_kept: bool = True):
        """Batched ``scale_outliers -> normalise -> dtw_subsequence`` (MotifSeq.py:180-209,437-439).

        signals: int16 [total_samples]; offsets: int64 [n_reads+1]; models: one float64 vector or a list.
        -> (hits, n_kept): hits is a structured array [n_reads, n_models] with fields start, end,
        dist (host mode) or a uint8 CUDA tensor [n_reads, n_models, 16] holding the same records
        (device mode; see :func:`hits_from_torch`).  start = -1: read empty after outlier removal;
        -2: MAD == 0.
        """
        mvec, moffs = self._pack_models(models)
        params = _cabi.MotifParams(_cabi.SCALE[scale], int(scale_low), int(scale_hi), _cabi.PRECISION[precision])
        n_models = moffs.size - 1
        if _is_torch(signals):
            import torch
            if signals.dtype not in (torch.int16, torch.float64) or offsets.dtype != torch.int64:
                raise TypeError("device mode needs int16 (or float64) signals and int64 offsets")
            if not (signals.is_cuda and offsets.is_cuda and signals.is_contiguous() and offsets.is_contiguous()):
                raise ValueError("device mode needs contiguous CUDA tensors")
            n_reads = offsets.numel() - 1
            dev = signals.device
            hits = out if out is not None else torch.empty((n_reads, n_models, 16), dtype=torch.uint8, device=dev)
            kept = torch.empty(n_reads, dtype=torch.int32, device=dev) if want_kept else None
            self._use_torch_stream()
            if signals.dtype == torch.float64:
                _cabi.check(self._lib.sqk_motifseq_f64(
                    self._h, signals.data_ptr(), offsets.data_ptr(), n_reads, mvec.ctypes.data, moffs.ctypes.data, n_models,
                    C.byref(params), _cabi.SQK_MEM_DEVICE, hits.data_ptr(), kept.data_ptr() if kept is not None else None))
                return hits, kept
            _cabi.check(self._lib.sqk_motifseq(
                self._h, signals.data_ptr(), offsets.data_ptr(), n_reads, int(max_read_len), mvec.ctypes.data,
                moffs.ctypes.data, n_models, C.byref(params), _cabi.SQK_MEM_DEVICE, hits.data_ptr(),
                kept.data_ptr() if kept is not None else None))
            return hits, kept
        signals = np.asarray(signals)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n_reads = offsets.size - 1
        hits = out if out is not None else np.zeros((n_reads, n_models), dtype=HIT_DTYPE)
        kept = np.zeros(n_reads, dtype=np.int32) if want_kept else None
        if signals.dtype.kind == "f":
            # float signal (e.g. pA values from a SquigglePull TSV): the float64 front end, same outputs
            signals = np.ascontiguousarray(signals, dtype=np.float64)
            _cabi.check(self._lib.sqk_motifseq_f64(
                self._h, signals.ctypes.data, offsets.ctypes.data, n_reads, mvec.ctypes.data, moffs.ctypes.data, n_models,
                C.byref(params), _cabi.SQK_MEM_HOST, hits.ctypes.data, kept.ctypes.data if kept is not None else None))
            return hits, kept
        signals = np.ascontiguousarray(signals, dtype=np.int16)
        _cabi.check(self._lib.sqk_motifseq(
            self._h, signals.ctypes.data, offsets.ctypes.data, n_reads, int(max_read_len), mvec.ctypes.data,
            moffs.ctypes.data, n_models, C.byref(params), _cabi.SQK_MEM_HOST, hits.ctypes.data,
            kept.ctypes.data if kept is not None else None))
        return hits, kept

    def motifseq_trace(self, signal, model, scale: str = "medmad", scale_low: int = 0, scale_hi: int = 1200,
                       precision: str = "fp64"):
        """One read, one model: -> (hit record, normalised signal, cost[-1, :]) -- what ``-x`` prints
        (MotifSeq.py:447) and view_region plots (MotifSeq.py:507-509)."""
        signal = np.ascontiguousarray(signal, dtype=np.int16)
        model = np.ascontiguousarray(model, dtype=np.float64)
        params = _cabi.MotifParams(_cabi.SCALE[scale], int(scale_low), int(scale_hi), _cabi.PRECISION[precision])
        last = np.zeros(signal.size, dtype=np.float64)
        norm = np.zeros(signal.size, dtype=np.float64)
        hit = np.zeros(1, dtype=HIT_DTYPE)
        n_out = C.c_int64(0)
        _cabi.check(self._lib.sqk_motifseq_trace(self._h, signal.ctypes.data, signal.size, model.ctypes.data, model.size,
                                                 C.byref(params), last.ctypes.data, norm.ctypes.data, signal.size,
                                                 C.byref(n_out), hit.ctypes.data))
        n = n_out.value
        return hit[0], norm[:n], last[:n]

    # ---- segmenter -----------------------------------------------------------------------
    def segmenter(self, signals, offsets, cfg: SegConfig = SegConfig(), max_read_len: int = 0, pa_offset=None,
                  pa_scale=None):
        """Batched ``sig[:Num] -> scale_outliers -> get_segs`` (segmenter.py:124-128,399-470).

        With ``pa_offset`` / ``pa_scale`` (float64 per read: channel offset and range/digitisation) every read
        is first converted like the reference's fast5 default, ``np.round((raw+offset)*scale, 2)``
        (segmenter.py:345-349), and limits/thresholds apply to the pA values.

        -> (segs int32 [n_reads, max_segs, 2], n_segs int32 [n_reads]); n_segs == 0 is the
        reference's ``False``; n_segs > max_segs means the row was truncated (raise max_segs).
        """
        p = _cabi.SegParams(cfg.error, cfg.corrector, cfg.window, cfg.seg_dist, cfg.std_scale, cfg.stall_len,
                            cfg.lim_low, cfg.lim_hi, cfg.Num, cfg.max_segs)
        if (pa_offset is None) != (pa_scale is None):
            raise ValueError("pa_offset and pa_scale go together")
        use_pa = pa_offset is not None
        if _is_torch(signals):
            import torch
            if signals.dtype not in (torch.int16, torch.float64) or offsets.dtype != torch.int64:
                raise TypeError("device mode needs int16 (or float64) signals and int64 offsets")
            n_reads = offsets.numel() - 1
            dev = signals.device
            segs = torch.zeros((n_reads, cfg.max_segs, 2), dtype=torch.int32, device=dev)
            nsegs = torch.zeros(n_reads, dtype=torch.int32, device=dev)
            self._use_torch_stream()
            if signals.dtype == torch.float64:
                if use_pa:
                    raise ValueError("pa_offset/pa_scale apply to raw int16 signals, not to float signals")
                _cabi.check(self._lib.sqk_segmenter_f64(self._h, signals.data_ptr(), offsets.data_ptr(), n_reads, C.byref(p),
                                                        _cabi.SQK_MEM_DEVICE, segs.data_ptr(), nsegs.data_ptr()))
                return segs, nsegs
            if use_pa:
                po = torch.as_tensor(pa_offset, dtype=torch.float64, device=dev).contiguous()
                ps = torch.as_tensor(pa_scale, dtype=torch.float64, device=dev).contiguous()
                _cabi.check(self._lib.sqk_segmenter_pa(self._h, signals.data_ptr(), offsets.data_ptr(), n_reads,
                                                       int(max_read_len), po.data_ptr(), ps.data_ptr(), C.byref(p),
                                                       _cabi.SQK_MEM_DEVICE, segs.data_ptr(), nsegs.data_ptr()))
                self._keepalive = (po, ps)     # stream-ordered use: keep the tensors alive until the next call
            else:
                _cabi.check(self._lib.sqk_segmenter(self._h, signals.data_ptr(), offsets.data_ptr(), n_reads,
                                                    int(max_read_len), C.byref(p), _cabi.SQK_MEM_DEVICE, segs.data_ptr(),
                                                    nsegs.data_ptr()))
            return segs, nsegs
        signals = np.asarray(signals)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n_reads = offsets.size - 1
        segs = np.zeros((n_reads, cfg.max_segs, 2), dtype=np.int32)
        nsegs = np.zeros(n_reads, dtype=np.int32)
        if signals.dtype.kind == "f":
            if use_pa:
                raise ValueError("pa_offset/pa_scale apply to raw int16 signals, not to float signals")
            signals = np.ascontiguousarray(signals, dtype=np.float64)
            _cabi.check(self._lib.sqk_segmenter_f64(self._h, signals.ctypes.data, offsets.ctypes.data, n_reads, C.byref(p),
                                                    _cabi.SQK_MEM_HOST, segs.ctypes.data, nsegs.ctypes.data))
            return segs, nsegs
        signals = np.ascontiguousarray(signals, dtype=np.int16)
        if use_pa:
            po = np.ascontiguousarray(pa_offset, dtype=np.float64)
            ps = np.ascontiguousarray(pa_scale, dtype=np.float64)
            if po.size != n_reads or ps.size != n_reads:
                raise ValueError("pa_offset / pa_scale need one value per read")
            _cabi.check(self._lib.sqk_segmenter_pa(self._h, signals.ctypes.data, offsets.ctypes.data, n_reads,
                                                   int(max_read_len), po.ctypes.data, ps.ctypes.data, C.byref(p),
                                                   _cabi.SQK_MEM_HOST, segs.ctypes.data, nsegs.ctypes.data))
        else:
            _cabi.check(self._lib.sqk_segmenter(self._h, signals.ctypes.data, offsets.ctypes.data, n_reads, int(max_read_len),
                                                C.byref(p), _cabi.SQK_MEM_HOST, segs.ctypes.data, nsegs.ctypes.data))
        return segs, nsegs


def _adapter(self, signals, offsets, cfg: AdapterConfig = AdapterConfig(), max_read_len: int = 0):
    """Batched dRNA adapter finder (dRNA_segmenter.py:86-176, slow5 branch): per read scale_outliers, threshold
    from kept samples [t_start, t_end), one-sided run detector, first segment.

    -> (segs int32 [n_reads, 2], found int32 [n_reads]); found == 0: the reference prints nothing for that read."""
    p = _cabi.AdapterParams(cfg.error, cfg.no_err_thresh, cfg.corrector, cfg.window, cfg.seg_dist, cfg.t_start, cfg.t_end,
                            cfg.std_scale, cfg.lim_low, cfg.lim_hi)
    if _is_torch(signals):
        import torch
        if signals.dtype != torch.int16 or offsets.dtype != torch.int64:
            raise TypeError("device mode needs int16 signals and int64 offsets")
        n_reads = offsets.numel() - 1
        segs = torch.zeros((n_reads, 2), dtype=torch.int32, device=signals.device)
        found = torch.zeros(n_reads, dtype=torch.int32, device=signals.device)
        self._use_torch_stream()
        _cabi.check(self._lib.sqk_adapter(self._h, signals.data_ptr(), offsets.data_ptr(), n_reads, int(max_read_len),
                                          C.byref(p), _cabi.SQK_MEM_DEVICE, segs.data_ptr(), found.data_ptr()))
        return segs, found
    signals = np.ascontiguousarray(signals, dtype=np.int16)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n_reads = offsets.size - 1
    segs = np.zeros((n_reads, 2), dtype=np.int32)
    found = np.zeros(n_reads, dtype=np.int32)
    _cabi.check(self._lib.sqk_adapter(self._h, signals.ctypes.data, offsets.ctypes.data, n_reads, int(max_read_len),
                                      C.byref(p), _cabi.SQK_MEM_HOST, segs.ctypes.data, found.ctypes.data))
    return segs, found


Context.adapter = _adapter


def _rollmean(self, signals, offsets, cfg: RollmeanConfig = RollmeanConfig(), max_read_len: int = 0):
    """Batched rolling-mean adapter finder (dRNA_segmenter.py:272-326, TSV branch): per read scale_outliers,
    t = rolling(w).mean(), bot = t.mean() - std_factor * t.std(), runs of t < bot, first segment of a plausible length.

    -> (segs int32 [n_reads, 2] = (start - shift, end - shift), found int32 [n_reads]); found == 0: the reference prints
    nothing for that read."""
    p = _cabi.RollmeanParams(cfg.w, cfg.seg_dist, cfg.lo_thresh, cfg.hi_thresh, cfg.shift, cfg.lim_low, cfg.lim_hi, 0,
                             cfg.std_factor)
    if _is_torch(signals):
        import torch
        if signals.dtype != torch.int16 or offsets.dtype != torch.int64:
            raise TypeError("device mode needs int16 signals and int64 offsets")
        n_reads = offsets.numel() - 1
        segs = torch.zeros((n_reads, 2), dtype=torch.int32, device=signals.device)
        found = torch.zeros(n_reads, dtype=torch.int32, device=signals.device)
        self._use_torch_stream()
        _cabi.check(self._lib.sqk_rollmean(self._h, signals.data_ptr(), offsets.data_ptr(), n_reads, int(max_read_len),
                                           C.byref(p), _cabi.SQK_MEM_DEVICE, segs.data_ptr(), found.data_ptr()))
        return segs, found
    signals = np.ascontiguousarray(signals, dtype=np.int16)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n_reads = offsets.size - 1
    segs = np.zeros((n_reads, 2), dtype=np.int32)
    found = np.zeros(n_reads, dtype=np.int32)
    _cabi.check(self._lib.sqk_rollmean(self._h, signals.ctypes.data, offsets.ctypes.data, n_reads, int(max_read_len),
                                       C.byref(p), _cabi.SQK_MEM_HOST, segs.ctypes.data, found.ctypes.data))
    return segs, found


Context.rollmean = _rollmean


def hits_from_torch(hits_u8) -> np.ndarray:
    """uint8 CUDA tensor [n, m, 16] of sqk_hit records -> structured numpy array [n, m]."""
    a = hits_u8.cpu().numpy()
    return a.reshape(a.shape[0], a.shape[1] * 16).view(HIT_DTYPE).reshape(a.shape[0], a.shape[1])


def segs_to_lists(segs: np.ndarray, n_segs: np.ndarray):
    """Batch output -> what get_segs returns per read: list of [start, end] or False."""
    cap = segs.shape[1]
    nl = n_segs.tolist()
    for r, n in enumerate(nl):
        if n < 0:
            raise ValueError(f"read {r} is longer than the max_read_len passed to segmenter()")
        if n > cap:
            raise OverflowError(f"read {r}: {n} segments > max_segs={cap}")
    sl = segs.tolist()
    return [sl[r][:n] if n else False for r, n in enumerate(nl)]


def test_segs(segs, cfg: SegConfig):
    """segmenter.test_segs (segmenter.py:473-494): -k rejects reads whose first segment starts after
    stall_start; -g rejects reads whose second segment starts more than gap_dist after the first ends
    (a single-segment read passes, as the reference's swallowed IndexError makes it)."""
    if not segs:
        return False
    if cfg.stall and segs[0][0] > cfg.stall_start:
        return False
    if cfg.gap and len(segs) > 1 and segs[1][0] > segs[0][1] + cfg.gap_dist:
        return False
    return segs


test_segs.__test__ = False  # not a pytest test
