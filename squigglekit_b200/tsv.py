"""SquigglePull signal text, batched: the ``-s`` input of MotifSeq.py / segmenter.py / dRNA_segmenter.py.

    fast5 <TAB> readID [<TAB> digitisation <TAB> offset <TAB> range <TAB> sampling_rate] <TAB> s0 <TAB> s1 ...

is what SquigglePull.py:243-253 prints and what MotifSeq.py:252-298 (signal from column 8) and segmenter.py:179-230
(signal from column 4) read back with one ``float()`` / ``int()`` per field in a Python list comprehension.  Here the
file is read in large binary blocks and ``sqk_tsv_parse`` (libsqk, C + OpenMP over the lines) turns every block into
the int16 batch layout the GPU path consumes -- samples land directly in a pinned buffer, so the following host-mode
call copies them without staging.  ``write_reads`` is the inverse (``sqk_tsv_format``), i.e. SquigglePull's
``print_data`` for raw signal.
"""
from __future__ import annotations

import ctypes as C
import gzip
from dataclasses import dataclass

import numpy as np

from . import _cabi

NO_SIGNAL, NOT_INT16, ALL_ZERO = 1, 2, 4
DEFAULT_PINNED = True       # Reader(pinned=None): page-locked sample buffer (the CPU test-suite, which has no CUDA runtime, clears this)


@dataclass
class Batch:
    """Consecutive lines of the file.  ``signals[offsets[i]:offsets[i+1]]`` are the int16 samples of line i (valid where
    ``status[i] == 0``); ``head(i)`` its leading columns; ``tail_text(i)`` the raw text of its signal columns;
    ``heads(k)`` the first k columns of every line (one C call for the batch)."""
    text: object            # buffer holding the batch's text (memoryview or mmap slice)
    base: int               # address of text[0]
    line_begin: np.ndarray
    sig_begin: np.ndarray
    offsets: np.ndarray
    status: np.ndarray
    signals: np.ndarray
    n: int

    def head(self, i: int):
        end = int(self.sig_begin[i])
        b = bytes(self.text[int(self.line_begin[i]):max(int(self.line_begin[i]), end - 1)])
        if self.status[i] & NO_SIGNAL:
            b = bytes(self.text[int(self.line_begin[i]):int(self.line_begin[i + 1])]).rstrip(b"\r\n")
        return b.decode("utf-8", "replace").split("\t")

    def heads_bytes(self, n_cols: int = 2) -> bytes:
        """The first ``n_cols`` columns of every line, ``c0 <TAB> c1 <NL>`` each (one C call for the batch)."""
        lib = _cabi.lib()
        need = -int(lib.sqk_tsv_heads(self.base, self.line_begin.ctypes.data, self.sig_begin.ctypes.data, self.n, n_cols, None, 0))
        if need <= 0:
            return b""
        out = bytearray(need)
        view = (C.c_char * need).from_buffer(out)
        lib.sqk_tsv_heads(self.base, self.line_begin.ctypes.data, self.sig_begin.ctypes.data, self.n, n_cols, C.addressof(view), need)
        del view
        return bytes(out)

    def heads(self, n_cols: int = 2):
        """-> list of lists: the first ``n_cols`` columns of every line."""
        return [ln.split("\t") for ln in self.heads_bytes(n_cols).decode("utf-8", "replace").split("\n")[:self.n]]

    def tail_text(self, i: int) -> str:
        return bytes(self.text[int(self.sig_begin[i]):int(self.line_begin[i + 1])]).rstrip(b"\r\n").decode("utf-8", "replace")

    def sig(self, i: int) -> np.ndarray:
        return self.signals[int(self.offsets[i]):int(self.offsets[i + 1])]


class _Slot:
    """One batch's worth of parse output: a slice of the (pinned) sample buffer and the per-line arrays."""

    def __init__(self, samples: np.ndarray, max_lines: int):
        self.samples = samples
        self.offsets = np.zeros(max_lines + 1, dtype=np.int64)
        self.line_begin = np.zeros(max_lines + 1, dtype=np.int64)
        self.sig_begin = np.zeros(max_lines, dtype=np.int64)
        self.status = np.zeros(max_lines, dtype=np.int32)


class Reader:
    """Iterate over a SquigglePull TSV as parsed batches of at most ``max_lines`` lines / ``max_samples`` samples.  A
    plain file is memory-mapped and parsed in place (no copy of the text is ever made); a .gz file is streamed through a
    buffer.  The sample buffer is pinned when ``pinned`` (needs the CUDA runtime, i.e. a GPU box).

    ``prefetch`` (plain files): the sample buffer is cut into two slots of ``max_samples // 2`` and a helper thread parses
    the next batch into the free slot while the caller works on the current one (its GPU call, its row formatting) --
    ``sqk_tsv_parse`` runs outside the GIL.  A batch is valid until the iterator is advanced, as without prefetch."""

    def __init__(self, path: str, start_col: int, max_lines: int = 16384, max_samples: int = 96 << 20,
                 block_bytes: int = 128 << 20, pinned=None, n_threads: int = 0, prefetch: bool = True):
        import mmap
        import os
        self.lib = _cabi.lib()
        self.start_col, self.max_lines, self.max_samples = int(start_col), int(max_lines), int(max_samples)
        self.block_bytes, self.n_threads = int(block_bytes), int(n_threads)
        self.mm = None
        self.fh = None
        if path.endswith(".gz"):
            self.fh = gzip.open(path, "rb")
        else:
            self._file = open(path, "rb")
            size = os.fstat(self._file.fileno()).st_size
            if size > 0:
                self.mm = mmap.mmap(self._file.fileno(), 0, access=mmap.ACCESS_READ)
                try:
                    self.mm.madvise(mmap.MADV_SEQUENTIAL)
                except Exception:
                    pass
                self.mm_arr = np.frombuffer(self.mm, dtype=np.uint8)
            self.size = size
        self.prefetch = (bool(prefetch) and self.fh is None and self.max_samples >= 16
                         and os.environ.get("SQK_TSV_PREFETCH", "1") != "0")     # (knob: single buffer, no helper thread)
        n_slots = 2 if self.prefetch else 1
        self.slot_samples = (self.max_samples // n_slots) & ~7 if n_slots > 1 else self.max_samples
        self._pinned = None
        if DEFAULT_PINNED if pinned is None else pinned:
            from .core import pinned_empty
            self._pinned = pinned_empty(self.max_samples, np.int16)
            big = self._pinned
        else:
            big = np.empty(self.max_samples, dtype=np.int16)
        self._slots = [_Slot(big[k * self.slot_samples:(k + 1) * self.slot_samples], self.max_lines) for k in range(n_slots)]
        self.samples = self._slots[0].samples
        self.buf = bytearray()
        self.pos = 0
        self.eof = False

    def _stop_helper(self):
        helper, self._helper = getattr(self, "_helper", None), None
        if helper is not None:
            t, stop, free = helper
            stop.set()
            free.put(None)
            t.join()

    def close(self):
        self._stop_helper()
        if self.fh is not None:
            self.fh.close()
        if self.mm is not None:
            self.mm_arr = None
            try:
                self.mm.close()
            except BufferError:
                pass
            self.mm = None
        if getattr(self, "_file", None) is not None:
            self._file.close()
            self._file = None
        if self._pinned is not None:
            from .core import pinned_free
            pinned_free(self._pinned)
            self._pinned = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _parse(self, addr, avail, final, slot):
        n_lines, consumed = C.c_int64(0), C.c_int64(0)
        rc = self.lib.sqk_tsv_parse(addr, avail, int(final), self.start_col, self.max_lines, slot.samples.size,
                                    self.n_threads, slot.samples.ctypes.data, slot.offsets.ctypes.data,
                                    slot.line_begin.ctypes.data, slot.sig_begin.ctypes.data, slot.status.ctypes.data,
                                    C.byref(n_lines), C.byref(consumed))
        if rc == _cabi.SQK_ERR_NOMEM:
            raise ValueError("a line of the signal file holds more samples than the batch buffer (%d)" % slot.samples.size)
        if rc != 0:
            raise RuntimeError("sqk_tsv_parse failed (%d)" % rc)
        return int(n_lines.value), int(consumed.value)

    def _batch(self, text, base, n, slot):
        return Batch(text, base, slot.line_begin[:n + 1].copy(), slot.sig_begin[:n].copy(), slot.offsets[:n + 1].copy(),
                     slot.status[:n].copy(), slot.samples, n)

    def _iter_mmap_prefetch(self):
        """The memory-mapped file with a helper thread one batch ahead (two slots)."""
        import queue
        import threading
        base = self.mm_arr.ctypes.data
        done, free = queue.Queue(), queue.Queue()
        for k in range(len(self._slots)):
            free.put(k)
        stop = threading.Event()

        def work():
            pos = 0
            try:
                while pos < self.size and not stop.is_set():
                    k = free.get()
                    if k is None:
                        break
                    n, used = self._parse(base + pos, self.size - pos, True, self._slots[k])
                    if n == 0:
                        break
                    done.put((k, pos, n, used))
                    pos += used
                done.put(None)
            except BaseException as e:                      # handed to the consumer, raised there
                done.put(e)

        t = threading.Thread(target=work, name="sqk-tsv-prefetch", daemon=True)
        self._helper = (t, stop, free)                       # close() stops it before the buffers go away
        t.start()
        try:
            while True:
                item = done.get()
                if item is None:
                    return
                if isinstance(item, BaseException):
                    raise item
                k, pos, n, used = item
                mv = memoryview(self.mm)[pos:pos + used]
                try:
                    yield self._batch(mv, base + pos, n, self._slots[k])
                finally:
                    mv.release()
                free.put(k)                                 # the caller is done with that batch: its slot may be refilled
        finally:
            self._stop_helper()

    def _fill(self):
        if self.pos:
            del self.buf[:self.pos]
            self.pos = 0
        chunk = self.fh.read(self.block_bytes)
        if not chunk:
            self.eof = True
        else:
            self.buf += chunk

    def __iter__(self):
        if self.fh is None:                                      # memory-mapped plain file
            if self.mm is None:
                return
            if self.prefetch:
                yield from self._iter_mmap_prefetch()
                return
            pos, base, slot = 0, self.mm_arr.ctypes.data, self._slots[0]
            while pos < self.size:
                n, used = self._parse(base + pos, self.size - pos, True, slot)
                if n == 0:
                    return
                mv = memoryview(self.mm)[pos:pos + used]
                try:
                    yield self._batch(mv, base + pos, n, slot)
                finally:
                    mv.release()
                pos += used
            return
        while True:                                              # streamed (.gz)
            if not self.eof and len(self.buf) - self.pos < self.block_bytes // 2:
                self._fill()
            avail = len(self.buf) - self.pos
            if avail == 0:
                if self.eof:
                    return
                continue
            view = (C.c_char * avail).from_buffer(self.buf, self.pos)
            try:
                n, used = self._parse(C.addressof(view), avail, self.eof, self._slots[0])
                if n == 0:
                    if self.eof:
                        return
                else:
                    mv = memoryview(self.buf)[self.pos:self.pos + used]
                    try:
                        yield self._batch(mv, C.addressof(view), n, self._slots[0])
                    finally:
                        mv.release()
                    self.pos += used
            finally:
                del view
            if n == 0:
                self._fill()                                      # the buffered text ends inside a line: bring in more


def format_reads(heads, signals: np.ndarray, offsets: np.ndarray, n_threads: int = 0) -> bytes:
    """``head <TAB> s0 <TAB> s1 ... <NL>`` per read (SquigglePull.py:251-253); heads: list of str, e.g. "f.fast5\\trid"."""
    lib = _cabi.lib()
    signals = np.ascontiguousarray(signals, dtype=np.int16)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n = offsets.size - 1
    enc = [h.encode() if isinstance(h, str) else bytes(h) for h in heads]
    hoff = np.zeros(n + 1, dtype=np.int64)
    np.cumsum([len(e) for e in enc], out=hoff[1:])
    hb = b"".join(enc)
    need = -int(lib.sqk_tsv_format(signals.ctypes.data, offsets.ctypes.data, n, hb, hoff.ctypes.data, n_threads, None, 0))
    out = bytearray(need)
    if need:
        view = (C.c_char * need).from_buffer(out)
        got = int(lib.sqk_tsv_format(signals.ctypes.data, offsets.ctypes.data, n, hb, hoff.ctypes.data, n_threads, C.addressof(view), need))
        del view
        assert got == need
    return bytes(out)


def write_reads(fh, heads, signals, offsets, n_threads: int = 0):
    fh.write(format_reads(heads, signals, offsets, n_threads))


def ndtr(z) -> np.ndarray:
    """``scipy.special.ndtr`` (what ``scipy.stats.norm.cdf`` evaluates) bit for bit, from libsqk (``sqk_ndtr``): the command
    lines do not import scipy."""
    z = np.asarray(z, dtype=np.float64)
    flat = np.ascontiguousarray(z).reshape(-1)
    out = np.empty_like(flat)
    _cabi.lib().sqk_ndtr(flat.ctypes.data, flat.size, out.ctypes.data)
    return out.reshape(z.shape)             # (0-d in, 0-d out: index with [()] for the numpy scalar)


def score_hits(hits: np.ndarray, mod_mean, mod_stdev, n_threads: int = 0):
    """Z-score, p-value and hit probability (MotifSeq.py:443-445) of every record of ``hits`` [n_reads, n_models] for the
    per-model ``mod_mean`` / ``mod_stdev`` -> three float64 arrays [n_reads, n_models] (``sqk_score_hits``)."""
    hits = np.ascontiguousarray(hits)
    n, m = hits.shape
    mm = np.ascontiguousarray(mod_mean, dtype=np.float64)
    ms = np.ascontiguousarray(mod_stdev, dtype=np.float64)
    if mm.size != m or ms.size != m:
        raise ValueError("one mod_mean / mod_stdev per model")
    zs, ps, hps = (np.empty((n, m), dtype=np.float64) for _ in range(3))
    _cabi.lib().sqk_score_hits(hits.ctypes.data, n, m, mm.ctypes.data, ms.ctypes.data, n_threads, zs.ctypes.data, ps.ctypes.data,
                               hps.ctypes.data)
    return zs, ps, hps


def format_hit_rows(heads: bytes, hits: np.ndarray, names, consts, zs: np.ndarray, ps: np.ndarray, hps: np.ndarray,
                    n_threads: int = 0) -> bytes:
    """The rows get_region_multi prints (MotifSeq.py:441-449) for a batch, formatted by libsqk (``sqk_tsv_format_rows``):
    heads = ``Batch.heads_bytes(2)``; hits structured [n_reads, n_models]; names / consts: per model, the motif name and the
    text of ``mod_mean <TAB> mod_stdev``; zs / ps / hps float64 [n_reads, n_models].  Floats come out as Python's repr()."""
    lib = _cabi.lib()
    n, m = hits.shape
    hits = np.ascontiguousarray(hits)
    zs, ps, hps = (np.ascontiguousarray(a, dtype=np.float64) for a in (zs, ps, hps))
    nb = b"".join(x.encode() + b"\0" for x in names)
    cb = b"".join(x.encode() + b"\0" for x in consts)
    heads = heads + b"\0"
    need = -int(lib.sqk_tsv_format_rows(heads, n, hits.ctypes.data, m, nb, cb, zs.ctypes.data, ps.ctypes.data, hps.ctypes.data,
                                        n_threads, None, 0))
    if need <= 0:
        return b""
    out = np.empty(need, dtype=np.uint8)                    # (an upper bound; not zero-filled, copied out once)
    got = int(lib.sqk_tsv_format_rows(heads, n, hits.ctypes.data, m, nb, cb, zs.ctypes.data, ps.ctypes.data, hps.ctypes.data,
                                      n_threads, out.ctypes.data, need))
    return out[:got].tobytes()


def format_seg_rows(heads: bytes, segs: np.ndarray, n_segs: np.ndarray, keep: np.ndarray, n_threads: int = 0) -> bytes:
    """The rows segmenter.py prints (``name <TAB> s0,e0,s1,e1,...``, segmenter.py:130-146) for the reads with ``keep``;
    heads = ``Batch.heads_bytes(1)``; segs [n_reads, max_segs, 2] / n_segs [n_reads] as ``Context.segmenter`` returns them."""
    lib = _cabi.lib()
    segs = np.ascontiguousarray(segs, dtype=np.int32)
    n_segs = np.ascontiguousarray(n_segs, dtype=np.int32)
    keep = np.ascontiguousarray(keep, dtype=np.uint8)
    n, cap_segs = segs.shape[0], segs.shape[1]
    heads = heads + b"\0"
    need = -int(lib.sqk_tsv_format_segs(heads, n, segs.ctypes.data, n_segs.ctypes.data, cap_segs, keep.ctypes.data, n_threads, None, 0))
    if need <= 0:
        return b""
    out = bytearray(need)
    view = (C.c_char * need).from_buffer(out)
    got = int(lib.sqk_tsv_format_segs(heads, n, segs.ctypes.data, n_segs.ctypes.data, cap_segs, keep.ctypes.data, n_threads,
                                      C.addressof(view), need))
    del view
    return bytes(out[:got])
