"""BLOW5 (binary SLOW5) reader -- the other raw-signal container SquiggleKit touches (SquigglePlot.py and
dRNA_segmenter.py open it through pyslow5, which is not installed here).  Host-side I/O only; yields what the
fast5 reader yields, so both CLIs accept ``--slow5 file.blow5`` as an extra input.

Supported: BLOW5 0.1.x / 0.2.x and later, record compression none / zlib / zstd, signal stored plain or svb-zd
(StreamVByte of zig-zag deltas, slow5tools' default).  The decoders are in ``codecs.py``; the plain and zlib layouts are
pinned on the reference's ``example/slow5/0.blow5``, the zstd / svb-zd layouts on records written by this module's own
encoder only (no slow5tools-written file with them was available), so every redundant length is checked: the svb-zd field
is accepted as ``u32 count | stream`` directly behind ``len_raw_signal`` or behind one more u64 byte count, whichever has
``count == len_raw_signal``.  ex-zd signal compression is refused.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

from . import codecs


class Slow5Error(Exception):
    pass


def read_blow5(path):
    """Yield dict(read_id, signal int16, digitisation, offset, range, sampling_rate, read_group) per record."""
    with open(path, "rb") as fh:
        buf = fh.read()
    if buf[:6] != b"BLOW5\x01":
        raise Slow5Error(f"{path}: not a BLOW5 file")
    major, minor, _patch = buf[6], buf[7], buf[8]
    rec_comp = buf[9]
    pos = 10
    sig_comp = 0
    if (major, minor) >= (0, 2):
        sig_comp = buf[pos]
        pos += 1
    if rec_comp not in (0, 1, 2):
        raise Slow5Error(f"{path}: record compression method {rec_comp} is not supported (none, zlib, zstd are)")
    if sig_comp not in (0, 1):
        raise Slow5Error(f"{path}: signal compression method {sig_comp} (ex-zd?) is not supported; use `-s none` or `-s svb-zd`")
    hdr_size = struct.unpack_from("<I", buf, 64)[0]
    pos = 68 + hdr_size
    end = len(buf)
    while pos + 8 <= end:
        if buf[pos:pos + 5] == b"5WOLB":
            break
        rec_size = struct.unpack_from("<Q", buf, pos)[0]
        pos += 8
        rec = buf[pos:pos + rec_size]
        pos += rec_size
        if rec_comp == 1:
            rec = zlib.decompress(rec)
        elif rec_comp == 2:
            try:
                rec = codecs.zstd_decompress(rec)
            except codecs.CodecError as e:
                raise Slow5Error(f"{path}: {e}") from e
        q = 0
        rid_len = struct.unpack_from("<H", rec, q)[0]; q += 2
        read_id = rec[q:q + rid_len].split(b"\0")[0].decode("ascii", "replace"); q += rid_len
        read_group = struct.unpack_from("<I", rec, q)[0]; q += 4
        digitisation, offset, rng, rate = struct.unpack_from("<dddd", rec, q); q += 32
        n = struct.unpack_from("<Q", rec, q)[0]; q += 8
        if sig_comp == 0:
            sig = np.frombuffer(rec, dtype="<i2", count=n, offset=q).copy()
        else:
            sig = None
            for skip in (0, 8):                               # u32 count | stream, directly or behind a u64 byte count
                if len(rec) >= q + skip + 4 and struct.unpack_from("<I", rec, q + skip)[0] == n:
                    try:
                        sig, _ = codecs.svb_zd_decode(rec[q + skip:], n)
                        break
                    except codecs.CodecError:
                        sig = None
            if sig is None:
                raise Slow5Error(f"{path}: read {read_id}: svb-zd signal field not understood")
        yield {"read_id": read_id, "signal": sig, "digitisation": digitisation, "offset": offset, "range": rng,
               "sampling_rate": rate, "read_group": read_group, "name": read_id}
