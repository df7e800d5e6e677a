"""BLOW5 (binary SLOW5) reader -- the other raw-signal container SquiggleKit touches (SquigglePlot.py and
dRNA_segmenter.py open it through pyslow5, which is not installed here).  Host-side I/O only; yields what the
fast5 reader yields, so both CLIs accept ``--slow5 file.blow5`` as an extra input.

Supported: BLOW5 0.1.x / 0.2.x, record compression none or zlib, uncompressed signal.  zstd records and
svb-zd signal compression are reported as unsupported (convert with ``slow5tools view -c zlib -s none``).
"""
from __future__ import annotations

import struct
import zlib

import numpy as np


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
    if rec_comp not in (0, 1):
        raise Slow5Error(f"{path}: record compression {rec_comp} (zstd) is not supported; use zlib or none")
    if sig_comp != 0:
        raise Slow5Error(f"{path}: signal compression {sig_comp} (svb-zd) is not supported; use `-s none`")
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
        q = 0
        rid_len = struct.unpack_from("<H", rec, q)[0]; q += 2
        read_id = rec[q:q + rid_len].split(b"\0")[0].decode("ascii", "replace"); q += rid_len
        read_group = struct.unpack_from("<I", rec, q)[0]; q += 4
        digitisation, offset, rng, rate = struct.unpack_from("<dddd", rec, q); q += 32
        n = struct.unpack_from("<Q", rec, q)[0]; q += 8
        sig = np.frombuffer(rec, dtype="<i2", count=n, offset=q).copy()
        yield {"read_id": read_id, "signal": sig, "digitisation": digitisation, "offset": offset, "range": rng,
               "sampling_rate": rate, "read_group": read_group, "name": read_id}
