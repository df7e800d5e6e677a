"""Decoders for the compressed signal containers either side of the hot path (host-side I/O):

* **VBZ** -- ONT's HDF5 filter 32020, what MinKNOW writes into fast5 files since 2020 (the reference needs the `vbz` HDF5
  plugin for these, README.md:100-104).  A chunk is ``u32 original_size | zstd frame``; the frame holds a StreamVByte
  stream (control bytes first, two bits per value: 1-4 data bytes) of the zig-zag-coded first differences of the
  samples.  Filter parameters ``cd_values = (version, integer_size, zig_zag, zstd_level)``; version 0 only (what the
  fast5 writers use; anything else is refused).
* **svb-zd** -- slow5lib's signal compression in BLOW5 records: ``u32 count | StreamVByte stream`` of the same zig-zag
  deltas; and **zstd** record compression.

zstd itself is the system's libzstd (ctypes), or pyarrow's codec when the library cannot be loaded.

Status: the StreamVByte / zig-zag / zstd layers are tested here against encoders written from the same published
format descriptions plus the real libzstd; **no file written by MinKNOW or slow5tools was available to check the
container layouts against**, so every length the formats carry redundantly (original size, value count, bytes consumed)
is verified and a mismatch raises instead of returning samples.
"""
from __future__ import annotations

import ctypes as C
import ctypes.util
import struct

import numpy as np


class CodecError(Exception):
    pass


_zstd = None


def _libzstd():
    global _zstd
    if _zstd is None:
        for name in (ctypes.util.find_library("zstd"), "libzstd.so.1", "libzstd.so"):
            if not name:
                continue
            try:
                lib = C.CDLL(name)
                lib.ZSTD_getFrameContentSize.restype = C.c_ulonglong
                lib.ZSTD_getFrameContentSize.argtypes = [C.c_void_p, C.c_size_t]
                lib.ZSTD_decompress.restype = C.c_size_t
                lib.ZSTD_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
                lib.ZSTD_isError.restype = C.c_uint
                lib.ZSTD_isError.argtypes = [C.c_size_t]
                _zstd = lib
                break
            except OSError:
                continue
        if _zstd is None:
            _zstd = False
    return _zstd


def zstd_decompress(data: bytes, max_size: int = 1 << 31) -> bytes:
    """One zstd frame -> bytes (the frame must carry its content size, as the writers of both containers make it)."""
    data = bytes(data)
    lib = _libzstd()
    if lib:
        n = lib.ZSTD_getFrameContentSize(data, len(data))
        if n >= (1 << 64) - 2:                                  # ZSTD_CONTENTSIZE_UNKNOWN / _ERROR
            raise CodecError("zstd frame without a content size (or not a zstd frame)")
        if n > max_size:
            raise CodecError(f"zstd frame claims {n} bytes")
        out = C.create_string_buffer(max(int(n), 1))
        got = lib.ZSTD_decompress(out, int(n), data, len(data))
        if lib.ZSTD_isError(got) or got != n:
            raise CodecError("zstd decompression failed")
        return out.raw[:int(n)]
    try:
        import pyarrow as pa
    except ImportError as e:
        raise CodecError("no zstd decoder available (libzstd not loadable, pyarrow not installed)") from e
    # pyarrow wants the size: read it from the frame header (magic, descriptor, [window], [dict id], content size)
    n = _zstd_content_size(data)
    if n is None or n > max_size:
        raise CodecError("zstd frame without a usable content size")
    return pa.decompress(data, decompressed_size=n, codec="zstd", asbytes=True)


def _zstd_content_size(data: bytes):
    if len(data) < 6 or data[:4] != b"\x28\xb5\x2f\xfd":
        return None
    fhd = data[4]
    fcs_flag, single, did_flag = fhd >> 6, (fhd >> 5) & 1, fhd & 3
    q = 5 + (0 if single else 1) + (0, 1, 2, 4)[did_flag]
    size = (1 if single else 0, 2, 4, 8)[fcs_flag] if fcs_flag else (1 if single else 0)
    if size == 0 or len(data) < q + size:
        return None
    v = int.from_bytes(data[q:q + size], "little")
    return v + 256 if size == 2 else v


def svb_decode(data, count: int):
    """StreamVByte (Lemire's format: ceil(count / 4) control bytes, then the data bytes; code c = c + 1 little-endian bytes)
    -> (uint32 [count], bytes consumed)."""
    buf = np.frombuffer(bytes(data), dtype=np.uint8)
    n_ctrl = (count + 3) // 4
    if count < 0 or buf.size < n_ctrl + count:                # (every value has a control code and at least one data byte)
        raise CodecError("StreamVByte stream shorter than its value count allows")
    ctrl = buf[:n_ctrl]
    codes = ((ctrl[:, None] >> np.array([0, 2, 4, 6], dtype=np.uint8)) & 3).reshape(-1)[:count].astype(np.int64)
    lens = codes + 1
    ends = np.cumsum(lens)
    total = int(ends[-1]) if count else 0
    if buf.size < n_ctrl + total:
        raise CodecError("StreamVByte stream shorter than its control bytes say")
    payload = buf[n_ctrl:n_ctrl + total]
    starts = ends - lens
    out = np.zeros(count, dtype=np.uint32)
    for k in range(4):                                          # byte k of every value that has one
        sel = lens > k
        out[sel] |= payload[starts[sel] + k].astype(np.uint32) << np.uint32(8 * k)
    return out, n_ctrl + total


def zigzag_delta_decode(z: np.ndarray) -> np.ndarray:
    """uint32 zig-zag codes of first differences (first value against 0) -> int32 values (wrapping, as the C code does)."""
    z = z.astype(np.uint32)
    d = ((z >> np.uint32(1)) ^ (np.uint32(0) - (z & np.uint32(1)))).astype(np.uint32)
    return np.cumsum(d, dtype=np.uint32).view(np.int32)


def vbz_decode(chunk, cd_values, itemsize: int) -> bytes:
    """One HDF5 chunk behind filter 32020 -> the bytes of the chunk's elements."""
    chunk = bytes(chunk)
    cd = list(cd_values) + [0] * 4
    version, int_size, zig_zag, zstd_level = cd[0], cd[1], cd[2], cd[3]
    if version != 0:
        raise CodecError(f"VBZ stream version {version} is not supported (version 0 only)")
    if len(chunk) < 4:
        raise CodecError("VBZ chunk shorter than its size header")
    original = struct.unpack_from("<I", chunk, 0)[0]
    body = chunk[4:]
    if zstd_level != 0:
        body = zstd_decompress(body)
    if int_size == 0:
        if len(body) != original:
            raise CodecError("VBZ chunk: size header and payload disagree")
        return body
    if int_size not in (1, 2, 4) or int_size != itemsize or original % int_size:
        raise CodecError(f"VBZ chunk: integer size {int_size} does not fit the dataset (element size {itemsize}, {original} bytes)")
    count = original // int_size
    vals, used = svb_decode(body, count)
    if used != len(body):
        raise CodecError("VBZ chunk: StreamVByte stream and payload length disagree")
    if zig_zag:
        vals = zigzag_delta_decode(vals)
    return vals.astype({1: "<i1", 2: "<i2", 4: "<i4"}[int_size] if zig_zag else {1: "<u1", 2: "<u2", 4: "<u4"}[int_size]).tobytes()


def svb_zd_decode(data, n_samples: int):
    """slow5lib's svb-zd signal field: ``u32 count | StreamVByte`` of zig-zag deltas -> (int16 [n_samples], bytes consumed)."""
    data = bytes(data)
    if len(data) < 4:
        raise CodecError("svb-zd field shorter than its count")
    count = struct.unpack_from("<I", data, 0)[0]
    if count != n_samples:
        raise CodecError(f"svb-zd field holds {count} values, the record says {n_samples}")
    vals, used = svb_decode(data[4:], count)
    out = zigzag_delta_decode(vals)
    if count and (out.min() < -32768 or out.max() > 32767):
        raise CodecError("svb-zd field decodes to values outside int16")
    return out.astype(np.int16), 4 + used


# ---- encoders: for the tests (and for writing fixtures); same formats ---------------------------------------------------

def svb_encode(vals: np.ndarray) -> bytes:
    v = np.ascontiguousarray(vals, dtype=np.uint32)
    codes = (v > 0xFF).astype(np.uint8) + (v > 0xFFFF) + (v > 0xFFFFFF)
    pad = (-v.size) % 4
    c4 = np.concatenate([codes, np.zeros(pad, np.uint8)]).reshape(-1, 4)
    ctrl = (c4[:, 0] | (c4[:, 1] << 2) | (c4[:, 2] << 4) | (c4[:, 3] << 6)).astype(np.uint8)
    b = v.view(np.uint8).reshape(-1, 4)
    keep = np.arange(4)[None, :] <= codes[:, None]
    return ctrl.tobytes() + b[keep].tobytes()


def zigzag_delta_encode(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x).astype(np.int32)
    d = np.diff(x, prepend=np.int32(0)).astype(np.int32)
    return ((d << 1) ^ (d >> 31)).view(np.uint32)
