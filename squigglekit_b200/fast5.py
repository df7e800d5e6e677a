"""Raw-signal extraction from fast5 files without h5py (the feeder in front of both hot paths:
process_fast5, MotifSeq.py:327-351 / segmenter.py:321-355; read_multi_fast5, segmenter.py:376-397).

h5py is not installed in the target image, so this is a small pure-Python reader for the subset of
HDF5 that MinKNOW / ont_fast5_api files use: superblock v0-v1, version-1 object headers with
continuation blocks, old-style groups (symbol-table B-trees + local heaps) and compact link
messages, contiguous / compact / chunked datasets with the deflate, shuffle, fletcher32 and VBZ (ONT's
filter 32020: zstd + StreamVByte + zig-zag delta, decoded in ``codecs.py`` -- the reference needs the
`vbz` HDF5 plugin for those files, README.md:100-104) filters, fixed- and variable-length string
attributes (global heap) and numeric scalar attributes.  If h5py happens to be importable it is NOT
used: one code path, tested here.

Host-side I/O only -- the samples it returns go to the GPU as int16.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

from . import codecs

UNDEF = 0xFFFFFFFFFFFFFFFF


class Fast5Error(Exception):
    pass


class _Dataset:
    def __init__(self):
        self.shape = None
        self.dtype = None
        self.layout = None      # ("contiguous", addr, size) | ("compact", bytes) | ("chunked", btree_addr, chunk_dims)
        self.filters = []


class _Object:
    def __init__(self, f, addr):
        self.f, self.addr = f, addr
        self.attrs = {}
        self.links = {}         # name -> object header address
        self.symtab = None      # (btree, heap)
        self.ds = _Dataset()
        self._parsed = False


class Fast5File:
    """Minimal read-only HDF5 view: ``f["Raw/Reads"]`` -> group, ``.keys()``, ``.attrs``, ``.read()``."""

    def __init__(self, path):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        b = self.buf
        if b[:8] != b"\x89HDF\r\n\x1a\n":
            raise Fast5Error(f"{path}: not an HDF5 file")
        ver = b[8]
        if ver > 1:
            raise Fast5Error(f"{path}: HDF5 superblock version {ver} not supported by the built-in reader")
        self.O, self.L = b[13], b[14]
        if self.O != 8 or self.L != 8:
            raise Fast5Error("only 8-byte offsets/lengths supported")
        pos = 24 + (4 if ver == 1 else 0)
        self.base = self._u64(pos)
        pos += 32                           # base, free-space, eof, driver-info addresses
        # root symbol table entry
        self.root = self._obj(self._u64(pos + 8))
        cache_type = self._u32(pos + 16)
        if cache_type == 1:
            self.root.symtab = (self._u64(pos + 24), self._u64(pos + 32))
        self._cache = {}

    # ---- primitives -----------------------------------------------------------------------------
    def _u16(self, p): return struct.unpack_from("<H", self.buf, p)[0]
    def _u32(self, p): return struct.unpack_from("<I", self.buf, p)[0]
    def _u64(self, p): return struct.unpack_from("<Q", self.buf, p)[0]

    def _obj(self, addr):
        return _Object(self, addr)

    # ---- object headers ---------------------------------------------------------------------------
    def _parse(self, o: _Object):
        if o._parsed:
            return o
        b = self.buf
        p = o.addr + self.base
        if b[p:p + 4] == b"OHDR":
            raise Fast5Error("version-2 object headers (libver='latest' files) are not supported by the built-in reader")
        if b[p] != 1:
            raise Fast5Error(f"object header version {b[p]} not supported")
        nmsg = self._u16(p + 2)
        size = self._u32(p + 8)
        blocks = [(p + 16, size)]
        seen = 0
        dtype_raw = None
        while blocks and seen < nmsg:
            q, left = blocks.pop(0)
            end = q + left
            while q + 8 <= end and seen < nmsg:
                mtype, msize, mflags = self._u16(q), self._u16(q + 2), b[q + 4]
                body = q + 8
                seen += 1
                if mtype == 0x10:
                    blocks.append((self._u64(body) + self.base, self._u64(body + 8)))
                elif mtype == 0x11:
                    o.symtab = (self._u64(body), self._u64(body + 8))
                elif mtype == 0x01:
                    o.ds.shape = self._dataspace(body)
                elif mtype == 0x03:
                    dtype_raw = body
                    o.ds.dtype = self._datatype(body)[0]
                elif mtype == 0x08:
                    o.ds.layout = self._layout(body)
                elif mtype == 0x0B:
                    o.ds.filters = self._filters(body)
                elif mtype == 0x0C:
                    name, val = self._attribute(body)
                    o.attrs[name] = val
                elif mtype == 0x06:
                    name, addr = self._link(body)
                    if addr is not None:
                        o.links[name] = addr
                elif mtype == 0x02 and msize >= 18:
                    # link info: dense (fractal heap) storage is beyond this reader
                    flags = b[body + 1]
                    off = body + 2 + (8 if flags & 1 else 0)
                    if self._u64(off) != UNDEF:
                        raise Fast5Error("dense link storage (fractal heap) is not supported by the built-in reader")
                q = body + msize
        o._parsed = True
        return o

    def _dataspace(self, p):
        b = self.buf
        ver, rank, flags = b[p], b[p + 1], b[p + 2]
        q = p + (8 if ver == 1 else 4)
        return tuple(self._u64(q + 8 * i) for i in range(rank))

    def _datatype(self, p):
        """-> (descriptor, total message size is not needed).  descriptor: numpy dtype, ('str', n), ('vlen_str',)"""
        b = self.buf
        cls, ver = b[p] & 0x0F, b[p] >> 4
        bits0 = b[p + 1]
        size = self._u32(p + 4)
        if cls == 0:
            signed = bool(bits0 & 0x08)
            big = bool(bits0 & 0x01)
            return (np.dtype(f"{'>' if big else '<'}{'i' if signed else 'u'}{size}"),)
        if cls == 1:
            big = bool(bits0 & 0x01)
            return (np.dtype(f"{'>' if big else '<'}f{size}"),)
        if cls == 3:
            return (("str", size),)
        if cls == 9:
            vtype = bits0 & 0x0F
            if vtype == 1:
                return (("vlen_str",),)
            base = self._datatype(p + 8)[0]
            return (("vlen", base),)
        if cls == 8:    # enum: treat as its base integer type
            return self._datatype(p + 8)
        return (("opaque", size),)

    def _layout(self, p):
        b = self.buf
        ver = b[p]
        if ver == 3:
            cls = b[p + 1]
            if cls == 0:
                n = self._u16(p + 2)
                return ("compact", bytes(b[p + 4:p + 4 + n]))
            if cls == 1:
                return ("contiguous", self._u64(p + 2), self._u64(p + 10))
            if cls == 2:
                nd = b[p + 2]
                bt = self._u64(p + 3)
                dims = tuple(self._u32(p + 11 + 4 * i) for i in range(nd))
                return ("chunked", bt, dims)
        elif ver in (1, 2):
            nd, cls = b[p + 1], b[p + 2]
            q = p + 8
            addr = None
            if cls != 0:
                addr = self._u64(q)
                q += 8
            dims = tuple(self._u32(q + 4 * i) for i in range(nd))
            q += 4 * nd
            if cls == 1:
                return ("contiguous", addr, None)
            if cls == 2:
                esize = self._u32(q)
                return ("chunked", addr, dims + (esize,))
            if cls == 0:
                n = self._u32(q)
                return ("compact", bytes(b[q + 4:q + 4 + n]))
        raise Fast5Error(f"data layout version {ver} not supported")

    def _filters(self, p):
        b = self.buf
        ver, n = b[p], b[p + 1]
        q = p + (8 if ver == 1 else 2)
        out = []
        for _ in range(n):
            fid = self._u16(q)
            if ver == 1 or fid >= 256:
                nlen = self._u16(q + 2)
                q += 2
            else:
                nlen = 0
            ncd = self._u16(q + 4)
            q += 6
            if ver == 1:
                nlen = (nlen + 7) & ~7
            q += nlen
            cd = [self._u32(q + 4 * i) for i in range(ncd)]
            q += 4 * ncd
            if ver == 1 and ncd % 2:
                q += 4
            out.append((fid, cd))
        return out

    def _link(self, p):
        b = self.buf
        flags = b[p + 1]
        q = p + 2
        ltype = 0
        if flags & 0x08:
            ltype = b[q]; q += 1
        if flags & 0x04:
            q += 8
        if flags & 0x10:
            q += 1
        lsz = 1 << (flags & 3)
        nlen = int.from_bytes(b[q:q + lsz], "little")
        q += lsz
        name = bytes(b[q:q + nlen]).decode("utf-8", "replace")
        q += nlen
        return name, (self._u64(q) if ltype == 0 else None)

    def _attribute(self, p):
        b = self.buf
        ver = b[p]
        nsz, tsz, ssz = self._u16(p + 2), self._u16(p + 4), self._u16(p + 6)
        q = p + 8 + (1 if ver == 3 else 0)
        pad = (lambda n: (n + 7) & ~7) if ver == 1 else (lambda n: n)
        name = bytes(b[q:q + nsz]).split(b"\0")[0].decode("utf-8", "replace")
        q += pad(nsz)
        dt = self._datatype(q)[0]
        q += pad(tsz)
        shape = self._dataspace(q) if ssz >= 4 else ()
        q += pad(ssz)
        count = int(np.prod(shape)) if shape else 1
        return name, self._decode(dt, q, count, shape)

    def _decode(self, dt, q, count, shape):
        b = self.buf
        if isinstance(dt, np.dtype):
            arr = np.frombuffer(b, dtype=dt, count=count, offset=q)
            return arr[0] if not shape else arr.reshape(shape).copy()
        if dt[0] == "str":
            vals = [bytes(b[q + i * dt[1]:q + (i + 1) * dt[1]]).split(b"\0")[0] for i in range(count)]
            return vals[0] if not shape else vals
        if dt[0] == "vlen_str":
            vals = []
            for i in range(count):
                ln = self._u32(q + 16 * i)
                vals.append(self._global_heap(self._u64(q + 16 * i + 4), self._u32(q + 16 * i + 12))[:ln].decode("utf-8", "replace"))
            return vals[0] if not shape else vals
        return None

    def _global_heap(self, addr, index):
        b = self.buf
        p = addr + self.base
        if b[p:p + 4] != b"GCOL":
            raise Fast5Error("bad global heap collection")
        size = self._u64(p + 8)
        q, end = p + 16, p + size
        while q + 16 <= end:
            idx, osz = self._u16(q), self._u64(q + 8)
            if idx == 0:
                break
            if idx == index:
                return bytes(b[q + 16:q + 16 + osz])
            q += 16 + ((osz + 7) & ~7)
        raise Fast5Error("global heap object not found")

    # ---- groups -------------------------------------------------------------------------------------
    def _children(self, o: _Object):
        self._parse(o)
        out = dict(o.links)
        if o.symtab:
            bt, heap = o.symtab
            hp = heap + self.base
            if self.buf[hp:hp + 4] != b"HEAP":
                raise Fast5Error("bad local heap")
            data = self._u64(hp + 24) + self.base
            self._walk_group_btree(bt, data, out)
        return out

    def _walk_group_btree(self, addr, heap_data, out):
        b = self.buf
        p = addr + self.base
        if b[p:p + 4] == b"SNOD":
            n = self._u16(p + 6)
            q = p + 8
            for _ in range(n):
                name_off, ohdr = self._u64(q), self._u64(q + 8)
                s = heap_data + name_off
                e = b.index(b"\0", s)
                out[bytes(b[s:e]).decode("utf-8", "replace")] = ohdr
                q += 40
            return
        if b[p:p + 4] != b"TREE":
            raise Fast5Error("bad group B-tree node")
        n = self._u16(p + 6)
        q = p + 24 + 8                     # skip key 0
        for _ in range(n):
            self._walk_group_btree(self._u64(q), heap_data, out)
            q += 16                        # child + next key

    # ---- public -------------------------------------------------------------------------------------
    def get(self, path):
        o = self.root
        for part in [x for x in path.split("/") if x]:
            kids = self._children(o)
            if part not in kids:
                raise KeyError(path)
            o = self._obj(kids[part])
        return Node(self, self._parse(o))

    __getitem__ = get

    def keys(self):
        return list(self._children(self.root).keys())


class Node:
    def __init__(self, f: Fast5File, o: _Object):
        self.f, self.o = f, o

    @property
    def attrs(self):
        return self.o.attrs

    def keys(self):
        return list(self.f._children(self.o).keys())

    def __getitem__(self, path):
        o = self.o
        for part in [x for x in path.split("/") if x]:
            kids = self.f._children(o)
            if part not in kids:
                raise KeyError(path)
            o = self.f._obj(kids[part])
        return Node(self.f, self.f._parse(o))

    def read(self) -> np.ndarray:
        ds, f = self.o.ds, self.f
        if ds.layout is None or not isinstance(ds.dtype, np.dtype):
            raise Fast5Error("not a numeric dataset")
        n = int(np.prod(ds.shape)) if ds.shape else 1
        kind = ds.layout[0]
        if kind == "compact":
            return np.frombuffer(ds.layout[1], dtype=ds.dtype, count=n).reshape(ds.shape).copy()
        if kind == "contiguous":
            if ds.layout[1] == UNDEF:
                return np.zeros(ds.shape, dtype=ds.dtype)
            return np.frombuffer(f.buf, dtype=ds.dtype, count=n, offset=ds.layout[1] + f.base).reshape(ds.shape).copy()
        # chunked (1-D only: that is what Signal datasets are)
        if len(ds.shape) != 1:
            raise Fast5Error("only 1-D chunked datasets are supported")
        for fid, _ in ds.filters:
            if fid not in (1, 2, 3, 32020):
                raise Fast5Error(f"HDF5 filter {fid} not supported")
        out = np.zeros(ds.shape[0], dtype=ds.dtype)
        chunk_len = ds.layout[2][0]
        esize = ds.dtype.itemsize
        for off, addr, size, mask in self._chunks(ds.layout[1], len(ds.layout[2])):
            raw = bytes(f.buf[addr + f.base: addr + f.base + size])
            for k, (fid, cd) in reversed(list(enumerate(ds.filters))):
                if mask & (1 << k):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    a = np.frombuffer(raw, dtype=np.uint8)
                    cnt = a.size // esize
                    raw = a[:cnt * esize].reshape(esize, cnt).T.tobytes()
                elif fid == 3:
                    raw = raw[:-4]
                elif fid == 32020:                            # ONT's VBZ (zstd + StreamVByte + zig-zag delta), codecs.py
                    try:
                        raw = codecs.vbz_decode(raw, cd, esize)
                    except codecs.CodecError as e:
                        raise Fast5Error(f"VBZ chunk at {addr}: {e}") from e
            vals = np.frombuffer(raw, dtype=ds.dtype, count=min(chunk_len, len(raw) // esize))
            take = min(vals.size, ds.shape[0] - off)
            out[off:off + take] = vals[:take]
        return out

    def _chunks(self, addr, ndims):
        f = self.f
        b = f.buf
        if addr == UNDEF:
            return
        p = addr + f.base
        if b[p:p + 4] != b"TREE":
            raise Fast5Error("bad chunk B-tree node")
        level, n = b[p + 5], f._u16(p + 6)
        key_size = 8 + 8 * ndims
        q = p + 24
        for _ in range(n):
            size, mask = f._u32(q), f._u32(q + 4)
            off0 = f._u64(q + 8)
            child = f._u64(q + key_size)
            if level == 0:
                yield off0, child, size, mask
            else:
                yield from self._chunks(child, ndims)
            q += key_size + 8


# ---- what the reference's extraction functions return -----------------------------------------------
def _txt(v):
    return v.decode("utf-8", "replace") if isinstance(v, (bytes, bytearray)) else str(v)


def is_multi_fast5(f: Fast5File) -> bool:
    return any(k.startswith("read_") for k in f.keys())


def read_single_fast5(path):
    """process_fast5 (MotifSeq.py:327-351, segmenter.py:321-355): first read under Raw/Reads.
    -> dict(signal int16 array, read_id, digitisation, offset, range, sampling_rate)"""
    f = Fast5File(path)
    reads = f["Raw/Reads"]
    first = reads.keys()[0]
    rd = reads[first]
    sig = rd["Signal"].read()
    rid = rd.attrs.get("read_id", b"")
    out = {"signal": np.ascontiguousarray(sig, dtype=np.int16), "read_id": _txt(rid), "name": first,
           "read_id_is_bytes": isinstance(rid, (bytes, bytearray))}
    try:
        ch = f["UniqueGlobalKey/channel_id"].attrs
        out.update(digitisation=float(ch["digitisation"]), offset=float(ch["offset"]), range=float(ch["range"]),
                   sampling_rate=float(ch["sampling_rate"]))
    except KeyError:
        pass
    return out


def read_multi_fast5(path):
    """read_multi_fast5 (segmenter.py:376-397): every read_* group of a multi-read file, in file order.
    -> dict read_name -> dict as above"""
    f = Fast5File(path)
    out = {}
    for name in f.keys():
        if not name.startswith("read_"):
            continue
        g = f[name]
        rec = {"signal": np.ascontiguousarray(g["Raw/Signal"].read(), dtype=np.int16),
               "read_id": _txt(g["Raw"].attrs.get("read_id", b"")), "name": name}
        ch = g["channel_id"].attrs
        rec.update(digitisation=float(ch["digitisation"]), offset=float(ch["offset"]), range=float(ch["range"]),
                   sampling_rate=float(ch["sampling_rate"]))
        out[name] = rec
    return out
