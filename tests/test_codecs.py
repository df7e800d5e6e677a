"""CPU tests of the compressed-container decoders (squigglekit_b200/codecs.py) and of the readers that use them: ONT's VBZ
filter in fast5 files, zstd records and svb-zd signals in BLOW5 files.  The encoders are this repo's own (same published
formats) plus the real libzstd through pyarrow: no MinKNOW- or slow5tools-written file with these codecs exists in the
reference tree, which is why the decoders check every redundant length (see the module docstring)."""
import os
import struct
import zlib

import numpy as np
import pytest

from squigglekit_b200 import codecs, fast5, slow5

pa = pytest.importorskip("pyarrow")


def zstd(b: bytes) -> bytes:
    return pa.compress(b, codec="zstd", asbytes=True)


def squiggle(n, seed=0):
    rng = np.random.default_rng(seed)
    x = 500 + np.cumsum(rng.integers(-6, 7, n)) + rng.integers(-3, 4, n)
    x[rng.integers(0, max(n, 1), n // 50)] = rng.integers(-32768, 32767, n // 50)     # spikes: 3- and 4-byte codes
    return np.clip(x, -32768, 32767).astype(np.int16)


@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 5, 127, 4096, 36978])
def test_streamvbyte_zigzag_round_trip(n):
    x = squiggle(n, n)
    z = codecs.zigzag_delta_encode(x)
    s = codecs.svb_encode(z)
    assert len(s) == (n + 3) // 4 + int(((z > 0xFF).astype(int) + (z > 0xFFFF) + (z > 0xFFFFFF) + 1).sum())
    v, used = codecs.svb_decode(s + b"tail", n)
    assert used == len(s) and np.array_equal(v, z)
    assert np.array_equal(codecs.zigzag_delta_decode(v).astype(np.int16), x)


def test_streamvbyte_known_bytes():
    """The format by hand: four values 1, 0x1234, 0x123456, 0x12345678 -> control byte 0b11100100, then 1+2+3+4 data bytes
    little-endian; a fifth value opens a second control byte."""
    vals = np.array([1, 0x1234, 0x123456, 0x12345678, 7], dtype=np.uint32)
    want = bytes([0b11100100, 0b00000000, 0x01, 0x34, 0x12, 0x56, 0x34, 0x12, 0x78, 0x56, 0x34, 0x12, 0x07])
    assert codecs.svb_encode(vals) == want
    got, used = codecs.svb_decode(want, 5)
    assert used == len(want) and got.tolist() == vals.tolist()
    # zig-zag: 0, -1, 1, -2, 2 -> 0, 1, 2, 3, 4 on the differences
    assert codecs.zigzag_delta_encode(np.array([0, -1, 0, -2, 0], dtype=np.int16)).tolist() == [0, 1, 2, 3, 4]
    assert codecs.zigzag_delta_decode(np.array([0, 1, 2, 3, 4], dtype=np.uint32)).tolist() == [0, -1, 0, -2, 0]


def test_vbz_chunk_and_errors():
    x = squiggle(5000, 3)
    stream = codecs.svb_encode(codecs.zigzag_delta_encode(x))
    chunk = struct.pack("<I", 2 * x.size) + zstd(stream)
    assert np.array_equal(np.frombuffer(codecs.vbz_decode(chunk, (0, 2, 1, 1), 2), dtype="<i2"), x)
    # without zstd (level 0), without zig-zag, 4-byte integers, plain bytes (integer size 0)
    assert np.array_equal(np.frombuffer(codecs.vbz_decode(struct.pack("<I", 2 * x.size) + stream, (0, 2, 1, 0), 2), dtype="<i2"), x)
    u = np.arange(0, 70000, 7, dtype=np.uint16)
    c2 = struct.pack("<I", 2 * u.size) + zstd(codecs.svb_encode(u.astype(np.uint32)))
    assert np.array_equal(np.frombuffer(codecs.vbz_decode(c2, (0, 2, 0, 1), 2), dtype="<u2"), u)
    assert codecs.vbz_decode(struct.pack("<I", 5) + zstd(b"hello"), (0, 0, 0, 1), 1) == b"hello"
    for bad, cd, why in ((struct.pack("<I", 2 * x.size + 2) + zstd(stream), (0, 2, 1, 1), "size header too large"),
                         (struct.pack("<I", 2 * x.size) + zstd(stream + b"x"), (0, 2, 1, 1), "trailing bytes"),
                         (struct.pack("<I", 2 * x.size) + stream, (0, 2, 1, 1), "not a zstd frame"),
                         (chunk, (1, 2, 1, 1), "version 1"), (chunk, (0, 4, 1, 1), "wrong integer size"), (b"ab", (0, 2, 1, 1), "short")):
        with pytest.raises(codecs.CodecError):
            codecs.vbz_decode(bad, cd, 2)


def test_zstd_both_decoders_agree(monkeypatch):
    data = bytes(squiggle(20000, 5).tobytes())
    frame = zstd(data)
    assert codecs.zstd_decompress(frame) == data
    monkeypatch.setattr(codecs, "_zstd", False)              # no libzstd: pyarrow's codec, size from the frame header
    assert codecs.zstd_decompress(frame) == data
    for n in (0, 1, 255, 256, 65791, 65792, 1 << 20):
        assert codecs._zstd_content_size(zstd(b"a" * n)) == n


def _chunk_keys(f, addr, ndims):
    """(position of the key in the file buffer, offset, chunk address, size, mask) of every raw-data chunk"""
    p = addr + f.base
    level, n = f.buf[p + 5], f._u16(p + 6)
    key_size = 8 + 8 * ndims
    q = p + 24
    for _ in range(n):
        child = f._u64(q + key_size)
        if level == 0:
            yield q, f._u64(q + 8), child, f._u32(q), f._u32(q + 4)
        else:
            yield from _chunk_keys(f, child, ndims)
        q += key_size + 8


def test_fast5_with_vbz_chunks(golden_dir, tmp_path):
    """The reference's example read with its Signal chunks re-encoded as VBZ in place (chunk bytes + the size in the chunk
    B-tree key; the filter list is swapped on the parsed dataset): the reader returns the same samples."""
    src = os.path.join(golden_dir, "test.fast5")
    want = fast5.read_single_fast5(src)["signal"]
    f = fast5.Fast5File(src)
    reads = f["Raw/Reads"]
    node = reads[reads.keys()[0]]["Signal"]
    ds = node.o.ds
    assert ds.layout[0] == "chunked"
    buf = bytearray(f.buf)
    n_chunks = 0
    for q, off, addr, size, mask in _chunk_keys(f, ds.layout[1], len(ds.layout[2])):
        take = min(ds.layout[2][0], want.size - off)
        chunk = want[off:off + take]
        if take < ds.layout[2][0]:                           # HDF5 stores whole chunks
            chunk = np.concatenate([chunk, np.zeros(ds.layout[2][0] - take, np.int16)])
        enc = struct.pack("<I", 2 * chunk.size) + zstd(codecs.svb_encode(codecs.zigzag_delta_encode(chunk)))
        if len(enc) > size:
            pytest.skip("VBZ chunk larger than the deflate chunk it would replace")
        buf[addr + f.base:addr + f.base + len(enc)] = enc
        struct.pack_into("<II", buf, q, len(enc), 0)
        n_chunks += 1
    assert n_chunks >= 1
    dst = tmp_path / "vbz.fast5"
    dst.write_bytes(bytes(buf))
    g = fast5.Fast5File(str(dst))
    reads = g["Raw/Reads"]
    node = reads[reads.keys()[0]]["Signal"]
    node.o.ds.filters = [(32020, [0, 2, 1, 1])]
    assert np.array_equal(node.read(), want)
    node.o.ds.filters = [(32020, [1, 2, 1, 1])]
    with pytest.raises(fast5.Fast5Error):
        node.read()


def _blow5(records, rec_comp, sig_comp, with_size):
    """A BLOW5 0.2.0 file with the given record / signal compression (writer for this test only)."""
    hdr = b"#slow5_version\t0.2.0\n#num_read_groups\t1\n#char*\tuint32_t\tdouble\tdouble\tdouble\tdouble\tuint64_t\tint16_t*\n#read_id\tread_group\tdigitisation\toffset\trange\tsampling_rate\tlen_raw_signal\traw_signal\n"
    out = bytearray(b"BLOW5\x01" + bytes([0, 2, 0, rec_comp, sig_comp]))
    out += b"\0" * (64 - len(out)) + struct.pack("<I", len(hdr)) + hdr
    for rid, sig in records:
        r = struct.pack("<H", len(rid) + 1) + rid.encode() + b"\0" + struct.pack("<I", 0) + struct.pack("<dddd", 8192.0, 6.0, 1467.61, 4000.0)
        r += struct.pack("<Q", sig.size)
        if sig_comp == 0:
            r += sig.astype("<i2").tobytes()
        else:
            field = struct.pack("<I", sig.size) + codecs.svb_encode(codecs.zigzag_delta_encode(sig))
            r += (struct.pack("<Q", len(field)) if with_size else b"") + field
        r += b"aux-bytes"
        if rec_comp == 1:
            r = zlib.compress(r)
        elif rec_comp == 2:
            r = zstd(r)
        out += struct.pack("<Q", len(r)) + r
    return bytes(out + b"5WOLB")


@pytest.mark.parametrize("rec_comp,sig_comp,with_size", [(0, 0, False), (1, 0, False), (2, 0, False), (1, 1, False), (2, 1, False),
                                                         (0, 1, True), (2, 1, True)])
def test_blow5_compressed_records(tmp_path, rec_comp, sig_comp, with_size):
    recs = [(f"read-{i}", squiggle(n, i)) for i, n in enumerate((1, 4, 777, 20000))]
    p = tmp_path / "c.blow5"
    p.write_bytes(_blow5(recs, rec_comp, sig_comp, with_size))
    got = list(slow5.read_blow5(str(p)))
    assert [g["read_id"] for g in got] == [r[0] for r in recs]
    for g, (_, sig) in zip(got, recs):
        assert g["signal"].dtype == np.int16 and np.array_equal(g["signal"], sig)
        assert (g["digitisation"], g["offset"], g["range"], g["sampling_rate"]) == (8192.0, 6.0, 1467.61, 4000.0)


def test_blow5_refuses_what_it_does_not_know(tmp_path):
    p = tmp_path / "x.blow5"
    p.write_bytes(_blow5([("r", squiggle(10))], 0, 2, False))
    with pytest.raises(slow5.Slow5Error):
        list(slow5.read_blow5(str(p)))
    bad = bytearray(_blow5([("r", squiggle(100))], 0, 1, False))
    i = bad.index(struct.pack("<Q", 100)) + 8
    struct.pack_into("<I", bad, i, 99)                        # the embedded count no longer matches len_raw_signal
    p.write_bytes(bytes(bad))
    with pytest.raises(slow5.Slow5Error):
        list(slow5.read_blow5(str(p)))


def test_corrupt_streams_raise_instead_of_allocating():
    """A count that the stream cannot possibly hold (a corrupt header) is refused before anything is sized by it."""
    with pytest.raises(codecs.CodecError):
        codecs.svb_decode(b"\x00" * 10, 1 << 31)
    with pytest.raises(codecs.CodecError):
        codecs.svb_zd_decode(struct.pack("<I", 1 << 30) + b"\x00" * 16, 1 << 30)
    with pytest.raises(codecs.CodecError):
        codecs.vbz_decode(struct.pack("<I", 0xFFFFFFF0) + zstd(b"\x00" * 8), (0, 2, 1, 1), 2)
    rng = np.random.default_rng(9)
    for _ in range(200):                                      # random garbage: an error or some array, never a crash
        blob = rng.integers(0, 256, int(rng.integers(0, 64)), dtype=np.uint8).tobytes()
        n = int(rng.integers(0, 80))
        try:
            v, used = codecs.svb_decode(blob, n)
            assert v.size == n and used <= len(blob)
        except codecs.CodecError:
            pass
