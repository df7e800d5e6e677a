"""The SquigglePull text reader / writer of libsqk (host code, no GPU): round trips, ragged and odd lines, block
boundaries, the flags that send a line to the float path."""
import gzip

import numpy as np
import pytest

from squigglekit_b200 import tsv


def _mk(rng, n, lo=1, hi=5000, vmin=-300, vmax=2500):
    reads = [rng.integers(vmin, vmax, int(rng.integers(lo, hi))).astype(np.int16) for _ in range(n)]
    off = np.zeros(n + 1, np.int64)
    np.cumsum([r.size for r in reads], out=off[1:])
    return reads, np.concatenate(reads), off


@pytest.mark.parametrize("start_col,gz", [(2, False), (6, False), (2, True)])
def test_round_trip(tmp_path, start_col, gz):
    rng = np.random.default_rng(5)
    reads, sig, off = _mk(rng, 300)
    extra = ["8192.0", "6.0", "1467.61", "4000.0"] if start_col == 6 else []
    heads = ["\t".join([f"file{i}.fast5", f"read-{i}"] + extra) for i in range(300)]
    text = tsv.format_reads(heads, sig, off)
    want = "".join(h + "\t" + "\t".join(str(int(v)) for v in r) + "\n" for h, r in zip(heads, reads)).encode()
    assert text == want                                       # SquigglePull.py:251-253, byte for byte
    path = tmp_path / ("sig.tsv.gz" if gz else "sig.tsv")
    (gzip.open if gz else open)(path, "wb").write(text)
    got_reads, got_heads = [], []
    # small blocks and small batches: lines straddle block boundaries, batches end on either limit
    with tsv.Reader(str(path), start_col, max_lines=37, max_samples=40000, block_bytes=30000, pinned=False) as rd:
        for b in rd:
            assert not b.status.any()
            assert b.heads(2) == [b.head(i)[:2] for i in range(b.n)]      # one C call for the batch == line by line
            assert b.heads(1) == [b.head(i)[:1] for i in range(b.n)]
            for i in range(b.n):
                got_reads.append(b.sig(i).copy()); got_heads.append("\t".join(b.head(i)))
    assert got_heads == heads
    assert len(got_reads) == 300 and all(np.array_equal(a, b) for a, b in zip(got_reads, reads))


def test_flags_and_odd_lines(tmp_path):
    lines = [
        "a.fast5\tr0\t1\t2\t3",                 # plain
        "b.fast5\tr1\t1.5\t2\t3",               # float field -> NOT_INT16
        "c.fast5\tr2",                          # no signal columns
        "d.fast5\tr3\t0\t0\t0",                 # all zero
        "e.fast5\tr4\t40000\t1",                # beyond int16
        "f.fast5\tr5\t-5\t+7\t12\t",            # signs, trailing tab
        "g.fast5\tr6\t1e3\t5",                  # exponent
        "h.fast5\tr7\t\t5",                     # empty field
        "i.fast5\tr8\t9\t8\t7\r",               # CRLF
        "j.fast5\tr9\t32767\t-32768",           # limits; last line without newline
    ]
    path = tmp_path / "odd.tsv"
    path.write_bytes("\n".join(lines).encode())
    st, sigs, heads, tails = [], [], [], []
    with tsv.Reader(str(path), 2, pinned=False) as rd:
        for b in rd:                            # (the unterminated last line arrives once the end of the file is known)
            st += b.status.tolist()
            for i in range(b.n):
                sigs.append(b.sig(i).tolist()); heads.append(b.head(i)); tails.append(b.tail_text(i))
    assert st == [0, tsv.NOT_INT16, tsv.NO_SIGNAL, tsv.ALL_ZERO, tsv.NOT_INT16, 0, tsv.NOT_INT16, tsv.NOT_INT16, 0, 0]
    assert sigs[0] == [1, 2, 3] and sigs[5] == [-5, 7, 12] and sigs[8] == [9, 8, 7] and sigs[9] == [32767, -32768]
    assert heads[2] == ["c.fast5", "r2"] and heads[1] == ["b.fast5", "r1"]
    assert tails[1] == "1.5\t2\t3" and tails[8] == "9\t8\t7"
    # a line larger than the whole sample buffer is an error, not a silent truncation
    with tsv.Reader(str(path), 2, max_samples=2, pinned=False) as rd:
        with pytest.raises(ValueError):
            list(rd)


def test_empty_file(tmp_path):
    path = tmp_path / "empty.tsv"
    path.write_bytes(b"")
    with tsv.Reader(str(path), 4, pinned=False) as rd:
        assert sum(1 for _ in rd) == 0
