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


def test_ndtr_is_scipys_bit_for_bit():
    """sqk_ndtr restates scipy.special.ndtr (Cephes ndtr.c: what scipy.stats.norm.cdf -- MotifSeq.py:444 -- evaluates) so
    that the command line needs no scipy import: same bits over the whole range, the branch points and the specials."""
    ndtr_ref = pytest.importorskip("scipy.special").ndtr
    rng = np.random.default_rng(11)
    z = np.concatenate([rng.normal(0, 3, 1_000_000), rng.uniform(-40, 40, 300_000), rng.normal(0, 0.3, 300_000),
                        np.sqrt(2.0) * np.array([1.0, -1.0, 8.0, -8.0, np.nextafter(1.0, 0), np.nextafter(1.0, 2), np.nextafter(8.0, 0)]),
                        np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-300, -1e-300, 5e-324, 37.6, -37.6, -38.6, 26.7, -26.7])])
    got, want = tsv.ndtr(z), ndtr_ref(z)
    assert np.array_equal(got.view(np.int64)[~np.isnan(want)], want.view(np.int64)[~np.isnan(want)])
    assert np.isnan(got[np.isnan(want)]).all()


def test_score_hits_matches_the_reference_expressions():
    """Z, p-value and hit probability as get_region_multi computes them (MotifSeq.py:441-445), per (read, model)."""
    st = pytest.importorskip("scipy.stats")
    from squigglekit_b200 import core
    rng = np.random.default_rng(12)
    hits = np.zeros((5000, 3), dtype=core.HIT_DTYPE)
    hits["dist"] = rng.random((5000, 3)) * 80
    hits["dist"][7, 1] = np.nan                              # a status record
    m, b, std, L = 2.90, -9.6, 0.08468, [10, 20, 7]
    means = [(m * l) + b for l in L]
    stdevs = [mm * std for mm in means]
    zs, ps, hps = tsv.score_hits(hits, means, stdevs)
    for c in range(3):
        Z = (hits["dist"][:, c] - means[c]) / stdevs[c]
        p = st.norm.cdf(Z)
        hp = (1 - p) * 100
        for got, want in ((zs[:, c], Z), (ps[:, c], p), (hps[:, c], hp)):
            ok = ~np.isnan(want)
            assert np.array_equal(got[ok].view(np.int64), want[ok].view(np.int64)) and np.isnan(got[~ok]).all()


def _py_line(line: str, start_col: int):
    """What sqk_tsv_parse must say about one line (without its newline): (status, samples or None)."""
    cols = line.split("\t")
    if len(cols) < start_col + 1:
        return tsv.NO_SIGNAL, None
    fields = cols[start_col:]
    if fields and fields[-1] == "":
        fields = fields[:-1]                     # a trailing tab does not open a field
    if not fields:
        return tsv.NO_SIGNAL, None
    if fields[-1].endswith("\r"):
        fields = fields[:-1] + [fields[-1][:-1]]  # CRLF
    vals = []
    for f in fields:
        body = f[1:] if f[:1] in ("+", "-") else f
        if not body.isascii() or not body.isdigit() or len(body) > 6 or not -32768 <= int(f) <= 32767:
            return tsv.NOT_INT16, None
        vals.append(int(f))
    return (tsv.ALL_ZERO if not any(vals) else 0), vals


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_parser_fuzz_against_python(tmp_path, seed):
    """Random lines -- plain integers of every width, signs, leading zeros, values around the int16 limits, empty fields,
    junk, floats, CRLF, trailing tabs, an unterminated last line -- against a Python statement of the rules."""
    rng = np.random.default_rng(seed)
    pool = ["0", "7", "42", "511", "1200", "32767", "32768", "-32768", "-32769", "+5", "-0", "007", "000123", "0000123", "99999",
            "123456", "1234567", "", "1.5", "1e3", "abc", "12a", "-", "+", " 5", "5 ", "65536", "-1"]
    lines = []
    for i in range(1500):
        kind = rng.random()
        n = int(rng.integers(0, 40))
        if kind < 0.6:
            f = [str(int(v)) for v in rng.integers(0, 1300, n)]
        elif kind < 0.75:
            f = [str(int(v)) for v in rng.integers(-40000, 40000, n)]
        elif kind < 0.8:
            f = ["0"] * n
        else:
            f = [pool[int(j)] for j in rng.integers(0, len(pool), n)]
        head = [f"f{i}.fast5", f"r{i}", "a", "b"][: int(rng.integers(1, 5))] if rng.random() < 0.1 else [f"f{i}.fast5", f"r{i}", "a", "b"]
        line = "\t".join(head + f)
        if rng.random() < 0.1:
            line += "\t"
        if rng.random() < 0.1:
            line += "\r"
        lines.append(line)
    text = "\n".join(lines) + ("\n" if seed != 3 else "")       # seed 3: the last line has no newline
    path = tmp_path / "fuzz.tsv"
    path.write_bytes(text.encode())
    got = []
    with tsv.Reader(str(path), 4, max_lines=257, pinned=False, n_threads=3) as rd:
        for b in rd:
            for i in range(b.n):
                got.append((int(b.status[i]), b.sig(i).tolist()))
    assert len(got) == len(lines)
    for i, line in enumerate(lines):
        st, vals = _py_line(line, 4)
        assert got[i][0] == st, (i, line, got[i][0], st)
        if vals is not None and not st & tsv.NOT_INT16:
            assert got[i][1] == vals, (i, line)


def _collect(path, **kw):
    rows = []
    with tsv.Reader(str(path), 2, pinned=False, **kw) as rd:
        for b in rd:
            for i in range(b.n):
                rows.append((b.head(i), int(b.status[i]), b.sig(i).tolist()))
    return rows


def test_prefetch_reader_equals_plain_reader(tmp_path):
    """The two-slot reader with its helper thread yields the same lines, in order, as the single-buffer one -- whatever the
    batch limits; leaving the loop early and errors raised by the parser do not hang it."""
    rng = np.random.default_rng(4)
    lines = []
    for i in range(3000):
        n = int(rng.integers(1, 60))
        lines.append("\t".join([f"f{i}.fast5", f"r{i}"] + [str(int(v)) for v in rng.integers(0, 1300, n)]))
    path = tmp_path / "p.tsv"
    path.write_bytes(("\n".join(lines) + "\n").encode())
    want = _collect(path, prefetch=False, max_lines=100000, max_samples=1 << 20)
    assert len(want) == 3000 and want[17][0] == ["f17.fast5", "r17"]
    for kw in ({"max_lines": 7}, {"max_lines": 256, "max_samples": 4096}, {"max_lines": 100000, "max_samples": 1 << 20},
               {"max_lines": 1, "max_samples": 128}):
        assert _collect(path, prefetch=True, **kw) == want, kw
        assert _collect(path, prefetch=False, **kw) == want, kw
    # early exit: the helper thread is told to stop and joined
    import threading
    before = threading.active_count()
    with tsv.Reader(str(path), 2, max_lines=5, pinned=False) as rd:
        for k, b in enumerate(rd):
            if k == 2:
                break
    assert threading.active_count() == before
    # a line that does not fit a slot: the parser's error surfaces in the consuming thread
    with tsv.Reader(str(path), 2, max_lines=64, max_samples=64, pinned=False) as rd:
        with pytest.raises(ValueError):
            list(rd)
    assert threading.active_count() == before


def test_prefetch_reader_consumer_error_and_close(tmp_path):
    """An exception in the consuming loop, or close() while a batch is pending, stops the helper thread before the sample
    buffer is released."""
    import threading
    lines = ["\t".join([f"f{i}.fast5", f"r{i}"] + ["5"] * 20) for i in range(500)]
    path = tmp_path / "e.tsv"
    path.write_bytes(("\n".join(lines) + "\n").encode())
    before = threading.active_count()
    with pytest.raises(RuntimeError):
        with tsv.Reader(str(path), 2, max_lines=10, pinned=False) as rd:
            for k, b in enumerate(rd):
                if k == 1:
                    raise RuntimeError("consumer failed")
    assert threading.active_count() == before
    rd = tsv.Reader(str(path), 2, max_lines=10, pinned=False)
    it = iter(rd)
    next(it)
    rd.close()                                               # generator still alive: close() must not leave the thread behind
    assert threading.active_count() == before
    del it
