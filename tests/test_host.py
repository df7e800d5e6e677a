"""CPU tests of the host-side mirror of the reference interface: model files, fast5 extraction,
test_segs, TSV row formatting, CLI flag surface."""
import io
import json
import os

import numpy as np
import pytest

import squigglekit_b200 as sqk
from squigglekit_b200 import cli_motifseq, cli_segmenter, fast5


def test_read_synth_model_matches_reference_output(golden_dir, tmp_path):
    exp = json.load(open(os.path.join(golden_dir, "example_expected.json")))
    ex = np.load(os.path.join(golden_dir, "example_read.npz"), allow_pickle=True)
    p = tmp_path / "m.model"
    p.write_text(exp["model_text"])
    for reader in (sqk.read_synth_model, sqk.read_model):
        model, order, L = reader(str(p))
        assert order == ["3_prime_end"] and L == [20]
        assert np.array_equal(model["3_prime_end"], ex["model"])


def test_read_synth_model_multiple_blocks(tmp_path):
    p = tmp_path / "two.model"
    p.write_text("#a\npos\tbase\tcurrent\tsd\tdwell\n0\tA\t0.5\t0.1\t2.4\n1\tC\t-1.0\t0.1\t1.5\n#b\n0\tG\t2.0\t0.2\t3.0\n")
    model, order, L = sqk.read_synth_model(str(p))
    assert order == ["a", "b"] and L == [2, 1]
    assert model["a"].tolist() == [0.5, 0.5, -1.0, -1.0] and model["b"].tolist() == [2.0, 2.0, 2.0]


def test_read_bait_model_fills_order_and_lengths(tmp_path):
    p = tmp_path / "bait.tsv"
    p.write_text("adapter\t12\tx\t1.0\t2.0\t3.5\npolyA\t8\tx\t-0.5\t0.25\n")
    model, order, L = sqk.read_model(str(p))
    assert order == ["adapter", "polyA"] and L == [12, 8]
    assert model["adapter"].tolist() == [1.0, 2.0, 3.5]


def test_fast5_reader_on_the_reference_example(golden_dir):
    rec = fast5.read_single_fast5(os.path.join(golden_dir, "test.fast5"))
    ex = np.load(os.path.join(golden_dir, "example_read.npz"), allow_pickle=True)
    assert rec["signal"].dtype == np.int16 and np.array_equal(rec["signal"], ex["raw"])
    assert rec["read_id"] == str(ex["read_id"])
    assert (rec["digitisation"], rec["offset"], rec["sampling_rate"]) == (8192.0, 16.0, 4000.0)
    assert float("{0:.2f}".format(rec["range"])) == 1493.94
    f = fast5.Fast5File(os.path.join(golden_dir, "test.fast5"))
    assert sorted(f.keys()) == ["Analyses", "Raw", "UniqueGlobalKey"]
    assert not fast5.is_multi_fast5(f)


def test_test_segs_semantics():
    cfg = sqk.SegConfig(stall=True, stall_start=300, gap=True, gap_dist=100)
    assert sqk.test_segs([[10, 200], [250, 500]], cfg) == [[10, 200], [250, 500]]
    assert sqk.test_segs([[301, 600]], cfg) is False                      # first segment starts too late
    assert sqk.test_segs([[10, 200], [301, 500]], cfg) is False           # second segment too far
    assert sqk.test_segs([[10, 200]], cfg) == [[10, 200]]                 # single segment: IndexError swallowed upstream
    assert sqk.test_segs(False, cfg) is False


def test_row_formatting_matches_reference_rows(golden_dir):
    """format_row reproduces the TSV rows the reference's get_region_multi printed (golden), given the
    same start/end/dist."""
    exp = json.load(open(os.path.join(golden_dir, "example_expected.json")))
    for scale in ("zscale", "medmad"):
        want = exp["tsv"][scale].rstrip("\n")
        f = want.split("\t")
        row = cli_motifseq.format_row(f[0], f[1], f[2], int(f[3]), int(f[4]), np.float64(float(f[6])), 2.90, -9.6, 0.08468, 20)
        assert row == want
    rows = json.load(open(os.path.join(golden_dir, "motifseq_rows.json")))
    g = np.load(os.path.join(golden_dir, "motifseq_golden.npz"))
    for key, L in (("motif80_zscale", 10), ("example163_medmad", 20)):
        lines = rows[key].rstrip("\n").split("\n")
        for line in lines[:5]:
            f = line.split("\t")
            assert cli_motifseq.format_row(f[0], f[1], f[2], int(f[3]), int(f[4]), np.float64(float(f[6])), 2.90, -9.6, 0.08468, L) == line


def test_cli_flag_surface_matches_reference():
    ms = {a.dest for a in cli_motifseq.build_parser()._actions}
    for flag in ("f5f", "f5_path", "signal", "scale", "fasta_input", "scrappie_model", "model", "sig_extract", "slope",
                 "intercept", "std_const", "view", "save", "img", "scale_hi", "scale_low", "version", "verbose"):
        assert flag in ms, flag
    sg = {a.dest for a in cli_segmenter.build_parser()._actions}
    for flag in ("ind", "f5_path", "signal", "single", "Num", "error", "corrector", "window", "seg_dist", "std_scale", "view",
                 "gap", "gap_dist", "stall", "test", "stall_len", "stall_start", "lim_hi", "lim_low", "raw_signal"):
        assert flag in sg, flag
    d = cli_segmenter.build_parser().parse_args(["-s", "x"])
    assert (d.error, d.corrector, d.window, d.seg_dist, d.std_scale, d.stall_len, d.stall_start, d.gap_dist, d.lim_hi, d.lim_low) == \
        (5, 50, 150, 50, 0.75, 0.25, 300, 3000, 900, 0)
    m = cli_motifseq.build_parser().parse_args(["-s", "x", "-m", "y"])
    assert (m.scale, m.slope, m.intercept, m.std_const, m.scale_hi, m.scale_low) == ("medmad", 2.90, -9.6, 0.08468, 1200, 0)


def test_blow5_reader_on_the_reference_example(golden_dir):
    from squigglekit_b200 import slow5
    recs = list(slow5.read_blow5(os.path.join(golden_dir, "example.blow5")))
    ex = np.load(os.path.join(golden_dir, "example_read.npz"), allow_pickle=True)
    assert len(recs) == 1
    r = recs[0]
    assert r["read_id"] == str(ex["read_id"]) and np.array_equal(r["signal"], ex["raw"])
    assert (r["digitisation"], r["offset"], r["sampling_rate"]) == (8192.0, 16.0, 4000.0)


def test_vectorised_rows_equal_the_per_read_rows():
    """cli_motifseq.format_rows (a batch at a time) prints exactly what format_row prints read by read -- the row
    get_region_multi prints (MotifSeq.py:441-449), Python float repr and all."""
    import squigglekit_b200 as sqk
    from squigglekit_b200 import cli_motifseq
    rng = np.random.default_rng(8)
    n, names, L = 500, ["m1", "polyA"], [20, 8]
    hits = np.zeros((n, 2), dtype=sqk.HIT_DTYPE)
    hits["start"] = rng.integers(0, 4000, (n, 2)); hits["end"] = hits["start"] + rng.integers(1, 300, (n, 2))
    hits["dist"] = np.abs(rng.normal(40, 30, (n, 2))) * 10.0 ** rng.integers(-6, 3, (n, 2))
    hits["dist"][3, 0] = 1e-300; hits["dist"][4, 1] = 1e22; hits["dist"][5, 0] = 48.4
    hits["start"][7, 1] = hits["end"][7, 1] = -1; hits["start"][9, 0] = hits["end"][9, 0] = -2
    heads = [(f"f{i}.fast5", f"r{i}") for i in range(n)]
    rows, skipped = cli_motifseq.format_rows(heads, names, hits, 2.90, -9.6, 0.08468, L)
    want = []
    for r in range(n):
        for c, name in enumerate(names):
            h = hits[r, c]
            if int(h["start"]) < 0:
                continue
            want.append(cli_motifseq.format_row(heads[r][0], heads[r][1], name, int(h["start"]), int(h["end"]), h["dist"], 2.90, -9.6,
                                                0.08468, L[c]))
    assert rows == want
    # format_row itself against the reference's expressions with the real scipy (MotifSeq.py:441-449)
    st = pytest.importorskip("scipy.stats")
    k = 0
    for r in range(n):
        for c, name in enumerate(names):
            h = hits[r, c]
            if int(h["start"]) < 0:
                continue
            start, end, dist = int(h["start"]), int(h["end"]), h["dist"]
            mod_mean = (2.90 * L[c]) + -9.6
            mod_stdev = mod_mean * 0.08468
            Z = (dist - mod_mean) / mod_stdev
            p_value = st.norm.cdf(Z)
            hit_P = (1 - p_value) * 100
            ref = "\t".join("{}".format(x) for x in [heads[r][0], heads[r][1], name, start, end, end - start, dist, mod_mean, mod_stdev, Z, p_value, hit_P])
            assert want[k] == ref
            k += 1
    assert sorted(skipped) == [(7, -1), (9, -2)]
    # ... and so does the formatter in libsqk (floats as Python's repr writes them)
    hb = ("\n".join(f"{a}\t{b}" for a, b in heads) + "\n").encode()
    text, skipped_c = cli_motifseq.format_rows_bytes(hb, names, hits, 2.90, -9.6, 0.08468, L)
    assert text == ("\n".join(want) + "\n").encode()
    assert sorted(skipped_c) == [(7, -1), (9, -2)]


def test_repr_of_floats_in_libsqk():
    """sqk_tsv_format_rows writes a double exactly as Python's repr() does: shortest round-trip digits, fixed notation
    in [1e-4, 1e16), exponents with at least two digits, subnormals, infinities."""
    import squigglekit_b200 as sqk
    from squigglekit_b200 import tsv
    rng = np.random.default_rng(12)
    xs = np.concatenate([rng.normal(0, 1, 60000) * 10.0 ** rng.integers(-30, 30, 60000), rng.integers(-10**6, 10**6, 5000).astype(float),
                         rng.integers(-10**6, 10**6, 5000) / 1000.0, rng.random(20000), np.ldexp(rng.random(3000), rng.integers(-1074, -1000, 3000)),
                         np.array([0.0, -0.0, 0.1, 0.2, 0.3, 1 / 3, 1e15, 1e16, 1e17, 123456.789, 5e-5, 1e-5, 0.0001, 0.001, 5e-324, 1.7976931348623157e308,
                                   2.2250738585072014e-308, 9007199254740993.0, 0.30000000000000004, np.inf, -np.inf, 1e22, 1e23, 1e100, 1e-100])])
    hits = np.zeros((xs.size, 1), dtype=sqk.HIT_DTYPE)
    hits["dist"][:, 0] = xs; hits["end"] = 1
    out = tsv.format_hit_rows(b"h\n" * xs.size, hits, ["n"], ["c"], xs[:, None], xs[:, None], xs[:, None]).split(b"\n")[:-1]
    assert len(out) == xs.size
    for x, line in zip(xs.tolist(), out):
        f = line.decode().split("\t")
        assert f[5] == repr(x) and f[7] == repr(x) and f[9] == repr(x), (x, f)


def test_segmenter_rows_batched_equal_per_read(capsys):
    """cli_segmenter.emit_batch (arrays + libsqk's row writer) prints what emit prints read by read, for every combination of
    the -u / -k / -g switches."""
    import io
    import types
    import squigglekit_b200 as sqk
    from squigglekit_b200 import cli_segmenter
    rng = np.random.default_rng(21)
    n, cap = 400, 6
    nsegs = rng.integers(0, cap + 1, n).astype(np.int32)
    segs = np.zeros((n, cap, 2), np.int32)
    for r in range(n):
        pos = np.sort(rng.integers(0, 9000, 2 * cap))
        segs[r] = pos.reshape(cap, 2)
        segs[r, nsegs[r]:] = 0
    names = [f"read_{i}.fast5" for i in range(n)]
    hb = ("\n".join(names) + "\n").encode()
    for test in (False, True):
        for stall in (False, True):
            for gap in (False, True):
                args = types.SimpleNamespace(test=test, stall=stall, gap=gap, stall_start=300, gap_dist=3000)
                cfg = sqk.SegConfig(stall=stall, gap=gap, stall_start=300, gap_dist=3000, max_segs=cap)
                a, b = io.StringIO(), io.StringIO()
                cli_segmenter.emit(args, cfg, names, segs, nsegs, a)
                cli_segmenter.emit_batch(args, cfg, hb, segs, nsegs, b)
                assert a.getvalue() == b.getvalue() and a.getvalue().count("\n") > 50
    capsys.readouterr()


def test_clock_sampler_without_nvidia_smi(monkeypatch):
    """bench.py's clock sampler: no nvidia-smi (this container), a disabled sampler (ranks other than 0) and the wait for the
    first sample all end quietly with an empty summary; rows that did arrive are digested."""
    import time
    import bench
    monkeypatch.setenv("PATH", "/nonexistent")
    for enabled in (True, False):
        with bench.ClockSampler(0, enabled=enabled) as c:
            t0 = time.perf_counter()
            c.wait_first(timeout=5.0)
            assert time.perf_counter() - t0 < 1.0
            c.mark(); c.unmark()
            assert c.summary() == {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    c = bench.ClockSampler(0)
    now = time.perf_counter()
    c.rows = [(now - 1.0, ["1965", "1965", "200.0", "0x0", "Not Active", "Not Active", "Not Active", "Not Active"]),
              (now + 0.01, ["1950", "1965", "650.5", "0x4", "Not Active", "Not Active", "Not Active", "Active"])]
    c.mark(); c.t1 = now + 1.0
    s = c.summary()
    assert s["sm_mhz"] == 1950.0 and s["sm_max_mhz"] == 1965.0 and s["reasons"] == ["sw_power_cap"] and s["samples"] == 1
    assert s["samples_incl_warmup"] == 2 and s["sm_mhz_min_incl_warmup"] == 1950.0
