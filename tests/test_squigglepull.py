"""CPU tests of the SquigglePull.py drop-in (fast5 -> signal TSV, host-side only): its rows equal the rows the reference's own
extract_f5_all + print_data produced (tests/golden/make_squigglepull_golden.py ran them unmodified over an h5py stand-in),
in raw and pA mode, with and without -i; and what it writes is what the -s readers parse back."""
import io
import json
import os
import shutil
import tarfile
import zlib

import numpy as np
import pytest

from squigglekit_b200 import cli_squigglepull, fast5, tsv

REF_TAR = "/root/reference/example/example_fast5s.tar"


def _check(line, want):
    f = line.split("\t")
    assert len(f) == want["n_fields"]
    assert f[:12] == want["head"] and f[-3:] == want["tail"]
    assert zlib.crc32(line.encode()) == want["crc32"]


@pytest.mark.parametrize("flags", [["-r"], ["-r", "-i"], [], ["-i"]])
def test_rows_equal_the_reference_rows(golden_dir, tmp_path, flags):
    gold = json.load(open(os.path.join(golden_dir, "squigglepull_golden.json")))
    case = next(c for c in gold["cases"] if c["raw_signal"] == ("-r" in flags) and c["extra_info"] == ("-i" in flags))
    d = tmp_path / "f5"
    d.mkdir()
    shutil.copy(os.path.join(golden_dir, "test.fast5"), d / "test.fast5")
    out = io.StringIO()
    cli_squigglepull.main(["-p", str(d)] + flags, out=out)
    lines = out.getvalue().split("\n")
    assert lines[-1] == "" and len(lines) == 2
    _check(lines[0], case["rows"][0])
    if os.path.exists(REF_TAR):                               # the other golden rows need the reference's tar (build container)
        with tarfile.open(REF_TAR) as tf:
            for m in gold["tar_members"]:
                tf.extract(m, tmp_path, filter="data")
        out = io.StringIO()
        cli_squigglepull.main(["-p", str(tmp_path / "paper_fast5s")] + flags, out=out)
        got = {ln.split("\t", 1)[0]: ln for ln in out.getvalue().split("\n") if ln}
        for name, want in zip(gold["files"][1:], case["rows"][1:]):
            _check(got[name], want)


def test_written_file_reads_back(golden_dir, tmp_path):
    """SquigglePull.py -r -i output is what MotifSeq.py -s / segmenter.py -s consume: column 6 on is the raw signal."""
    d = tmp_path / "f5"
    d.mkdir()
    shutil.copy(os.path.join(golden_dir, "test.fast5"), d / "a.fast5")
    shutil.copy(os.path.join(golden_dir, "test.fast5"), d / "b.fast5")
    out = io.StringIO()
    cli_squigglepull.main(["-p", str(d), "-r", "-i"], out=out)
    p = tmp_path / "sig.tsv"
    p.write_text(out.getvalue())
    want = fast5.read_single_fast5(os.path.join(golden_dir, "test.fast5"))["signal"]
    n = 0
    with tsv.Reader(str(p), 6, pinned=False) as rd:
        for b in rd:
            for i in range(b.n):
                assert b.status[i] == 0 and np.array_equal(b.sig(i), want)
                assert b.head(i)[1] == "db4ae416-40c2-45c2-9cc9-7d49c5711a7c" and b.head(i)[2:] == ["8192.0", "16.0", "1493.94", "4000.0"]
                n += 1
    assert n == 2


def test_arguments_and_missing_directory(tmp_path, capsys):
    with pytest.raises(SystemExit):
        cli_squigglepull.main(["-p", str(tmp_path / "nope")])
    assert "is not an existing directory" in capsys.readouterr().err
    a = cli_squigglepull.build_parser().parse_args(["-p", "x", "-t", "multi", "-v", "-r", "-i"])
    assert (a.path, a.type, a.verbose, a.raw_signal, a.extra_info) == ("x", "multi", True, True, True)
