"""GPU end-to-end tests of the drop-in command lines: their stdout equals what the reference scripts
printed (golden rows generated with the reference's own get_region_multi / get_segs)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(script, *args):
    p = subprocess.run([sys.executable, os.path.join(ROOT, script), *args], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout, p.stderr


def write_model(tmp_path, golden_dir):
    exp = json.load(open(os.path.join(golden_dir, "example_expected.json")))
    m = tmp_path / "CATCTATCCAGGGTTAAATT.model"
    m.write_text(exp["model_text"])
    return str(m), exp


@pytest.mark.parametrize("scale", ["zscale", "medmad"])
def test_motifseq_fast5_dir_config1(golden_dir, tmp_path, scale):
    """BASELINE config 1: MotifSeq.py -p <dir with test.fast5> -m example .model."""
    model, exp = write_model(tmp_path, golden_dir)
    d = tmp_path / "f5"
    d.mkdir()
    os.symlink(os.path.join(golden_dir, "test.fast5"), d / "test.fast5")
    out, err = run("MotifSeq.py", "-p", str(d), "-m", model, "-l", scale)
    lines = out.rstrip("\n").split("\n")
    assert lines[0].split("\t")[:4] == ["fast5", "readID", "model", "start"]
    want = exp["tsv"][scale].rstrip("\n").split("\t")
    got = lines[1].split("\t")
    assert got[1] == "b'{}'".format(want[1])          # h5py hands the reference bytes: it prints b'...'
    assert got[0] == want[0] and got[2:] == want[2:]
    assert "preliminary experimental modeling only" in err


def test_motifseq_signal_tsv_with_extract(golden_dir, tmp_path):
    model, exp = write_model(tmp_path, golden_dir)
    ex = np.load(os.path.join(golden_dir, "example_read.npz"), allow_pickle=True)
    tsv = tmp_path / "sig.tsv"
    with open(tsv, "w") as f:
        f.write("\t".join(["test.fast5", str(ex["read_id"])] + ["x"] * 6 + [str(int(v)) for v in ex["raw"]]) + "\n")
    out, _ = run("MotifSeq.py", "-s", str(tsv), "-m", model, "-l", "zscale", "-x")
    lines = out.rstrip("\n").split("\n")
    assert lines[0].endswith("normalised_signal")
    assert lines[1] == exp["tsv"]["zscale_x"].rstrip("\n")


def test_segmenter_signal_tsv(golden_dir, tmp_path):
    g = np.load(os.path.join(golden_dir, "segmenter_inputs.npz"))
    gold = json.load(open(os.path.join(golden_dir, "segmenter_golden.json")))
    sig, off = g["signals"], g["offsets"]
    tsv = tmp_path / "sig.tsv"
    with open(tsv, "w") as f:
        for r in range(off.size - 1):
            f.write("\t".join([f"read{r}.fast5", "id", "a", "b"] + [str(int(v)) for v in sig[off[r]:off[r + 1]]]) + "\n")
    for case, flags in ((gold[0], []), (gold[1], ["-k", "-u"]), (gold[2], ["-k", "-g", "-u", "-b", "800"])):
        out, err = run("segmenter.py", "-s", str(tsv), *flags)
        want = []
        for r, rec in enumerate(case["reads"]):
            if rec is None:
                continue
            segs = rec["tested"] if rec["tested"] is not None else rec["segs"]
            if segs:
                want.append(f"read{r}.fast5\t" + ",".join(f"{a},{b}" for a, b in segs))
        assert out.rstrip("\n").split("\n") == want, flags
        assert err.endswith("Done")


def test_segmenter_single_fast5(golden_dir):
    exp = json.load(open(os.path.join(golden_dir, "example_expected.json")))
    path = os.path.join(golden_dir, "test.fast5")
    out, _ = run("segmenter.py", "-i", path, "--single", "--raw_signal")
    assert out.rstrip("\n") == path + "\t" + ",".join(f"{a},{b}" for a, b in exp["segs_raw"])


def test_segmenter_single_fast5_default_is_pa(golden_dir):
    """Without --raw_signal the reference segments pA rounded to 2 decimals (segmenter.py:345-349)."""
    exp = json.load(open(os.path.join(golden_dir, "example_expected.json")))
    path = os.path.join(golden_dir, "test.fast5")
    out, _ = run("segmenter.py", "-i", path, "--single")
    assert out.rstrip("\n") == path + "\t" + ",".join(f"{a},{b}" for a, b in exp["segs_pA"])
    assert exp["segs_pA"] != exp["segs_raw"]


def test_clis_accept_float_pa_tsv(golden_dir, tmp_path):
    """SquigglePull without -r prints pA floats; both tools must give what the reference's float branch gives."""
    g = np.load(os.path.join(golden_dir, "float_signal_golden.npz"))
    want_segs = json.load(open(os.path.join(golden_dir, "float_signal_segs.json")))
    sig, off = g["signals"], g["offsets"]
    model, exp = write_model(tmp_path, golden_dir)
    bait = tmp_path / "motif80.tsv"
    from squigglekit_b200 import synth
    bait.write_text("m\t10\tx\t" + "\t".join(repr(float(v)) for v in synth.make_motif()) + "\n")
    ms_tsv, sg_tsv = tmp_path / "ms.tsv", tmp_path / "sg.tsv"
    with open(ms_tsv, "w") as f1, open(sg_tsv, "w") as f2:
        for r in range(off.size - 1):
            vals = [repr(float(v)) for v in sig[off[r]:off[r + 1]]]
            if not vals:
                continue
            f1.write("\t".join([f"read{r}.fast5", f"id{r}"] + ["x"] * 6 + vals) + "\n")
            f2.write("\t".join([f"read{r}.fast5", "id", "a", "b"] + vals) + "\n")
    out, _ = run("MotifSeq.py", "-s", str(ms_tsv), "-m", str(bait), "-l", "zscale", "-scale_hi", "300", "-scale_low", "40")
    rows = {l.split("\t")[0]: l.split("\t") for l in out.rstrip("\n").split("\n")[1:]}
    for r in range(off.size - 1):
        if g["zscale_start"][r] == -9 or off[r + 1] == off[r]:
            continue
        f = rows[f"read{r}.fast5"]
        assert (int(f[3]), int(f[4]), f[6]) == (int(g["zscale_start"][r]), int(g["zscale_end"][r]), repr(float(g["zscale_dist"][r])))
    out, _ = run("segmenter.py", "-s", str(sg_tsv), "-lim_hi", "160", "-lim_low", "30")
    got = dict(l.split("\t") for l in out.rstrip("\n").split("\n") if l)
    for r, w in enumerate(want_segs["segs"]):
        key = f"read{r}.fast5"
        if w:
            assert got[key] == ",".join(f"{a},{b}" for a, b in w), r
        else:
            assert key not in got


def test_clis_accept_blow5(golden_dir, tmp_path):
    model, exp = write_model(tmp_path, golden_dir)
    path = os.path.join(golden_dir, "example.blow5")
    out, _ = run("MotifSeq.py", "--slow5", path, "-m", model, "-l", "medmad")
    got = out.rstrip("\n").split("\n")[1].split("\t")
    want = exp["tsv"]["medmad"].rstrip("\n").split("\t")
    assert got[1:] == want[1:]
    out, _ = run("segmenter.py", "--slow5", path)
    assert out.rstrip("\n").split("\t")[1] == ",".join(f"{a},{b}" for a, b in exp["segs_pA"])
    out, _ = run("segmenter.py", "--slow5", path, "--raw_signal")
    assert out.rstrip("\n").split("\t")[1] == ",".join(f"{a},{b}" for a, b in exp["segs_raw"])
