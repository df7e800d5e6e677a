"""GPU parity tests of the rolling-mean adapter finder (dRNA_segmenter.py TSV branch, :272-326):
libsqk through the C ABI vs the golden vectors produced by the reference's own loop (real pandas) and vs the CPU oracle
on random reads and parameter sets."""
import importlib.util
import io
import json
import os

import numpy as np
import pytest

import oracle
import squigglekit_b200 as sqk

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def inputs():
    spec = importlib.util.spec_from_file_location("rollmean_inputs", os.path.join(ROOT, "tests", "golden", "rollmean_inputs.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def as_lists(segs, found):
    return [[int(segs[r, 0]), int(segs[r, 1])] if found[r] > 0 else None for r in range(found.shape[0])]


@pytest.mark.parametrize("mode", ["host", "device"])
def test_rollmean_golden(ctx, golden_dir, mode):
    want = json.load(open(os.path.join(golden_dir, "rollmean_golden.json")))
    sig, off = inputs().concatenated()
    if mode == "device":
        import torch
        segs, found = ctx.rollmean(torch.from_numpy(sig).cuda(), torch.from_numpy(off).cuda())
        torch.cuda.synchronize()
        segs, found = segs.cpu().numpy(), found.cpu().numpy()
    else:
        segs, found = ctx.rollmean(sig, off)
    assert as_lists(segs, found) == want["segments"]
    segs, found = ctx.rollmean(sig, off[:13], sqk.RollmeanConfig(w=700))
    assert as_lists(segs, found) == want["w700_first12"]


@pytest.mark.parametrize("seed", range(5))
def test_rollmean_random_parameters_vs_oracle(ctx, seed):
    rng = np.random.default_rng(500 + seed)
    reads = []
    for _ in range(48):
        n = int(rng.integers(0, 30000))
        lvl = rng.uniform(350, 700)
        sig = lvl + rng.normal(0, rng.uniform(2, 40), n)
        for _ in range(int(rng.integers(0, 5))):                  # low / high stretches of all lengths
            if n < 10:
                break
            at = int(rng.integers(0, n)); ln = int(rng.integers(1, 9000))
            sig[at:at + ln] += rng.choice([-1, 1]) * rng.uniform(20, 250)
        if rng.random() < 0.3 and n:
            sig[rng.integers(0, n, 25)] = rng.choice([-5, 0, 1200, 5000, -3000], 25)
        reads.append(np.clip(np.rint(sig), -32768, 32767).astype(np.int16))
    off = np.zeros(len(reads) + 1, np.int64)
    np.cumsum([r.size for r in reads], out=off[1:])
    sig = np.concatenate(reads)
    cfg = sqk.RollmeanConfig(w=int(rng.choice([1, 3, 50, 500, 2000, 4000])), seg_dist=int(rng.choice([0, 100, 1500, 5000])),
                             lo_thresh=int(rng.choice([0, 50, 2000])), hi_thresh=int(rng.choice([3000, 200000])),
                             shift=int(rng.choice([0, 1000])), std_factor=float(rng.choice([0.0, 0.5, 1.0, -0.25])),
                             lim_low=int(rng.choice([0, 100, -40000])), lim_hi=int(rng.choice([1200, 900, 40000])))
    ocfg = oracle.RollmeanCfg(cfg.w, cfg.seg_dist, cfg.lo_thresh, cfg.hi_thresh, cfg.shift, cfg.std_factor)
    wsegs, wfound = oracle.rollmean_batch(sig, off, ocfg, cfg.lim_low, cfg.lim_hi)
    segs, found = ctx.rollmean(sig, off, cfg)
    assert np.array_equal(found, wfound)
    assert np.array_equal(segs, wsegs)


def test_rollmean_unaligned_device_view(ctx):
    """Reads that start at odd sample offsets inside a larger device buffer (offsets[0] != 0)."""
    import torch
    sig, off = inputs().concatenated()
    pad = np.r_[np.full(3, 777, np.int16), sig]
    want_segs, want_found = oracle.rollmean_batch(sig, off)
    segs, found = ctx.rollmean(torch.from_numpy(pad).cuda(), torch.from_numpy(off + 3).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(found.cpu().numpy(), want_found) and np.array_equal(segs.cpu().numpy(), want_segs)


def test_rollmean_cli_tsv(ctx, golden_dir, tmp_path):
    from squigglekit_b200 import cli_drna_segmenter
    want = json.load(open(os.path.join(golden_dir, "rollmean_golden.json")))["segments"]
    reads = inputs().reads()[:10]
    path = tmp_path / "sig.tsv"
    with open(path, "w") as fh:
        for i, r in enumerate(reads):
            fh.write("\t".join([f"f{i}.fast5", f"read{i}", "a", "b"] + [str(int(v)) for v in r]) + "\n")
    buf = io.StringIO()
    cli_drna_segmenter.main(["-s", str(path)], out=buf)
    exp = [f"f{i}.fast5\tread{i}\t{w[0]}\t{w[1]}" for i, w in enumerate(want[:10]) if w is not None]
    assert [l for l in buf.getvalue().split("\n") if l] == exp


def test_rollmean_bad_arguments(ctx):
    sig = np.zeros(100, np.int16); off = np.array([0, 100], np.int64)
    with pytest.raises(sqk.SqkError):
        ctx.rollmean(sig, off, sqk.RollmeanConfig(w=0))
    with pytest.raises(sqk.SqkError):
        ctx.rollmean(sig, off, sqk.RollmeanConfig(w=70000))
