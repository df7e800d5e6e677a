"""GPU parity tests of the dRNA adapter finder (dRNA_segmenter.py slow5 branch): libsqk through
the C ABI vs the golden vectors produced by the reference's own loop and vs the CPU oracle on random reads."""
import io
import json
import os

import numpy as np
import pytest

import oracle
import squigglekit_b200 as sqk
from squigglekit_b200 import synth

pytestmark = pytest.mark.gpu


def as_lists(segs, found):
    return [[int(segs[r, 0]), int(segs[r, 1])] if found[r] > 0 else None for r in range(found.shape[0])]


@pytest.mark.parametrize("mode", ["host", "device"])
def test_adapter_golden(ctx, golden_dir, mode):
    g = np.load(os.path.join(golden_dir, "adapter_inputs.npz"))
    want = json.load(open(os.path.join(golden_dir, "adapter_golden.json")))["segments"]
    if mode == "device":
        import torch
        segs, found = ctx.adapter(torch.from_numpy(g["signals"]).cuda(), torch.from_numpy(g["offsets"]).cuda())
        torch.cuda.synchronize()
        segs, found = segs.cpu().numpy(), found.cpu().numpy()
    else:
        segs, found = ctx.adapter(g["signals"], g["offsets"])
    assert as_lists(segs, found) == want


@pytest.mark.parametrize("seed", range(4))
def test_adapter_random_parameters_vs_oracle(ctx, seed):
    rng = np.random.default_rng(4000 + seed)
    reads = []
    for _ in range(96):
        n = int(rng.integers(0, 14000))
        cut = int(rng.integers(0, max(n, 1)))
        sig = np.where(np.arange(n) < cut, rng.uniform(380, 480), rng.uniform(540, 700)) + rng.normal(0, rng.uniform(3, 30), n)
        sig += np.repeat(rng.normal(0, 30, n // 10 + 1), 10)[:n]
        if n and rng.random() < 0.4:
            sig[rng.integers(0, n, 12)] = rng.choice([-3, 0, 1200, 2500], 12)
        reads.append(np.clip(np.rint(sig), -32768, 32767).astype(np.int16))
    off = np.zeros(len(reads) + 1, np.int64); np.cumsum([r.size for r in reads], out=off[1:])
    sig = np.concatenate(reads)
    cfg = sqk.AdapterConfig(error=int(rng.integers(0, 9)), no_err_thresh=int(rng.integers(0, 4000)), corrector=int(rng.integers(1, 1500)),
                            window=int(rng.integers(1, 300)), seg_dist=int(rng.integers(0, 2000)), t_start=int(rng.integers(0, 2000)),
                            t_end=0, std_scale=float(rng.uniform(0.0, 1.5)),
                            lim_low=int(rng.integers(-10, 300)), lim_hi=int(rng.integers(700, 1400)))
    # seed 3 keeps the statistics slice empty (t_end <= t_start): NaN threshold, nothing may be found
    cfg.t_end = cfg.t_start + (int(rng.integers(200, 5000)) if seed != 3 else -int(rng.integers(0, 50)))
    cfg.t_end = max(cfg.t_end, 0)
    ocfg = oracle.AdapterCfg(cfg.error, cfg.no_err_thresh, cfg.corrector, cfg.window, cfg.seg_dist, cfg.t_start, cfg.t_end, cfg.std_scale)
    wsegs, wfound = oracle.adapter_batch(sig, off, ocfg, cfg.lim_low, cfg.lim_hi)
    segs, found = ctx.adapter(sig, off, cfg)
    assert as_lists(segs, found) == as_lists(wsegs, wfound), cfg
    assert (found.sum() > 0) == (seed != 3)


def test_adapter_cli_on_example_blow5(ctx, golden_dir):
    from squigglekit_b200 import cli_drna_segmenter, slow5
    path = os.path.join(golden_dir, "example.blow5")
    buf = io.StringIO()
    cli_drna_segmenter.main(["-f", path], out=buf)
    want = json.load(open(os.path.join(golden_dir, "adapter_golden.json")))
    recs = list(slow5.read_blow5(path))
    lines = [l for l in buf.getvalue().split("\n") if l]
    exp = ["{}\t{}\t{}".format(r["read_id"], *want["segments"][i]) for i, r in enumerate(recs[:want["n_real_reads"]])
           if want["segments"][i] is not None]
    assert lines == exp


def test_adapter_bad_arguments(ctx):
    sig = np.zeros(100, np.int16); off = np.array([0, 100], np.int64)
    with pytest.raises(sqk.SqkError):
        ctx.adapter(sig, off, sqk.AdapterConfig(corrector=0))
