"""Randomised GPU-vs-oracle tests: random motifs, read shapes, outlier windows and segmenter parameters.
Seeds are fixed, so failures reproduce; the bar is the same as everywhere else (bit-exact)."""
import numpy as np
import pytest

import oracle
import squigglekit_b200 as sqk
from squigglekit_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(6))
def test_motifseq_random_shapes(ctx, seed):
    rng = np.random.default_rng(1000 + seed)
    n_motif = int(rng.choice([1, 3, 6, 11, 24, 47, 80, 97, 163, 255, 400, 777]))
    levels = rng.standard_normal(n_motif // int(rng.integers(1, 9)) + 1)
    motif = np.repeat(levels, int(rng.integers(1, 9)))[:n_motif].astype(np.float64)
    if motif.size < n_motif:
        motif = np.concatenate([motif, rng.standard_normal(n_motif - motif.size)])
    lengths = [int(v) for v in rng.integers(0, 6000, 40)]
    sig, off = synth.ragged_reads_np(lengths, motif if n_motif < 500 else None, seed=seed)
    lo = int(rng.integers(-50, 450))
    hi = int(rng.integers(560, 1500))
    scale = ["zscale", "medmad", "none"][seed % 3]
    want, kept_w = oracle.motifseq_batch(sig, off, motif, lo=lo, hi=hi, scale=scale, full_matrix=False)
    hits, kept = ctx.motifseq(sig, off, motif, scale=scale, scale_low=lo, scale_hi=hi)
    ok = (kept_w > 0) & np.isfinite(want["dist"])
    assert np.array_equal(kept, kept_w)
    h = hits[:, 0]
    assert np.array_equal(h["start"][ok], want["start"][ok]) and np.array_equal(h["end"][ok], want["end"][ok])
    assert np.array_equal(h["dist"][ok], want["dist"][ok])
    assert (h["start"][~ok] < 0).all()


@pytest.mark.parametrize("seed", range(8))
def test_segmenter_random_parameters(ctx, seed):
    rng = np.random.default_rng(2000 + seed)
    sig, off = synth.segmenter_reads_np(96, int(rng.integers(500, 5000)), seed=seed)
    extra, eoff = synth.ragged_reads_np([int(v) for v in rng.integers(0, 400, 12)], seed=seed)
    sig = np.concatenate([sig, extra])
    off = np.concatenate([off, off[-1] + eoff[1:]])
    cfg = sqk.SegConfig(error=int(rng.integers(0, 12)), corrector=int(rng.integers(0, 60)), window=int(rng.integers(1, 200)),
                        seg_dist=int(rng.integers(0, 300)), std_scale=float(rng.uniform(0.2, 1.6)),
                        stall_len=float(rng.uniform(0.0, 1.0)), lim_low=int(rng.integers(-10, 420)),
                        lim_hi=int(rng.integers(600, 1300)), Num=int(rng.choice([0, 0, 300, 2000, -50])), max_segs=1024)
    ocfg = oracle.SegCfg(cfg.error, cfg.corrector, cfg.window, cfg.seg_dist, cfg.std_scale, cfg.stall_len)
    n = off.size - 1
    po = rng.integers(-30, 40, n).astype(float)
    ps = np.round(rng.uniform(1100, 1600, n), 2) / 8192.0
    for pa in (False, True):
        if pa:
            cfg.lim_low, cfg.lim_hi = int(rng.integers(-5, 70)), int(rng.integers(110, 300))
            want, want_n = oracle.segmenter_batch_pa(sig, off, po, ps, ocfg, cfg.lim_low, cfg.lim_hi, cfg.Num, cfg.max_segs)
            segs, nsegs = ctx.segmenter(sig, off, cfg, pa_offset=po, pa_scale=ps)
        else:
            want, want_n = oracle.segmenter_batch(sig, off, ocfg, cfg.lim_low, cfg.lim_hi, cfg.Num, cfg.max_segs)
            segs, nsegs = ctx.segmenter(sig, off, cfg)
        assert np.array_equal(nsegs, want_n), (cfg, pa, np.nonzero(nsegs != want_n)[0][:8])
        assert (want_n <= cfg.max_segs).all()
        for r in range(n):
            assert np.array_equal(segs[r, :nsegs[r]], want[r, :want_n[r]]), (cfg, pa, r)
