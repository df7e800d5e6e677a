"""The statistics kernel's rare paths (csrc/sqk_stats3.cuh), each forced on purpose and compared with the CPU oracle bit
for bit through all three modes: leaves with outliers inside (compacted in place), more outliers than the exception list
holds (whole read compacted in place), bursts, outliers on leaf boundaries and read edges, unaligned read starts, reads of
every length around the leaf / slot boundaries."""
import numpy as np
import pytest

import oracle
import squigglekit_b200 as sqk
from squigglekit_b200 import synth

pytestmark = pytest.mark.gpu


def _reads_with_outliers(rng, counts, n=4096, burst=False):
    reads = []
    for k in counts:
        s = np.clip(np.rint(rng.normal(500, 70, n)), 1, 899).astype(np.int16)
        if k:
            if burst:
                at = int(rng.integers(0, n - k)) if k < n else 0
                pos = np.arange(at, at + min(k, n))
            else:
                pos = rng.choice(n, size=min(k, n), replace=False)
            s[pos] = rng.choice([-7, 0, 1200, 1500, 32767, -32768], size=pos.size)
        reads.append(s)
    off = np.zeros(len(reads) + 1, np.int64)
    np.cumsum([r.size for r in reads], out=off[1:])
    return np.concatenate(reads), off


@pytest.mark.parametrize("burst", [False, True])
@pytest.mark.parametrize("scale", ["zscale", "medmad"])
def test_outlier_counts_motifseq(ctx, scale, burst):
    rng = np.random.default_rng(31 + burst)
    counts = [0, 1, 2, 3, 8, 31, 32, 33, 34, 64, 65, 200, 1000, 4000, 4095, 4096] * 2
    sig, off = _reads_with_outliers(rng, counts, burst=burst)
    motif = synth.make_motif()
    want, want_kept = oracle.motifseq_batch(sig, off, motif, scale=scale, full_matrix=False)
    hits, kept = ctx.motifseq(sig, off, motif, scale=scale)
    assert np.array_equal(kept, want_kept)
    ok = want["start"] >= 0
    # (medmad on a read whose MAD is 0 is a disclosed difference: status -2 instead of the reference's NaN row)
    deg = hits["start"][:, 0] == -2
    assert np.array_equal(hits["start"][:, 0][~deg], want["start"][~deg])
    assert np.array_equal(hits["end"][:, 0][~deg], want["end"][~deg])
    assert np.array_equal(hits["dist"][:, 0][~deg & ok], want["dist"][~deg & ok])


@pytest.mark.parametrize("burst", [False, True])
def test_outlier_counts_segmenter(ctx, burst):
    rng = np.random.default_rng(41 + burst)
    counts = [0, 1, 2, 5, 31, 32, 33, 40, 100, 700, 4000, 4096] * 2
    sig, off = _reads_with_outliers(rng, counts, burst=burst)
    for r in range(len(counts)):                                  # a stall-like plateau so that there are segments to find
        sig[off[r] + 20: off[r] + 420] = np.clip(np.rint(rng.normal(505, 4, 400)), 1, 899).astype(np.int16)
    cfg = sqk.SegConfig(max_segs=32)
    want, wn = oracle.segmenter_batch(sig, off, oracle.SegCfg(), 0, 900, 0, 32)
    segs, nsegs = ctx.segmenter(sig, off, cfg)
    assert np.array_equal(nsegs, wn)
    m = np.arange(32)[None, :, None] < wn[:, None, None]
    assert np.array_equal(np.where(m, segs, 0), np.where(m, want, 0))


def test_every_length_and_alignment(ctx):
    """Reads of every length 1..300 and around the leaf / slot boundaries, concatenated so that every 16-byte alignment
    occurs, a few outliers sprinkled in: zscale statistics through the DTW result, segmenter thresholds through the segments."""
    rng = np.random.default_rng(53)
    lengths = list(range(1, 300)) + [1023, 1024, 1025, 2047, 2048, 2049, 4088, 4090, 4095, 4096, 4097, 4104, 5000, 8175, 8176, 8177, 9000]
    reads = []
    for n in lengths:
        s = np.clip(np.rint(rng.normal(500, 70, n)), 1, 899).astype(np.int16)
        if n > 5:
            s[rng.choice(n, size=min(3, n // 3), replace=False)] = rng.choice([0, 1300, -4])
        reads.append(s)
    off = np.zeros(len(reads) + 1, np.int64)
    np.cumsum([r.size for r in reads], out=off[1:])
    sig = np.concatenate(reads)
    motif = synth.make_motif()[:24]
    want, want_kept = oracle.motifseq_batch(sig, off, motif, scale="zscale", full_matrix=False)
    hits, kept = ctx.motifseq(sig, off, motif, scale="zscale")
    assert np.array_equal(kept, want_kept)
    assert np.array_equal(hits["start"][:, 0], want["start"]) and np.array_equal(hits["end"][:, 0], want["end"])
    assert np.array_equal(hits["dist"][:, 0], want["dist"], equal_nan=True)
    wsegs, wn = oracle.segmenter_batch(sig, off, oracle.SegCfg(), 0, 900, 0, 16)
    segs, nsegs = ctx.segmenter(sig, off, sqk.SegConfig(max_segs=16))
    assert np.array_equal(nsegs, wn)
    m = np.arange(16)[None, :, None] < np.minimum(wn, 16)[:, None, None]
    assert np.array_equal(np.where(m, segs, 0), np.where(m, wsegs, 0))
