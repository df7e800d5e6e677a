"""Motifs longer than one pass of the DTW kernel holds (> 1024 points; mlpy takes any length and MotifSeq.py:382-405
expands a fasta to ~9 points per base): the float64 row-block path of libsqk against the CPU oracle, bit for bit."""
import numpy as np
import pytest

import oracle
from squigglekit_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_motif,scale", [(1025, "zscale"), (1500, "medmad"), (2000, "zscale"), (2048, "zscale"), (3100, "zscale")])
def test_long_motif_equals_oracle(ctx, n_motif, scale):
    rng = np.random.default_rng(n_motif)
    motif = np.repeat(rng.normal(0, 1, n_motif // 8 + 1), 8)[:n_motif] + rng.normal(0, 0.05, n_motif)
    lengths = [5000, 0, 3000, 4097, 2600, 700, 4500, 1]          # shorter than the motif, empty, one sample
    if scale == "medmad":
        lengths[-1] = 9          # (a one-sample read has MAD == 0: libsqk reports status -2 there, a disclosed difference)
    sig, off = synth.ragged_reads_np(lengths, None, seed=77 + n_motif)
    # plant a noisy copy of (part of) the motif in two reads so that the alignment is not trivial
    for r, at in ((0, 900), (3, 10)):
        ln = min(n_motif, lengths[r] - at - 1)
        sig[off[r] + at: off[r] + at + ln] = np.clip(np.rint(510 + 80 * motif[:ln] + rng.normal(0, 8, ln)), 1, 1199).astype(np.int16)
    want, want_kept = oracle.motifseq_batch(sig, off, motif, scale=scale, full_matrix=False)
    hits, kept = ctx.motifseq(sig, off, motif, scale=scale)
    assert np.array_equal(kept, want_kept)
    assert np.array_equal(hits["start"][:, 0], want["start"])
    assert np.array_equal(hits["end"][:, 0], want["end"])
    assert np.array_equal(hits["dist"][:, 0], want["dist"], equal_nan=True)


def test_long_and_short_motifs_in_one_call(ctx):
    rng = np.random.default_rng(5)
    short = synth.make_motif()
    long = np.repeat(rng.normal(0, 1, 200), 8)[:1300]
    sig, off, _ = synth.motifseq_reads_np(24, 4096, short)
    hits, _ = ctx.motifseq(sig, off, [short, long, short[:40]], scale="zscale")
    for c, m in enumerate((short, long, short[:40])):
        want, _ = oracle.motifseq_batch(sig, off, m, scale="zscale", full_matrix=False)
        assert np.array_equal(hits["start"][:, c], want["start"]) and np.array_equal(hits["end"][:, c], want["end"])
        assert np.array_equal(hits["dist"][:, c], want["dist"])
