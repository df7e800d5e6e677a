"""GPU parity tests of the exact two-pass DTW plan (float32 lower-bound scan + float64 windows):
forced on and off through the C ABI, it must return the oracle's start / end / dist bit for bit -- including the
paths that only trigger on awkward reads (short-read jobs, cluster overflow, tainted windows, full-length fallback)."""
import os

import numpy as np
import pytest

import oracle
import squigglekit_b200 as sqk
from squigglekit_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture()
def planned(ctx):
    """ctx whose plan is restored to auto afterwards."""
    yield ctx
    ctx.set_dtw_plan("auto")
    ctx.set_dtw_lanes(0)
    os.environ.pop("SQK_LB_WINDOW", None)


def check(ctx, sig, off, motif, scale="zscale", lo=0, hi=1200, what=""):
    want, kept_w = oracle.motifseq_batch(sig, off, motif, lo=lo, hi=hi, scale=scale, full_matrix=False)
    hits, kept = ctx.motifseq(sig, off, motif, scale=scale, scale_low=lo, scale_hi=hi)
    h = hits[:, 0]
    ok = (kept_w > 0) & np.isfinite(want["dist"])
    assert np.array_equal(kept, kept_w), what
    bad = np.nonzero((h["start"][ok] != want["start"][ok]) | (h["end"][ok] != want["end"][ok]) | (h["dist"][ok] != want["dist"][ok]))[0]
    assert bad.size == 0, f"{what}: {bad.size} reads differ, first {bad[:5]}: got {h[ok][bad[:3]]} want {want[ok][bad[:3]]}"
    assert (h["start"][~ok] < 0).all(), what
    return hits


@pytest.mark.parametrize("plan", ["single_pass", "two_pass"])
@pytest.mark.parametrize("scale", ["zscale", "medmad"])
def test_plans_match_oracle(planned, plan, scale):
    planned.set_dtw_plan(plan)
    motif = synth.make_motif()
    sig, off, _ = synth.motifseq_reads_np(512, 4096, motif)
    planned.enable_timing(True); planned.timing(reset=True)
    check(planned, sig, off, motif, scale, what=f"{plan}/{scale}")
    kt = planned.timing(reset=True); planned.enable_timing(False)
    if plan == "two_pass":
        assert kt["dtw_lb"]["launches"] >= 1 and kt["dtw_win"]["launches"] >= 1 and kt["dtw"]["launches"] == 0
    else:
        assert kt["dtw"]["launches"] >= 1 and kt["dtw_lb"]["launches"] == 0


def test_two_pass_proves_nearly_every_benchmark_read(planned):
    import torch
    planned.set_dtw_plan("two_pass")
    motif = synth.make_motif()
    sig, off, _ = synth.motifseq_reads_np(2048, 4096, motif)
    hits_t, _ = planned.motifseq(torch.from_numpy(sig).cuda(), torch.from_numpy(off).cuda(), motif, scale="zscale", max_read_len=4096)
    torch.cuda.synchronize()
    pc = planned.plan_counters()
    assert 2048 <= pc["windows"] <= 2048 * 1.2, pc
    assert pc["fallback_reads"] <= 20, pc
    want, _ = oracle.motifseq_batch(sig, off, motif, scale="zscale", full_matrix=False)
    h = sqk.hits_from_torch(hits_t)[:, 0]
    assert np.array_equal(h["start"], want["start"]) and np.array_equal(h["end"], want["end"]) and np.array_equal(h["dist"], want["dist"])


def test_two_pass_ragged_and_degenerate_reads(planned):
    planned.set_dtw_plan("two_pass")
    motif = synth.make_motif()
    lengths = [0, 1, 5, 79, 80, 81, 200, 543, 544, 545, 1087, 1088, 3000, 9000, 70000, 2, 4097, 0, 12345]
    sig, off = synth.ragged_reads_np(lengths, motif, seed=3)
    # an all-outlier read and a constant read (MAD == 0 under medmad)
    extra = np.concatenate([np.full(2000, 3000, np.int16), np.full(3000, 500, np.int16)])
    sig = np.concatenate([sig, extra]); off = np.concatenate([off, off[-1] + np.array([2000, 5000])])
    for scale in ("zscale", "medmad"):
        check(planned, sig, off, motif, scale, what=f"ragged/{scale}")
    # unaligned view: offsets[0] != 0 and an odd start address
    check(planned, sig[3:], off[3:] - 3, motif, "zscale", what="unaligned")


@pytest.mark.parametrize("n_motif", [5, 24, 47, 80, 97, 163, 400, 777, 1024])
def test_two_pass_motif_lengths(planned, n_motif):
    planned.set_dtw_plan("two_pass")
    rng = np.random.default_rng(n_motif)
    motif = np.repeat(rng.standard_normal(n_motif // 5 + 1), 5)[:n_motif].astype(np.float64)
    lengths = [int(v) for v in rng.integers(4 * n_motif, 4 * n_motif + 9000, 24)]
    sig, off = synth.ragged_reads_np(lengths, motif if n_motif < 500 else None, seed=n_motif)
    check(planned, sig, off, motif, ["zscale", "medmad"][n_motif % 2], what=f"N={n_motif}")


@pytest.mark.parametrize("lanes", [4, 8, 16, 32])
def test_two_pass_lane_layouts(planned, lanes):
    planned.set_dtw_plan("two_pass")
    planned.set_dtw_lanes(lanes)
    motif = synth.make_motif()[: {4: 80, 8: 77, 16: 80, 32: 70}[lanes]]
    sig, off, _ = synth.motifseq_reads_np(96, 3000, motif)
    check(planned, sig, off, motif, "zscale", what=f"lanes={lanes}")


def test_two_pass_tie_heavy_integer_reads(planned):
    planned.set_dtw_plan("two_pass")
    rng = np.random.default_rng(17)
    motif = np.repeat(rng.integers(495, 525, 10), 8).astype(np.float64)     # raw units, scale "none"
    reads = [np.repeat(rng.integers(490, 530, 700), 6)[: int(rng.integers(2500, 4200))].astype(np.int16) for _ in range(64)]
    reads.append(np.full(5000, 500, np.int16))                               # every column ties
    r2 = np.full(6000, 500, np.int16); r2[::2] += 1; reads.append(r2)        # two-level plateau: huge clusters
    off = np.zeros(len(reads) + 1, np.int64); np.cumsum([r.size for r in reads], out=off[1:])
    sig = np.concatenate(reads)
    check(planned, sig, off, motif, "none", what="ties/none")
    check(planned, sig, off, synth.make_motif(), "zscale", what="ties/zscale")


def test_two_pass_small_window_second_attempt_and_fallback(planned):
    """A window far too small for the alignment taints the minimum.  With the second attempt (windows of W2 columns) nearly
    every such read is settled there; without it (W2 = 0), or with a second window that is too small as well, every such
    read must be re-run in full.  Same bits every way."""
    import torch
    planned.set_dtw_plan("two_pass")
    motif = synth.make_motif()
    sig, off, _ = synth.motifseq_reads_np(256, 4096, motif)
    dsig, doff = torch.from_numpy(sig).cuda(), torch.from_numpy(off).cuda()
    try:
        os.environ["SQK_LB_WINDOW"] = "12"
        check(planned, sig, off, motif, "zscale", what="window=12, default second attempt")
        planned.motifseq(dsig, doff, motif, scale="zscale", max_read_len=4096)
        torch.cuda.synchronize()
        pc = planned.plan_counters()
        assert pc["second_attempt_windows"] >= 128 and pc["fallback_reads"] <= 8, pc
        os.environ["SQK_LB_WINDOW2"] = "0"
        check(planned, sig, off, motif, "zscale", what="window=12, no second attempt")
        planned.motifseq(dsig, doff, motif, scale="zscale", max_read_len=4096)
        torch.cuda.synchronize()
        pc = planned.plan_counters()
        assert pc["second_attempt_windows"] == 0 and pc["fallback_reads"] >= 128, pc
        os.environ["SQK_LB_WINDOW2"] = "20"
        check(planned, sig, off, motif, "zscale", what="window=12, second attempt of 20 columns")
        planned.motifseq(dsig, doff, motif, scale="zscale", max_read_len=4096)
        torch.cuda.synchronize()
        pc = planned.plan_counters()
        assert pc["second_attempt_windows"] >= 128 and pc["fallback_reads"] >= 100, pc
    finally:
        os.environ.pop("SQK_LB_WINDOW", None)
        os.environ.pop("SQK_LB_WINDOW2", None)


def test_two_pass_multiple_models_and_outlier_windows(planned):
    planned.set_dtw_plan("two_pass")
    rng = np.random.default_rng(23)
    models = [synth.make_motif(), np.repeat(rng.standard_normal(21), 7)[:140], rng.standard_normal(33)]
    sig, off, _ = synth.motifseq_reads_np(128, 5000, models[0])
    hits, kept = planned.motifseq(sig, off, models, scale="medmad", scale_low=300, scale_hi=700)
    for m, model in enumerate(models):
        want, kept_w = oracle.motifseq_batch(sig, off, np.ascontiguousarray(model), lo=300, hi=700, scale="medmad", full_matrix=False)
        assert np.array_equal(kept, kept_w)
        h = hits[:, m]
        assert np.array_equal(h["start"], want["start"]) and np.array_equal(h["end"], want["end"]) and np.array_equal(h["dist"], want["dist"]), m


def test_two_pass_reads_beyond_proof_length(planned):
    """Reads longer than SQK_LB_MAX_LEN (2^18 kept samples) are outside the lower-bound proof: one full-length job."""
    planned.set_dtw_plan("two_pass")
    motif = synth.make_motif()
    sig, off = synth.ragged_reads_np([300000, 5000], motif, seed=9)
    check(planned, sig, off, motif, "zscale", what="long")
