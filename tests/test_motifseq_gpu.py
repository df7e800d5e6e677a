"""GPU parity tests for the MotifSeq hot path: libsqk (through the C ABI) vs the CPU oracle and the
committed golden fixtures.  Bar: start/end bit-exact AND dist bit-exact in fp64 mode (every fp64
op is a single IEEE rounding in the same order as the reference's C loop); fp32 mode: dist within
1e-4 relative, index mismatch rate reported and bounded."""
import json
import os

import numpy as np
import pytest

import oracle
import squigglekit_b200 as sqk
from squigglekit_b200 import synth

pytestmark = pytest.mark.gpu


def oracle_hits(signals, offsets, model, scale, lo=0, hi=1200):
    hits, kept = oracle.motifseq_batch(signals, offsets, model, lo=lo, hi=hi, scale=scale, full_matrix=False)
    return hits, kept


def assert_hits_equal(got, want, kept_got=None, kept_want=None, what=""):
    assert np.array_equal(got["start"], want["start"]), f"{what}: start differs at {np.nonzero(got['start'] != want['start'])[0][:10]}"
    assert np.array_equal(got["end"], want["end"]), f"{what}: end differs at {np.nonzero(got['end'] != want['end'])[0][:10]}"
    a, b = got["dist"], want["dist"]
    same = (a == b) | (np.isnan(a) & np.isnan(b))
    assert same.all(), f"{what}: dist differs (bitwise) at {np.nonzero(~same)[0][:10]}"
    if kept_got is not None:
        assert np.array_equal(kept_got, kept_want), f"{what}: n_kept differs"


@pytest.mark.parametrize("scale", ["zscale", "medmad"])
@pytest.mark.parametrize("mname", ["motif80", "example163"])
def test_golden_set(ctx, golden_dir, scale, mname):
    g = np.load(os.path.join(golden_dir, "motifseq_golden.npz"))
    hits, kept = ctx.motifseq(g["signals"], g["offsets"], g["model_" + mname], scale=scale)
    key = f"{mname}_{scale}"
    want_start, want_end, want_dist = g[key + "_start"], g[key + "_end"], g[key + "_dist"]
    # reads whose MAD is 0 (constant / single-sample reads under medmad) have an all-NaN/inf normalised
    # signal in the reference; libsqk reports status -2 for them instead of DTW-ing NaNs
    pinned = (want_start != -2) & np.isfinite(want_dist)
    assert np.array_equal(hits["start"][pinned, 0], want_start[pinned])
    assert np.array_equal(hits["end"][pinned, 0], want_end[pinned])
    assert np.array_equal(hits["dist"][pinned, 0], want_dist[pinned])
    assert np.array_equal(kept[pinned], g[key + "_kept"][pinned])
    if (~pinned).any():
        assert (hits["start"][~pinned, 0] == -2).all() and np.isnan(hits["dist"][~pinned, 0]).all()


@pytest.mark.parametrize("scale", ["zscale", "medmad"])
def test_example_read(ctx, golden_dir, scale):
    """BASELINE config 1: example/CATCTATCCAGGGTTAAATT.model vs example/test.fast5."""
    ex = np.load(os.path.join(golden_dir, "example_read.npz"), allow_pickle=True)
    want = json.load(open(os.path.join(golden_dir, "example_expected.json")))["tsv"][scale].split("\t")
    raw = ex["raw"]
    offs = np.array([0, raw.size], dtype=np.int64)
    hits, kept = ctx.motifseq(raw, offs, ex["model"], scale=scale)
    assert int(hits["start"][0, 0]) == int(want[3])
    assert int(hits["end"][0, 0]) == int(want[4])
    assert repr(float(hits["dist"][0, 0])) == want[6]
    assert int(kept[0]) == 36977


@pytest.mark.parametrize("scale", ["zscale", "medmad"])
@pytest.mark.parametrize("mode", ["host", "device"])
def test_synthetic_vs_oracle(ctx, scale, mode):
    motif = synth.make_motif()
    sig, off, _ = synth.motifseq_reads_np(384, 4096, motif)
    want, kept_w = oracle_hits(sig, off, motif, scale)
    if mode == "device":
        import torch
        hits_t, kept_t = ctx.motifseq(torch.from_numpy(sig).cuda(), torch.from_numpy(off).cuda(), motif, scale=scale,
                                      max_read_len=4096)
        torch.cuda.synchronize()
        hits, kept = sqk.hits_from_torch(hits_t), kept_t.cpu().numpy()
    else:
        hits, kept = ctx.motifseq(sig, off, motif, scale=scale)
    assert_hits_equal(hits[:, 0], want, kept, kept_w, f"{scale}/{mode}")


def test_ragged_and_degenerate_reads(ctx):
    motif = synth.make_motif()
    lengths = [0, 1, 2, 3, 7, 8, 9, 15, 16, 17, 79, 80, 81, 127, 128, 129, 255, 1000, 4095, 4097, 0, 0, 5000, 33, 20000, 64, 6]
    sig, off = synth.ragged_reads_np(lengths, motif)
    # an all-outlier read and a read that keeps a single sample
    extra = np.array([0, -5, 1200, 1500, 3000], dtype=np.int16)
    one = np.array([1300, 1300, 500, 1400], dtype=np.int16)
    sig = np.concatenate([sig, extra, one])
    off = np.concatenate([off, [off[-1] + extra.size, off[-1] + extra.size + one.size]])
    for scale in ("zscale", "medmad"):
        want, kept_w = oracle_hits(sig, off, motif, scale)
        hits, kept = ctx.motifseq(sig, off, motif, scale=scale)
        h = hits[:, 0]
        empty = kept_w == 0
        assert (h["start"][empty] == -1).all() and (h["end"][empty] == -1).all() and np.isnan(h["dist"][empty]).all()
        # medmad with MAD == 0 (short / constant reads) is reported as -2, the oracle yields NaN there
        degenerate = (~empty) & ~np.isfinite(want["dist"])
        if scale == "medmad":
            assert (h["start"][degenerate] == -2).all()
        else:
            assert not degenerate.any()
        ok = ~(empty | degenerate)
        assert_hits_equal(h[ok], want[ok], kept, kept_w, f"ragged/{scale}")


def test_unaligned_offsets(ctx):
    """offsets[0] != 0 and reads starting at odd sample positions (16-byte block hulls cross reads)."""
    motif = synth.make_motif()
    sig, off = synth.ragged_reads_np([777, 1001, 13, 2048, 333, 4099], motif)
    pad = np.full(5, 400, dtype=np.int16)
    sig2 = np.concatenate([pad, sig, pad])
    off2 = off + 5
    want, kept_w = oracle_hits(sig, off, motif, "zscale")
    hits, kept = ctx.motifseq(sig2, off2, motif, scale="zscale")
    assert_hits_equal(hits[:, 0], want, kept, kept_w, "unaligned host")
    import torch
    dsig = torch.from_numpy(sig2).cuda()
    for shift in (0, 1, 3):       # device buffers whose base address is not 16-byte aligned
        view = dsig[shift:]
        hits_t, kept_t = ctx.motifseq(view, torch.from_numpy(off2 - shift).cuda(), motif, scale="zscale")
        torch.cuda.synchronize()
        assert_hits_equal(sqk.hits_from_torch(hits_t)[:, 0], want, kept_t.cpu().numpy(), kept_w, f"unaligned device shift {shift}")


@pytest.mark.parametrize("n_motif", [1, 2, 3, 4, 5, 7, 9, 16, 33, 79, 80, 81, 100, 160, 161, 163, 200, 320, 321, 500, 640, 641, 1000, 1024])
def test_motif_lengths(ctx, n_motif):
    rng = np.random.default_rng(n_motif)
    motif = np.repeat(rng.standard_normal((n_motif + 3) // 4), 4)[:n_motif].astype(np.float64)
    sig, off = synth.ragged_reads_np([1500, 400, n_motif, max(1, n_motif - 1), 2500, 64, 3000, 900], motif)
    want, kept_w = oracle_hits(sig, off, motif, "zscale")
    hits, kept = ctx.motifseq(sig, off, motif, scale="zscale")
    assert_hits_equal(hits[:, 0], want, kept, kept_w, f"N={n_motif}")


@pytest.mark.parametrize("lanes", [4, 8, 16, 32])
def test_lane_layouts_agree(ctx, lanes):
    """Every lanes-per-read layout of the kernel must give the oracle's answer (N=80 and N=163)."""
    for n_motif in (80, 163):
        motif = synth.make_motif(n_levels=(n_motif + 7) // 8, dwell=8)[:n_motif]
        sig, off, _ = synth.motifseq_reads_np(96, 2048, motif, seed=77 + lanes)
        want, kept_w = oracle_hits(sig, off, motif, "zscale")
        ctx.set_dtw_lanes(lanes)
        try:
            hits, kept = ctx.motifseq(sig, off, motif, scale="zscale")
        except sqk.SqkError as e:
            assert e.code == -4      # this layout cannot hold that motif: refused loudly, not mis-computed
            continue
        finally:
            ctx.set_dtw_lanes(0)
        assert_hits_equal(hits[:, 0], want, kept, kept_w, f"lanes={lanes} N={n_motif}")


def test_tie_heavy_integer_reads(ctx):
    """Few distinct levels, no noise, integer-valued motif: exact ties in min3 and in the last-row argmin
    pin the tie order (diag, then left, then up; first argmin)."""
    rng = np.random.default_rng(3)
    motif = np.repeat(rng.integers(-2, 3, 12), 5).astype(np.float64)
    reads = []
    for r in range(64):
        lv = rng.integers(480, 540, rng.integers(20, 200)) // 10 * 10
        reads.append(np.repeat(lv, rng.integers(1, 12)).astype(np.int16))
    off = np.zeros(len(reads) + 1, dtype=np.int64)
    np.cumsum([x.size for x in reads], out=off[1:])
    sig = np.concatenate(reads)
    for scale in ("none", "zscale", "medmad"):
        want, kept_w = oracle_hits(sig, off, motif, scale)
        hits, kept = ctx.motifseq(sig, off, motif, scale=scale)
        ok = np.isfinite(want["dist"])
        assert_hits_equal(hits[ok, 0], want[ok], kept, kept_w, f"ties/{scale}")


def test_multiple_models(ctx):
    m1 = synth.make_motif()
    m2 = synth.make_motif(n_levels=20, dwell=8, seed=9)[:163]
    m3 = synth.make_motif(n_levels=3, dwell=3, seed=4)
    sig, off, _ = synth.motifseq_reads_np(64, 3000, m1)
    hits, kept = ctx.motifseq(sig, off, [m1, m2, m3], scale="medmad")
    for k, m in enumerate((m1, m2, m3)):
        want, kept_w = oracle_hits(sig, off, m, "medmad")
        assert_hits_equal(hits[:, k], want, kept, kept_w, f"model {k}")


def test_outlier_window_args(ctx):
    motif = synth.make_motif()
    sig, off, _ = synth.motifseq_reads_np(48, 2000, motif)
    for lo, hi in ((0, 1200), (300, 700), (-100, 32767), (450, 560)):
        want, kept_w = oracle_hits(sig, off, motif, "zscale", lo, hi)
        hits, kept = ctx.motifseq(sig, off, motif, scale="zscale", scale_low=lo, scale_hi=hi)
        ok = kept_w > 0
        assert_hits_equal(hits[ok, 0], want[ok], kept, kept_w, f"window ({lo},{hi})")


def test_trace_last_row_and_signal(ctx, golden_dir):
    """sqk_motifseq_trace: normalised signal == sklearn zscale bit for bit; cost[-1,:] == oracle's."""
    ex = np.load(os.path.join(golden_dir, "example_read.npz"), allow_pickle=True)
    raw = ex["raw"][:6000]
    kept = raw[(raw > 0) & (raw < 1200)]
    z, _, _ = oracle.zscale(kept)
    d, s, e, last = oracle.dtw_subsequence_rolling(ex["model"], z, want_last_row=True)
    hit, norm, row = ctx.motifseq_trace(raw, ex["model"], scale="zscale")
    assert np.array_equal(norm, z)
    assert np.array_equal(row, last)
    assert (int(hit["start"]), int(hit["end"]), float(hit["dist"])) == (s, e, d)
    assert int(np.argmin(row)) == e


def test_fp32_fast_mode(ctx):
    motif = synth.make_motif()
    sig, off, _ = synth.motifseq_reads_np(2048, 4096, motif)
    want, _ = oracle_hits(sig, off, motif, "zscale")
    hits, _ = ctx.motifseq(sig, off, motif, scale="zscale", precision="fp32")
    h = hits[:, 0]
    rel = np.abs(h["dist"] - want["dist"]) / np.maximum(want["dist"], 1e-9)
    assert rel.max() < 1e-4, rel.max()
    mism = float(np.mean((h["start"] != want["start"]) | (h["end"] != want["end"])))
    print(f"fp32 fast mode index mismatch rate: {mism:.4%}")
    assert mism < 0.02


def test_bench_size_properties(ctx):
    """BASELINE config 3 size (100k x 4096, N=80) in device mode: size-independent properties +
    an oracle check of a fixed sub-sample + shard invariance (what multi-GPU sharding relies on)."""
    import torch
    motif = synth.make_motif()
    R, M = 100_000, 4096
    sig = synth.motifseq_reads_torch(R, M, motif, "cuda")
    off = torch.arange(R + 1, dtype=torch.int64, device="cuda") * M
    hits_t, kept_t = ctx.motifseq(sig.view(-1), off, motif, scale="zscale", max_read_len=M)
    torch.cuda.synchronize()
    h = sqk.hits_from_torch(hits_t)[:, 0]
    kept = kept_t.cpu().numpy()
    assert (kept > 0).all() and (kept <= M).all()
    assert (h["start"] >= 0).all() and (h["start"] <= h["end"]).all() and (h["end"] < kept).all()
    assert np.isfinite(h["dist"]).all() and (h["dist"] >= 0).all()
    # idempotence
    hits2, _ = ctx.motifseq(sig.view(-1), off, motif, scale="zscale", max_read_len=M)
    torch.cuda.synchronize()
    assert torch.equal(hits_t, hits2)
    # shard invariance: second half alone == second half of the whole
    half = R // 2
    hits3, _ = ctx.motifseq(sig.view(-1), off[half:].contiguous(), motif, scale="zscale", max_read_len=M)
    torch.cuda.synchronize()
    assert torch.equal(hits_t[half:], hits3)
    # oracle on a fixed sub-sample
    idx = np.arange(0, R, R // 512)[:512]
    sub = sig[torch.from_numpy(idx).cuda()].cpu().numpy().reshape(-1)
    suboff = np.arange(idx.size + 1, dtype=np.int64) * M
    want, kept_w = oracle_hits(sub, suboff, motif, "zscale")
    assert_hits_equal(h[idx], want, kept[idx], kept_w, "bench-size subsample")
    # planted motifs are found: at least 45 % of reads have a hit with a small distance
    assert (h["dist"] < 25).mean() > 0.4


def test_very_long_reads_use_global_staging(ctx):
    """Reads longer than the shared-memory staging window of the stats kernel (~113k samples) take the global
    scratch path; the DTW itself streams any length."""
    motif = synth.make_motif()
    sig, off = synth.ragged_reads_np([300_000, 120_000, 5000, 250_001], motif)
    for scale in ("zscale", "medmad"):
        want, kept_w = oracle_hits(sig, off, motif, scale)
        hits, kept = ctx.motifseq(sig, off, motif, scale=scale)
        assert_hits_equal(hits[:, 0], want, kept, kept_w, f"long/{scale}")
    segs, nsegs = ctx.segmenter(sig, off, sqk.SegConfig(max_segs=64))
    want_s, want_n = oracle.segmenter_batch(sig, off, oracle.SegCfg(), 0, 900, 0, 64)
    assert np.array_equal(nsegs, want_n)
    for r in range(nsegs.size):
        assert np.array_equal(segs[r, :nsegs[r]], want_s[r, :want_n[r]])


def test_argument_errors_are_reported_not_crashed(ctx):
    motif = synth.make_motif()
    sig, off, _ = synth.motifseq_reads_np(4, 500, motif)
    hits, kept = ctx.motifseq(sig[:0], np.zeros(1, dtype=np.int64), motif)          # zero reads is fine
    assert hits.shape == (0, 1)
    with pytest.raises(sqk.SqkError) as ei:
        ctx.motifseq(sig, off[::-1].copy(), motif)                                   # offsets not monotone
    assert ei.value.code == -1
    with pytest.raises(sqk.SqkError) as ei:
        ctx.motifseq(sig, off, [motif] * 300)                                        # more models than one call takes
    assert ei.value.code == -4
    with pytest.raises(sqk.SqkError):
        ctx.segmenter(sig, off, sqk.SegConfig(corrector=-1))
    with pytest.raises(sqk.SqkError):
        ctx.segmenter(sig, off, sqk.SegConfig(max_segs=0))
    # the context is still usable afterwards
    want, _ = oracle_hits(sig, off, motif, "zscale")
    hits, _ = ctx.motifseq(sig, off, motif, scale="zscale")
    assert_hits_equal(hits[:, 0], want, what="after errors")


@pytest.mark.parametrize("chunk", [4096, 50_000, 1 << 20])
def test_host_pipeline_chunking(ctx, chunk):
    """Host mode splits the batch into chunks (two slots, copies overlapped with kernels, results through
    pinned staging).  Tiny chunks force many of them, with ragged reads straddling every kind of boundary,
    a read longer than a chunk, empty reads, pageable and pinned buffers: results must not depend on it."""
    motif = synth.make_motif()
    rng = np.random.default_rng(chunk)
    lengths = [int(v) for v in rng.integers(0, 9000, 120)] + [0, 0, 70_000, 1, 5000]
    sig, off = synth.ragged_reads_np(lengths, motif)
    want, kept_w = oracle_hits(sig, off, motif, "zscale")
    ok = kept_w > 0
    ctx.set_chunk_samples(chunk)
    try:
        hits, kept = ctx.motifseq(sig, off, [motif, motif[:33]], scale="zscale")
        p_sig = sqk.pinned_empty(sig.size, np.int16); p_sig[:] = sig
        p_hits = sqk.pinned_empty((len(lengths), 2), sqk.HIT_DTYPE)
        hits2, kept2 = ctx.motifseq(p_sig, off, [motif, motif[:33]], scale="zscale", out=p_hits)
        segs, nsegs = ctx.segmenter(sig, off, sqk.SegConfig(max_segs=64))
    finally:
        ctx.set_chunk_samples(0)
    assert_hits_equal(hits[ok, 0], want[ok], kept, kept_w, f"chunk={chunk}")
    assert np.array_equal(hits.view(np.uint8), np.asarray(hits2).view(np.uint8)) and np.array_equal(kept, kept2)
    want2, _ = oracle_hits(sig, off, motif[:33], "zscale")
    assert_hits_equal(hits[ok, 1], want2[ok], what=f"chunk={chunk} model 2")
    want_s, want_n = oracle.segmenter_batch(sig, off, oracle.SegCfg(), 0, 900, 0, 64)
    assert np.array_equal(nsegs, want_n)
    for r in range(nsegs.size):
        assert np.array_equal(segs[r, :nsegs[r]], want_s[r, :want_n[r]])
    sqk.pinned_free(p_sig); sqk.pinned_free(p_hits)


def test_understated_max_read_len_is_flagged_not_overrun(ctx):
    """Device mode trusts the caller's max_read_len for sizing the staging window; a longer read must come back
    flagged (-3 / n_segs -1), never overrun shared memory."""
    import torch
    motif = synth.make_motif()
    sig, off = synth.ragged_reads_np([1000, 200_000, 3000], motif)
    want, kept_w = oracle_hits(sig, off, motif, "zscale")
    hits_t, kept_t = ctx.motifseq(torch.from_numpy(sig).cuda(), torch.from_numpy(off).cuda(), motif, scale="zscale",
                                  max_read_len=3000)
    torch.cuda.synchronize()
    h = sqk.hits_from_torch(hits_t)[:, 0]
    assert int(h["start"][1]) == -3 and int(kept_t[1]) == -1
    assert_hits_equal(h[[0, 2]], want[[0, 2]], what="neighbours of the flagged read")
    segs, nsegs = ctx.segmenter(torch.from_numpy(sig).cuda(), torch.from_numpy(off).cuda(), sqk.SegConfig(), max_read_len=3000)
    torch.cuda.synchronize()
    assert int(nsegs[1]) == -1
    with pytest.raises(ValueError):
        sqk.segs_to_lists(segs.cpu().numpy(), nsegs.cpu().numpy())


# ---- float64 signals: the reference's `-s` path on a TSV of floats ---------------------------
def test_float_signal_golden(ctx, golden_dir):
    g = np.load(os.path.join(golden_dir, "float_signal_golden.npz"))
    motif = synth.make_motif()
    for scale in ("zscale", "medmad"):
        hits, kept = ctx.motifseq(g["signals"], g["offsets"], motif, scale=scale, scale_low=40, scale_hi=300)
        ok = g[scale + "_start"] != -9
        assert np.array_equal(kept, g[scale + "_kept"])
        assert np.array_equal(hits["start"][ok, 0], g[scale + "_start"][ok])
        assert np.array_equal(hits["end"][ok, 0], g[scale + "_end"][ok])
        assert np.array_equal(hits["dist"][ok, 0], g[scale + "_dist"][ok])
        assert (hits["start"][~ok, 0] < 0).all()
    want = json.load(open(os.path.join(golden_dir, "float_signal_segs.json")))
    segs, nsegs = ctx.segmenter(g["signals"], g["offsets"], sqk.SegConfig(lim_low=want["lim_low"], lim_hi=want["lim_hi"], max_segs=64))
    got = sqk.segs_to_lists(segs, nsegs)
    for r, w in enumerate(want["segs"]):
        assert got[r] == (w if w else False), r


@pytest.mark.parametrize("mode", ["host", "device"])
def test_float_signal_vs_oracle(ctx, mode):
    """Arbitrary doubles (not multiples of 0.01), ragged lengths, negative values, both scalings, both tools."""
    motif = synth.make_motif()
    rng = np.random.default_rng(12)
    lengths = [3000, 0, 1, 2, 9, 4096, 130, 700, 12000, 33]
    sig_i, off = synth.ragged_reads_np(lengths, motif)
    sig = (sig_i.astype(np.float64) - 500.0) * 0.1773 + rng.standard_normal(sig_i.size) * 1e-3
    for scale in ("zscale", "medmad", "none"):
        want, kept_w = oracle.motifseq_batch_f64(sig, off, motif, lo=-60, hi=90, scale=scale)
        if mode == "device":
            import torch
            h_t, k_t = ctx.motifseq(torch.from_numpy(sig).cuda(), torch.from_numpy(off).cuda(), motif, scale=scale, scale_low=-60, scale_hi=90)
            torch.cuda.synchronize()
            hits, kept = sqk.hits_from_torch(h_t), k_t.cpu().numpy()
        else:
            hits, kept = ctx.motifseq(sig, off, motif, scale=scale, scale_low=-60, scale_hi=90)
        ok = (kept_w > 0) & np.isfinite(want["dist"])
        assert np.array_equal(kept, kept_w)
        assert_hits_equal(hits[ok, 0], want[ok], what=f"f64 {scale}/{mode}")
        assert (hits["start"][~ok, 0] < 0).all()
    for params in (dict(), dict(error=80, corrector=0, window=10), dict(Num=1000)):
        cfg = sqk.SegConfig(lim_low=-60, lim_hi=90, max_segs=256, **params)
        ocfg = oracle.SegCfg(cfg.error, cfg.corrector, cfg.window, cfg.seg_dist, cfg.std_scale, cfg.stall_len)
        want_s, want_n = oracle.segmenter_batch_f64(sig, off, ocfg, -60, 90, cfg.Num, 256)
        if mode == "device":
            import torch
            s_t, n_t = ctx.segmenter(torch.from_numpy(sig).cuda(), torch.from_numpy(off).cuda(), cfg)
            torch.cuda.synchronize()
            segs, nsegs = s_t.cpu().numpy(), n_t.cpu().numpy()
        else:
            segs, nsegs = ctx.segmenter(sig, off, cfg)
        assert np.array_equal(nsegs, want_n), params
        for r in range(nsegs.size):
            assert np.array_equal(segs[r, :nsegs[r]], want_s[r, :want_n[r]]), (params, r)


def test_float_signal_that_is_integer_valued_matches_int16_path(ctx):
    """Feeding raw integers as float64 must give exactly what the int16 kernels give."""
    motif = synth.make_motif()
    sig, off, _ = synth.motifseq_reads_np(64, 3000, motif)
    for scale in ("zscale", "medmad"):
        a, ka = ctx.motifseq(sig, off, motif, scale=scale)
        b, kb = ctx.motifseq(sig.astype(np.float64), off, motif, scale=scale)
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8)) and np.array_equal(ka, kb)
    s1, n1 = ctx.segmenter(sig, off, sqk.SegConfig())
    s2, n2 = ctx.segmenter(sig.astype(np.float64), off, sqk.SegConfig())
    assert np.array_equal(n1, n2) and np.array_equal(s1, s2)
