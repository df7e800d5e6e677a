"""GPU parity tests for the segmenter hot path: libsqk vs golden outputs of the reference's own
get_segs / test_segs (tests/golden/segmenter_golden.json) and vs the CPU oracle.  Integer work:
bit-exact."""
import json
import os

import numpy as np
import pytest

import oracle
import squigglekit_b200 as sqk
from squigglekit_b200 import synth

pytestmark = pytest.mark.gpu


def cfg_from(params: dict) -> sqk.SegConfig:
    c = sqk.SegConfig(max_segs=512)
    for k, v in params.items():
        if k == "test":
            continue
        setattr(c, k, v)
    return c


def test_golden_reference_outputs(ctx, golden_dir):
    g = np.load(os.path.join(golden_dir, "segmenter_inputs.npz"))
    gold = json.load(open(os.path.join(golden_dir, "segmenter_golden.json")))
    for case in gold:
        cfg = cfg_from(case["params"])
        segs, nsegs = ctx.segmenter(g["signals"], g["offsets"], cfg)
        got = sqk.segs_to_lists(segs, nsegs)
        for r, want in enumerate(case["reads"]):
            if want is None:            # empty after filtering: the reference raises; we report no segments
                assert got[r] is False
                continue
            assert got[r] == want["segs"], (case["params"], r)
            if want["tested"] is not None:
                assert sqk.test_segs(got[r], cfg) == want["tested"], (case["params"], r)


def test_example_read(ctx, golden_dir):
    ex = np.load(os.path.join(golden_dir, "example_read.npz"), allow_pickle=True)
    want = json.load(open(os.path.join(golden_dir, "example_expected.json")))["segs_raw"]
    raw = ex["raw"]
    segs, nsegs = ctx.segmenter(raw, np.array([0, raw.size], dtype=np.int64), sqk.SegConfig())
    assert sqk.segs_to_lists(segs, nsegs)[0] == want


PARAM_SETS = [
    dict(),
    dict(error=80, corrector=0, window=10),
    dict(error=2, corrector=3, window=40, seg_dist=200, std_scale=0.5),
    dict(error=0, window=30, std_scale=1.5),
    dict(Num=1500),
    dict(Num=-100),
    dict(lim_hi=600, lim_low=420, window=60),
    dict(stall_len=0.0),
    dict(window=1, error=1, corrector=1, seg_dist=0),
]


@pytest.mark.parametrize("params", PARAM_SETS)
@pytest.mark.parametrize("mode", ["host", "device"])
def test_synthetic_vs_oracle(ctx, params, mode):
    sig, off = synth.segmenter_reads_np(256, 4096)
    extra, eoff = synth.ragged_reads_np([0, 1, 2, 9, 149, 150, 151, 700, 10000, 3])
    sig = np.concatenate([sig, extra])
    off = np.concatenate([off, off[-1] + eoff[1:]])
    cfg = cfg_from(params)
    ocfg = oracle.SegCfg(cfg.error, cfg.corrector, cfg.window, cfg.seg_dist, cfg.std_scale, cfg.stall_len)
    want, want_n = oracle.segmenter_batch(sig, off, ocfg, cfg.lim_low, cfg.lim_hi, cfg.Num, cfg.max_segs)
    if mode == "device":
        import torch
        s_t, n_t = ctx.segmenter(torch.from_numpy(sig).cuda(), torch.from_numpy(off).cuda(), cfg)
        torch.cuda.synchronize()
        segs, nsegs = s_t.cpu().numpy(), n_t.cpu().numpy()
    else:
        segs, nsegs = ctx.segmenter(sig, off, cfg)
    assert np.array_equal(nsegs, want_n), np.nonzero(nsegs != want_n)[0][:10]
    assert (want_n <= cfg.max_segs).all()
    for r in range(nsegs.size):
        assert np.array_equal(segs[r, :nsegs[r]], want[r, :want_n[r]]), r


def test_overflow_is_reported(ctx):
    sig, off = synth.segmenter_reads_np(8, 4096)
    cfg = sqk.SegConfig(window=5, error=0, seg_dist=0, std_scale=2.0, max_segs=2)
    segs, nsegs = ctx.segmenter(sig, off, cfg)
    assert (nsegs > 2).any()
    with pytest.raises(OverflowError):
        sqk.segs_to_lists(segs, nsegs)


def test_config2_size(ctx):
    """BASELINE config 2: 10k synthetic 4k-sample reads, -ku.  Full-size run, oracle check of all reads."""
    import torch
    R, M = 10_000, 4096
    sig = synth.segmenter_reads_torch(R, M, "cuda")
    off = torch.arange(R + 1, dtype=torch.int64, device="cuda") * M
    cfg = sqk.SegConfig(stall=True)
    s_t, n_t = ctx.segmenter(sig.view(-1), off, cfg, max_read_len=M)
    torch.cuda.synchronize()
    segs, nsegs = s_t.cpu().numpy(), n_t.cpu().numpy()
    ocfg = oracle.SegCfg()
    want, want_n = oracle.segmenter_batch(sig.cpu().numpy().reshape(-1), off.cpu().numpy(), ocfg, 0, 900, 0, cfg.max_segs)
    assert np.array_equal(nsegs, want_n)
    mask = np.arange(cfg.max_segs)[None, :, None] < want_n[:, None, None]
    assert np.array_equal(np.where(mask, segs, 0), np.where(mask, want, 0))
    # the planted stall is found near the start in most reads
    first = segs[:, 0, 0][want_n > 0]
    assert (first <= cfg.stall_start).mean() > 0.8


def _pa_cal(golden_dir):
    gold = json.load(open(os.path.join(golden_dir, "segmenter_pa_golden.json")))
    off = np.array(gold["offset"])
    scale = np.array([float("{0:.2f}".format(v)) for v in gold["range"]]) / gold["digitisation"]
    return gold, off, scale


def test_pa_mode_golden_reference_outputs(ctx, golden_dir):
    """a9: convert_to_pA_numpy + np.round(.., 2) before get_segs -- the reference's fast5 default."""
    g = np.load(os.path.join(golden_dir, "segmenter_inputs.npz"))
    gold, off, scale = _pa_cal(golden_dir)
    for case in gold["cases"]:
        cfg = cfg_from(case["params"])
        segs, nsegs = ctx.segmenter(g["signals"], g["offsets"], cfg, pa_offset=off, pa_scale=scale)
        got = sqk.segs_to_lists(segs, nsegs)
        for r, want in enumerate(case["reads"]):
            if want is None:
                assert got[r] is False
                continue
            assert got[r] == want["segs"], (case["params"], r)
            if want["tested"] is not None:
                assert sqk.test_segs(got[r], cfg) == want["tested"], (case["params"], r)


def test_pa_mode_example_read(ctx, golden_dir):
    ex = np.load(os.path.join(golden_dir, "example_read.npz"), allow_pickle=True)
    want = json.load(open(os.path.join(golden_dir, "example_expected.json")))["segs_pA"]
    raw = ex["raw"]
    segs, nsegs = ctx.segmenter(raw, np.array([0, raw.size], dtype=np.int64), sqk.SegConfig(),
                                pa_offset=[16.0], pa_scale=[float("{0:.2f}".format(1493.94)) / 8192.0])
    assert sqk.segs_to_lists(segs, nsegs)[0] == want


@pytest.mark.parametrize("mode", ["host", "device"])
def test_pa_mode_synthetic_vs_oracle(ctx, mode):
    sig, off = synth.segmenter_reads_np(300, 4096, seed=21)
    extra, eoff = synth.ragged_reads_np([0, 1, 2, 9, 150, 151, 700, 9000])
    sig = np.concatenate([sig, extra])
    off = np.concatenate([off, off[-1] + eoff[1:]])
    n = off.size - 1
    rng = np.random.default_rng(8)
    po = rng.integers(-30, 40, n).astype(float)
    ps = np.round(rng.uniform(1100, 1600, n), 2) / 8192.0
    for params in (dict(), dict(lim_hi=120, lim_low=40, window=60), dict(error=80, corrector=0, window=10), dict(Num=2000)):
        cfg = cfg_from(params)
        ocfg = oracle.SegCfg(cfg.error, cfg.corrector, cfg.window, cfg.seg_dist, cfg.std_scale, cfg.stall_len)
        want, want_n = oracle.segmenter_batch_pa(sig, off, po, ps, ocfg, cfg.lim_low, cfg.lim_hi, cfg.Num, cfg.max_segs)
        if mode == "device":
            import torch
            s_t, n_t = ctx.segmenter(torch.from_numpy(sig).cuda(), torch.from_numpy(off).cuda(), cfg, pa_offset=po, pa_scale=ps)
            torch.cuda.synchronize()
            segs, nsegs = s_t.cpu().numpy(), n_t.cpu().numpy()
        else:
            segs, nsegs = ctx.segmenter(sig, off, cfg, pa_offset=po, pa_scale=ps)
        assert np.array_equal(nsegs, want_n), (params, np.nonzero(nsegs != want_n)[0][:10])
        for r in range(n):
            assert np.array_equal(segs[r, :nsegs[r]], want[r, :want_n[r]]), (params, r)
