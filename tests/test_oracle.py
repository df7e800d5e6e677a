"""CPU tests of the oracle itself: the checker is only worth something if it is pinned.
(i) against the committed golden fixtures (outputs of the reference's own code, see
tests/golden/make_golden.py); (ii) against numpy / sklearn run live; (iii) against an independent
numpy implementation and brute-force path enumeration; (iv) against the reference tree itself when
it is present (build container only)."""
import json
import os

import numpy as np
import pytest

import oracle
from oracle import numpy_ref, refload
from squigglekit_b200 import synth


def test_pairwise_sum_matches_numpy_and_golden(golden_dir):
    gold = json.load(open(os.path.join(golden_dir, "pairwise_sums.json")))
    rng = np.random.default_rng(gold["rng_seed"])
    for n_s, rec in gold["sums"].items():
        n = int(n_s)
        a = (rng.integers(1, 1200, n).astype(np.float64) - 511.37) ** 2
        assert float(np.sum(a)).hex() == rec["seed_vals_sum"], f"numpy changed its summation at n={n}"
        assert oracle.np_sum(a) == float(np.sum(a)), n
    for n in [1, 5, 8, 9, 127, 128, 129, 1000, 4095, 4096, 20000, 36977, 50000, 100003]:
        a = rng.standard_normal(n) ** 2 * 1e3
        assert oracle.np_sum(a) == float(np.sum(a)), n


def test_normalisation_matches_sklearn_and_numpy():
    motif = synth.make_motif()
    sig, off, _ = synth.motifseq_reads_np(24, 3000, motif)
    for r in range(24):
        s = sig[off[r]:off[r + 1]].astype(int)
        s = s[(s > 0) & (s < 1200)]
        z, mu, sd = oracle.zscale(s)
        assert np.array_equal(z, numpy_ref.zscale_np(s))
        m, med, mad = oracle.medmad(s)
        assert np.array_equal(m, numpy_ref.medmad_np(s))
        assert med == float(np.median(s)) and mad == float(np.median(np.abs(s - med)))
    # sigma == 0 -> sklearn divides by 1.0
    z, _, sd = oracle.zscale(np.full(50, 500))
    assert sd == 1.0 and not z.any()
    # sklearn's "mean not close to zero" corrections cannot trigger on int16-range data
    worst = np.full(50000, 32767)
    worst[0] = 32766
    assert np.array_equal(oracle.zscale(worst)[0], numpy_ref.zscale_np(worst))


def test_dtw_three_implementations_agree():
    rng = np.random.default_rng(2)
    motif = synth.make_motif()
    for trial in range(6):
        y = rng.standard_normal(rng.integers(1, 700))
        x = motif[: rng.integers(1, 81)]
        d, cost, (px, py) = oracle.dtw_subsequence(x, y)
        d2, s2, e2 = oracle.dtw_subsequence_rolling(x, y)
        d3, cost3, path3 = numpy_ref.dtw_rows(x, y)
        assert d == d2 == d3
        assert (py[0], py[-1]) == (s2, e2) == (path3[0][1], path3[-1][1])
        assert np.array_equal(cost, cost3)
        assert list(zip(px.tolist(), py.tolist())) == path3


def test_dtw_tie_order_on_integers():
    """Integer-valued inputs make exact ties common; all three implementations must break them alike."""
    rng = np.random.default_rng(7)
    for trial in range(200):
        x = rng.integers(-2, 3, rng.integers(1, 9)).astype(float)
        y = rng.integers(-2, 3, rng.integers(1, 40)).astype(float)
        d, cost, (px, py) = oracle.dtw_subsequence(x, y)
        d2, s2, e2 = oracle.dtw_subsequence_rolling(x, y)
        d3, _, path3 = numpy_ref.dtw_rows(x, y)
        assert (d, py[0], py[-1]) == (d2, s2, e2) == (d3, path3[0][1], path3[-1][1])
        assert e2 == int(np.argmin(cost[-1]))


def test_dtw_distance_is_the_true_minimum():
    rng = np.random.default_rng(3)
    for trial in range(150):
        x = rng.integers(-3, 4, rng.integers(1, 5)).astype(float)
        y = rng.integers(-3, 4, rng.integers(1, 7)).astype(float)
        assert oracle.dtw_subsequence(x, y)[0] == numpy_ref.brute_min_cost(x, y)


def test_motifseq_goldens(golden_dir):
    """The batch oracle reproduces the rows the reference's get_region_multi printed."""
    g = np.load(os.path.join(golden_dir, "motifseq_golden.npz"))
    for mname in ("motif80", "example163"):
        for scale in ("zscale", "medmad"):
            key = f"{mname}_{scale}"
            for full in (True, False):
                hits, kept = oracle.motifseq_batch(g["signals"], g["offsets"], g["model_" + mname], scale=scale, full_matrix=full)
                ok = (g[key + "_start"] != -2) & np.isfinite(g[key + "_dist"])
                assert np.array_equal(hits["start"][ok], g[key + "_start"][ok])
                assert np.array_equal(hits["end"][ok], g[key + "_end"][ok])
                assert np.array_equal(hits["dist"][ok], g[key + "_dist"][ok])
                assert np.array_equal(kept[ok], g[key + "_kept"][ok])


def test_example_read_goldens(golden_dir):
    ex = np.load(os.path.join(golden_dir, "example_read.npz"), allow_pickle=True)
    exp = json.load(open(os.path.join(golden_dir, "example_expected.json")))
    raw = ex["raw"]
    off = np.array([0, raw.size], dtype=np.int64)
    for scale in ("zscale", "medmad"):
        f = exp["tsv"][scale].split("\t")
        hits, kept = oracle.motifseq_batch(raw, off, ex["model"], scale=scale)
        assert (int(hits["start"][0]), int(hits["end"][0]), repr(float(hits["dist"][0]))) == (int(f[3]), int(f[4]), f[6])
    segs, n = oracle.segmenter_batch(raw, off, oracle.SegCfg(), 0, 900, 0, 16)
    assert segs[0, :n[0]].tolist() == exp["segs_raw"]


def test_get_segs_goldens(golden_dir):
    g = np.load(os.path.join(golden_dir, "segmenter_inputs.npz"))
    gold = json.load(open(os.path.join(golden_dir, "segmenter_golden.json")))
    for case in gold:
        p = case["params"]
        cfg = oracle.SegCfg(**{k: p[k] for k in ("error", "corrector", "window", "seg_dist", "std_scale", "stall_len") if k in p})
        segs, n = oracle.segmenter_batch(g["signals"], g["offsets"], cfg, p.get("lim_low", 0), p.get("lim_hi", 900),
                                         p.get("Num", 0), 512)
        for r, want in enumerate(case["reads"]):
            got = segs[r, :n[r]].tolist() if n[r] else False
            assert got == (want["segs"] if want is not None else False), (p, r)


@pytest.mark.skipif(not refload.available(), reason="reference tree not present (GPU box)")
def test_against_reference_tree_live():
    seg = refload.load("segmenter")
    ms = refload.load("MotifSeq")
    ssig, soff = synth.segmenter_reads_np(12, 2500, seed=99)
    for params in (dict(), dict(error=80, corrector=0, window=10), dict(error=1, corrector=2, window=25, std_scale=1.1)):
        a = refload.Args(**params)
        cfg = oracle.SegCfg(a.error, a.corrector, a.window, a.seg_dist, a.std_scale, a.stall_len)
        for r in range(12):
            s = seg.scale_outliers(ssig[soff[r]:soff[r + 1]].astype(int)[:-1], a)
            assert seg.get_segs(s, a) == oracle.get_segs(s, cfg, max_segs=512)
    model, order, L = ms.read_synth_model(os.path.join(refload.REFERENCE_ROOT, "example", "CATCTATCCAGGGTTAAATT.model"))
    assert (order, L, len(model[order[0]])) == (["3_prime_end"], [20], 163)


def test_get_segs_pa_goldens(golden_dir):
    """fast5 default path (pA conversion first): oracle vs the reference's own convert_to_pA_numpy + get_segs."""
    g = np.load(os.path.join(golden_dir, "segmenter_inputs.npz"))
    gold = json.load(open(os.path.join(golden_dir, "segmenter_pa_golden.json")))
    off = np.array(gold["offset"])
    scale = np.array([float("{0:.2f}".format(v)) for v in gold["range"]]) / gold["digitisation"]
    for case in gold["cases"]:
        p = case["params"]
        cfg = oracle.SegCfg(**{k: p[k] for k in ("error", "corrector", "window", "seg_dist", "std_scale", "stall_len") if k in p})
        segs, n = oracle.segmenter_batch_pa(g["signals"], g["offsets"], off, scale, cfg, p.get("lim_low", 0),
                                            p.get("lim_hi", 900), p.get("Num", 0), 512)
        for r, want in enumerate(case["reads"]):
            got = segs[r, :n[r]].tolist() if n[r] else False
            assert got == (want["segs"] if want is not None else False), (p, r)


def test_pa_conversion_matches_numpy_round():
    rng = np.random.default_rng(4)
    raw = rng.integers(-200, 2000, 5000).astype(np.int16)
    for offset, rg, dig in ((16.0, 1493.94, 8192.0), (-7.0, 1234.56, 8192.0), (3.0, 1467.61, 2048.0)):
        want = np.round((raw.astype(int) + offset) * (rg / dig), 2)
        assert np.array_equal(oracle.convert_to_pa(raw, offset, rg / dig), want)


def test_adapter_oracle_matches_reference_loop(golden_dir):
    """dRNA_segmenter.py's slow5-branch loop (run from the reference file by tests/golden/make_adapter_golden.py)
    vs the C restatement, on the committed vectors: the reference's example BLOW5 read + synthetic dRNA-like reads."""
    import json
    g = np.load(os.path.join(golden_dir, "adapter_inputs.npz"))
    want = json.load(open(os.path.join(golden_dir, "adapter_golden.json")))["segments"]
    segs, found = oracle.adapter_batch(g["signals"], g["offsets"])
    got = [[int(segs[r, 0]), int(segs[r, 1])] if found[r] else None for r in range(len(want))]
    assert got == want
    assert sum(w is not None for w in want) >= 30 and any(w is None for w in want)


def _rollmean_inputs():
    import importlib.util
    spec = importlib.util.spec_from_file_location("rollmean_inputs", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rollmean_inputs.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_rollmean_oracle_matches_reference_loop(golden_dir):
    """dRNA_segmenter.py's TSV-branch loop (run from the reference file, `w` injected, real pandas doing the rolling mean:
    tests/golden/make_rollmean_golden.py) vs the C restatement on the seeded inputs, both window lengths."""
    import json
    want = json.load(open(os.path.join(golden_dir, "rollmean_golden.json")))
    sig, off = _rollmean_inputs().concatenated()
    segs, found = oracle.rollmean_batch(sig, off)
    got = [[int(segs[r, 0]), int(segs[r, 1])] if found[r] else None for r in range(off.size - 1)]
    assert got == want["segments"]
    assert sum(w is not None for w in got) >= 30 and any(w is None for w in got)
    segs, found = oracle.rollmean_batch(sig, off[:13], oracle.RollmeanCfg(w=700))
    assert [[int(segs[r, 0]), int(segs[r, 1])] if found[r] else None for r in range(12)] == want["w700_first12"]


def test_rollmean_threshold_arithmetic_is_pandas():
    """bot = t.mean() - 0.5 * t.std() of t = rolling(w).mean(): the oracle's Kahan add/remove + nanops restatement against
    real pandas (when installed), bit for bit, including windows longer than the read and constant reads."""
    pd = pytest.importorskip("pandas")
    import warnings
    rng = np.random.default_rng(11)
    for trial in range(60):
        n = int(rng.integers(1, 12000)); w = int(rng.choice([1, 2, 7, 100, 2000, 2000, 5000]))
        sig = rng.integers(1, 1200, n)
        if trial % 7 == 0:
            sig[:] = sig[0]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            t = pd.Series(sig).rolling(window=w).mean()
            mn, sd = t.mean(), t.std()
            bot = mn - sd * 0.5
        _, thr = oracle.rollmean_seg(sig, oracle.RollmeanCfg(w=w), want_thresholds=True)
        for a, b in ((thr[0], bot), (thr[1], mn), (thr[2], sd)):
            assert a == b or (np.isnan(a) and np.isnan(b)), (trial, n, w, thr, bot, mn, sd)
