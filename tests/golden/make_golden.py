"""Regenerate the golden fixtures in this directory.  Runs ONLY in the build container
(needs /root/reference); the fixtures it writes are committed and are what travels.

What is "reference output" here and what is not:
  * segmenter goldens  -- produced by the reference's own get_segs / test_segs /
    scale_outliers (segmenter.py:311-318,399-494), imported unmodified via oracle/refload.py.
  * normalisation      -- produced by the real sklearn.preprocessing.scale and the literal
    numpy med-MAD expression of MotifSeq.py:192-200.
  * MotifSeq TSV rows  -- produced by the reference's own get_region_multi
    (MotifSeq.py:431-456: start/end extraction, mod_mean, Z, norm.cdf, row formatting) and
    read_synth_model (MotifSeq.py:354-379), with mlpy.dtw_subsequence replaced by the
    oracle's restatement (mlpy is not installable here) -> DTW numbers are "parity
    unpinned", everything around them is reference code.

usage:  python tests/golden/make_golden.py
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from oracle import numpy_ref, refload  # noqa: E402
from squigglekit_b200 import synth  # noqa: E402

REF = refload.REFERENCE_ROOT


def example_read():
    """Raw/Reads/Read_58517/Signal of example/test.fast5: one deflate chunk, 36 978 int16 samples
    (superblock v0, chunk of 47 670 B at file offset 12 808, no shuffle filter)."""
    blob = open(os.path.join(REF, "example", "test.fast5"), "rb").read()
    raw = np.frombuffer(zlib.decompress(blob[12808:12808 + 47670]), dtype="<i2")
    assert raw.size == 36978
    return raw


def run_region_multi(ms, args, sig, model, m_order, L, fast5, read_id):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        ms.get_region_multi(args, sig, model, m_order, fast5, read_id, args.slope, args.intercept,
                            args.std_const, L)
    return buf.getvalue()


def normalise(ms, raw, args):
    sig = np.array(raw, dtype=int)
    sig = ms.scale_outliers(sig, args)
    if args.scale == "zscale":
        return numpy_ref.zscale_np(sig)
    return numpy_ref.medmad_np(sig)


def main():
    ms = refload.load("MotifSeq")
    seg = refload.load("segmenter")

    # ---- example read + example model (BASELINE config 1) ---------------------------------
    raw = example_read()
    model_path = os.path.join(REF, "example", "CATCTATCCAGGGTTAAATT.model")
    model, m_order, L = ms.read_synth_model(model_path)
    name = m_order[0]
    ex = {"raw": raw, "model": np.array(model[name], dtype=np.float64), "L": np.array(L), "name": name,
          "read_id": "db4ae416-40c2-45c2-9cc9-7d49c5711a7c"}
    tsv = {}
    for scale in ("zscale", "medmad"):
        args = refload.Args(scale=scale)
        sig = normalise(ms, raw, args)
        tsv[scale] = run_region_multi(ms, args, sig, model, m_order, L, "test.fast5", ex["read_id"])
        args_x = refload.Args(scale=scale, sig_extract=True)
        tsv[scale + "_x"] = run_region_multi(ms, args_x, sig, model, m_order, L, "test.fast5", ex["read_id"])
    sargs = refload.Args()
    s = np.array(raw, dtype=int)[:-1]
    ex_segs_raw = seg.get_segs(seg.scale_outliers(s, sargs), sargs)
    pa = np.round(seg.convert_to_pA_numpy(np.array(raw, dtype=int), 8192.0, float("{0:.2f}".format(1493.94)), 16.0), 2)
    ex_segs_pa = seg.get_segs(seg.scale_outliers(np.array(pa[:-1], dtype=float), sargs), sargs)
    np.savez_compressed(os.path.join(HERE, "example_read.npz"), **ex)
    with open(os.path.join(HERE, "example_expected.json"), "w") as f:
        json.dump({"tsv": tsv, "segs_raw": ex_segs_raw, "segs_pA": ex_segs_pa,
                   "model_text": open(model_path).read()}, f, indent=1)

    # ---- synthetic MotifSeq set: ragged lengths, two models, both scalings -----------------
    motif80 = synth.make_motif()
    lengths = [1500, 81, 80, 79, 2000, 1, 2, 7, 8, 9, 127, 128, 129, 1033, 640, 3000, 257, 512, 2047, 999,
               1200, 1800, 333, 4096]
    signals, offsets = synth.ragged_reads_np(lengths, motif80)
    # tie-heavy read: few distinct levels, no noise
    rng = np.random.default_rng(5)
    ties = np.repeat(rng.integers(400, 620, 60), 12).astype(np.int16)
    const = np.full(300, 500, dtype=np.int16)           # sigma == 0 -> sklearn scale 1.0
    signals = np.concatenate([signals, ties, const])
    offsets = np.concatenate([offsets, [offsets[-1] + ties.size, offsets[-1] + ties.size + const.size]])
    models = {"motif80": motif80, "example163": ex["model"]}
    Ls = {"motif80": 10, "example163": 20}
    out = {"signals": signals, "offsets": offsets.astype(np.int64)}
    rows = {}
    for mname, mvec in models.items():
        out["model_" + mname] = mvec
        for scale in ("zscale", "medmad"):
            args = refload.Args(scale=scale)
            starts, ends, dists, kept, text = [], [], [], [], []
            for r in range(offsets.size - 1):
                rd = signals[offsets[r]:offsets[r + 1]]
                if scale == "medmad" and r == offsets.size - 2:
                    # constant read: mad == 0 -> inf/nan signal; documented degenerate case, not pinned
                    starts.append(-2); ends.append(-2); dists.append(np.nan); kept.append(rd.size); continue
                sig = normalise(ms, rd, args)
                kept.append(sig.size)
                t = run_region_multi(ms, args, sig, {mname: mvec}, [mname], [Ls[mname]], f"read{r}.fast5", f"id{r}")
                f = t.rstrip("\n").split("\t")
                starts.append(int(f[3])); ends.append(int(f[4])); dists.append(float(f[6]))
                text.append(t)
            key = f"{mname}_{scale}"
            out[key + "_start"] = np.array(starts, dtype=np.int32)
            out[key + "_end"] = np.array(ends, dtype=np.int32)
            out[key + "_dist"] = np.array(dists, dtype=np.float64)
            out[key + "_kept"] = np.array(kept, dtype=np.int32)
            rows[key] = "".join(text)
    np.savez_compressed(os.path.join(HERE, "motifseq_golden.npz"), **out)
    with open(os.path.join(HERE, "motifseq_rows.json"), "w") as f:
        json.dump(rows, f, indent=1)

    # ---- segmenter: real get_segs / test_segs on a synthetic set, several parameter sets ---
    ssig, soff = synth.segmenter_reads_np(24, 3000)
    extra, eoff = synth.ragged_reads_np([1, 2, 10, 149, 150, 151, 400, 5000])
    ssig = np.concatenate([ssig, extra])
    soff = np.concatenate([soff, soff[-1] + eoff[1:]])
    param_sets = [
        dict(),                                                     # defaults
        dict(stall=True, test=True),                                # -ku
        dict(stall=True, gap=True, test=True, gap_dist=800),        # -kgu
        dict(error=80, corrector=0, window=10),                     # live err-=1 branch
        dict(error=2, corrector=3, window=40, seg_dist=200, std_scale=0.5),
        dict(Num=1000),
        dict(lim_hi=600, lim_low=420, window=60),
    ]
    seg_out = []
    for ps in param_sets:
        a = refload.Args(**ps)
        per_read = []
        for r in range(soff.size - 1):
            sig = np.array(ssig[soff[r]:soff[r + 1]], dtype=int)[:a.Num]
            sig = seg.scale_outliers(sig, a)
            if sig.size == 0:
                per_read.append(None)            # reference raises on sig.min() of an empty array
                continue
            segs = seg.get_segs(sig, a)
            tested = None
            if segs and a.test:
                with contextlib.redirect_stderr(io.StringIO()):
                    tested = seg.test_segs([list(s) for s in segs], a)
            per_read.append({"segs": segs if segs else False,
                             "tested": (tested if tested else False) if a.test else None})
        seg_out.append({"params": ps, "reads": per_read})
    np.savez_compressed(os.path.join(HERE, "segmenter_inputs.npz"), signals=ssig, offsets=soff.astype(np.int64))
    with open(os.path.join(HERE, "segmenter_golden.json"), "w") as f:
        json.dump(seg_out, f)

    # ---- segmenter, fast5 default path: pA conversion (convert_to_pA_numpy + np.round) before get_segs ----
    n_reads = soff.size - 1
    rng = np.random.default_rng(17)
    cal_offset = rng.integers(-30, 40, n_reads).astype(float)
    cal_range = np.round(rng.uniform(1100.0, 1600.0, n_reads), 2)
    digitisation = 8192.0
    pa_sets = [dict(), dict(stall=True, test=True), dict(lim_hi=120, lim_low=40, window=60),
               dict(error=80, corrector=0, window=10), dict(Num=1200, std_scale=0.5)]
    pa_out = []
    for ps in pa_sets:
        a = refload.Args(**ps)
        per_read = []
        for r in range(n_reads):
            raw_r = np.array(ssig[soff[r]:soff[r + 1]], dtype=int)
            rg = float("{0:.2f}".format(cal_range[r]))                       # segmenter.py:344
            pa = np.round(seg.convert_to_pA_numpy(raw_r, digitisation, rg, cal_offset[r]), 2)
            sig = np.array(pa[:a.Num], dtype=float)
            sig = seg.scale_outliers(sig, a)
            if sig.size == 0:
                per_read.append(None)
                continue
            segs = seg.get_segs(sig, a)
            tested = None
            if segs and a.test:
                with contextlib.redirect_stderr(io.StringIO()):
                    tested = seg.test_segs([list(s) for s in segs], a)
            per_read.append({"segs": segs if segs else False,
                             "tested": (tested if tested else False) if a.test else None})
        pa_out.append({"params": ps, "reads": per_read})
    with open(os.path.join(HERE, "segmenter_pa_golden.json"), "w") as f:
        json.dump({"offset": cal_offset.tolist(), "range": cal_range.tolist(), "digitisation": digitisation,
                   "cases": pa_out}, f)

    # ---- float (pA) signals, the reference's `-s` path: MotifSeq.py:270-298 and segmenter.py:198-211 ----------
    fsig_i, foff = synth.ragged_reads_np([1800, 900, 2500, 1, 0, 77, 1300, 640], motif80)
    fsig = np.round((fsig_i.astype(int) + 9.0) * (1456.78 / 8192.0), 2)          # what SquigglePull prints without -r
    fout = {"signals": fsig, "offsets": foff.astype(np.int64)}
    for scale in ("zscale", "medmad"):
        args = refload.Args(scale=scale, scale_hi=300, scale_low=40)
        st, en, di, kp = [], [], [], []
        for r in range(foff.size - 1):
            sig = np.array([float(i) for i in fsig[foff[r]:foff[r + 1]]])       # MotifSeq.py:270
            sig = ms.scale_outliers(sig, args)
            kp.append(sig.size)
            if sig.size == 0 or (scale == "medmad" and np.median(np.abs(sig - np.median(sig))) == 0):
                st.append(-9); en.append(-9); di.append(np.nan); continue
            y = numpy_ref.zscale_np(sig) if scale == "zscale" else numpy_ref.medmad_np(sig)
            t = run_region_multi(ms, args, y, {"m": motif80}, ["m"], [10], "f", "id")
            f = t.rstrip("\n").split("\t")
            st.append(int(f[3])); en.append(int(f[4])); di.append(float(f[6]))
        fout[scale + "_start"] = np.array(st, dtype=np.int32); fout[scale + "_end"] = np.array(en, dtype=np.int32)
        fout[scale + "_dist"] = np.array(di); fout[scale + "_kept"] = np.array(kp, dtype=np.int32)
    sargs_f = refload.Args(lim_hi=160, lim_low=30)
    fsegs = []
    for r in range(foff.size - 1):
        sig = np.array(fsig[foff[r]:foff[r + 1]], dtype=float)[:sargs_f.Num]
        sig = seg.scale_outliers(sig, sargs_f)
        fsegs.append(seg.get_segs(sig, sargs_f) if sig.size else None)
    np.savez_compressed(os.path.join(HERE, "float_signal_golden.npz"), **fout)
    with open(os.path.join(HERE, "float_signal_segs.json"), "w") as f:
        json.dump({"lim_hi": 160, "lim_low": 30, "segs": fsegs}, f)

    # ---- numpy pairwise-sum known answers (pins oracle.np_sum and the CUDA sigma tree) -----
    rng = np.random.default_rng(11)
    sums = {}
    for n in [1, 7, 8, 9, 127, 128, 129, 255, 256, 1000, 4095, 4096, 4097, 20000, 36977]:
        a = (rng.integers(1, 1200, n).astype(np.float64) - 511.37) ** 2
        sums[str(n)] = {"seed_vals_sum": float(np.sum(a)).hex()}
    with open(os.path.join(HERE, "pairwise_sums.json"), "w") as f:
        json.dump({"rng_seed": 11, "sums": sums}, f, indent=1)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
