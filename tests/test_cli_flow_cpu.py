"""CPU tests of the command lines' -s (signal file) flow: the batched text reader with its helper thread, the split into
the vectorised path and the per-line path (pA lines, empty lines, all-zero reads), the rows written by libsqk -- with the
GPU context replaced by a stand-in that answers from the CPU oracle.  The printed rows must be the ones the per-read
expressions of the reference give (MotifSeq.py:252-298 + :441-449; segmenter.py:179-230).  What the GPU computes is the
business of the -m gpu tests; this file is about everything around it."""
import io

import numpy as np
import pytest

import oracle
import squigglekit_b200 as sqk
from squigglekit_b200 import cli_motifseq, cli_segmenter, synth, tsv


class OracleContext:
    """Context.motifseq / Context.segmenter answered by the oracle (test double)."""

    def __init__(self):
        self.calls = 0

    def motifseq(self, signals, offsets, models, scale="zscale", scale_low=0, scale_hi=1200, precision="fp64", want_kept=True,
                 **kw):
        self.calls += 1
        signals, offsets = np.asarray(signals), np.asarray(offsets)
        n = offsets.size - 1
        hits = np.zeros((n, len(models)), dtype=sqk.HIT_DTYPE)
        for c, model in enumerate(models):
            fn = oracle.motifseq_batch_f64 if signals.dtype.kind == "f" else oracle.motifseq_batch
            h, kept = fn(signals, offsets, model, lo=scale_low, hi=scale_hi, scale=scale, full_matrix=False)
            for f in ("start", "end", "dist"):
                hits[f][:, c] = h[f]
        return hits, (kept if want_kept else None)

    def segmenter(self, signals, offsets, cfg, **kw):
        self.calls += 1
        signals, offsets = np.asarray(signals), np.asarray(offsets)
        ocfg = oracle.SegCfg(cfg.error, cfg.corrector, cfg.window, cfg.seg_dist, cfg.std_scale, cfg.stall_len)
        fn = oracle.segmenter_batch_f64 if signals.dtype.kind == "f" else oracle.segmenter_batch
        return fn(signals, offsets, ocfg, lim_lo=cfg.lim_low, lim_hi=cfg.lim_hi, num=cfg.Num, max_segs=cfg.max_segs)


@pytest.fixture()
def pageable(monkeypatch):
    monkeypatch.setattr(tsv, "DEFAULT_PINNED", False)       # no CUDA runtime here: the reader's buffer is plain memory


def _signal_file(tmp_path, n_reads, n_samples, motif, head_cols, odd=True):
    sig, off, _ = synth.motifseq_reads_np(n_reads, n_samples, motif, seed=99)
    lines, reads = [], []
    for r in range(n_reads):
        s = sig[off[r]:off[r + 1]]
        head = [f"b{r // 50}.fast5", f"read_{r:05d}"] + ["x"] * (head_cols - 2)
        lines.append("\t".join(head + [str(int(v)) for v in s]))
        reads.append((head[0], head[1], s))
    if odd:
        # the lines that leave the vectorised path: pA-like floats, an all-zero read, a line without signal columns
        pa = np.round(sig[off[3]:off[4]].astype(np.float64) * 0.1717 + 2.5, 2)
        lines.insert(40, "\t".join(["pa.fast5", "read_pa"] + ["x"] * (head_cols - 2) + [repr(float(v)) for v in pa]))
        reads.insert(40, ("pa.fast5", "read_pa", pa))
        lines.insert(90, "\t".join(["zero.fast5", "read_zero"] + ["x"] * (head_cols - 2) + ["0"] * 50))
        reads.insert(90, None)
        lines.insert(120, "\t".join(["short.fast5", "read_short"]))
        reads.insert(120, None)
    path = tmp_path / "signal.tsv"
    path.write_text("\n".join(lines) + "\n")
    return str(path), reads


@pytest.mark.parametrize("scale", ["zscale", "medmad"])
def test_motifseq_signal_file_flow(tmp_path, pageable, monkeypatch, scale):
    motif = synth.make_motif()
    path, reads = _signal_file(tmp_path, 300, 1024, motif, head_cols=8)
    monkeypatch.setattr(cli_motifseq, "BATCH_READS", 64)     # several batches, some clean, some with odd lines
    args = cli_motifseq.build_parser().parse_args(["-s", path, "-m", "unused.model", "-l", scale])
    model, m_order, L = {"m80": motif, "half": motif[:40].copy()}, ["m80", "half"], [10, 5]
    out = io.StringIO()
    ctx = OracleContext()
    cli_motifseq.run_signal_file(ctx, args, model, m_order, L, out)
    assert ctx.calls >= 5
    want = []
    for rec in reads:
        if rec is None:
            continue
        f5, rid, s = rec
        for c, name in enumerate(m_order):
            if s.dtype.kind == "f":
                h, _ = oracle.motifseq_batch_f64(s, np.array([0, s.size]), model[name], scale=scale, full_matrix=False)
            else:
                h, _ = oracle.motifseq_batch(s, np.array([0, s.size]), model[name], scale=scale, full_matrix=False)
            want.append(cli_motifseq.format_row(f5, rid, name, int(h["start"][0]), int(h["end"][0]), h["dist"][0], args.slope,
                                                args.intercept, args.std_const, L[c]))
    assert out.getvalue().rstrip("\n").split("\n") == want


@pytest.mark.parametrize("flags", [[], ["-k", "-u"], ["-k", "-g", "-u", "-b", "800"]])
def test_segmenter_signal_file_flow(tmp_path, pageable, monkeypatch, capsys, flags):
    motif = synth.make_motif()
    path, reads = _signal_file(tmp_path, 260, 2048, motif, head_cols=4)
    monkeypatch.setattr(cli_segmenter, "BATCH_READS", 50)
    args = cli_segmenter.build_parser().parse_args(["-s", path] + flags)
    cfg = cli_segmenter.config_from_args(args) if hasattr(cli_segmenter, "config_from_args") else None
    if cfg is None:
        cfg = sqk.SegConfig(error=args.error, corrector=args.corrector, window=args.window, seg_dist=args.seg_dist,
                            std_scale=args.std_scale, stall_len=args.stall_len, lim_low=args.lim_low, lim_hi=args.lim_hi,
                            Num=args.Num, stall=args.stall, stall_start=args.stall_start, gap=args.gap, gap_dist=args.gap_dist)
    out = io.StringIO()
    cli_segmenter.run_signal_file(OracleContext(), args, cfg, out)
    ocfg = oracle.SegCfg(cfg.error, cfg.corrector, cfg.window, cfg.seg_dist, cfg.std_scale, cfg.stall_len)
    want = []
    for rec in reads:
        if rec is None:
            continue
        f5, _, s = rec
        x = s[:cfg.Num] if cfg.Num else s[:-1]               # segmenter.py:104-105: Num = 0 drops the last sample
        kept = x[(x > cfg.lim_low) & (x < cfg.lim_hi)]
        found = oracle.get_segs(kept, ocfg)
        if not found:
            continue
        if args.test:
            found = sqk.test_segs(found, cfg)
            if not found:
                continue
        want.append(f5 + "\t" + ",".join(str(v) for ij in found for v in ij))
    got = out.getvalue().rstrip("\n").split("\n") if out.getvalue() else []
    assert got == want


def test_drna_segmenter_signal_file_flow(tmp_path, pageable, golden_dir, monkeypatch):
    """dRNA_segmenter.py -s: the golden reads (segments produced by the reference's own loop body with real pandas) through
    the batched reader, with one line that is not plain int16 (it takes the per-line path) and small batches."""
    import importlib.util
    import json
    import os
    from squigglekit_b200 import cli_drna_segmenter
    spec = importlib.util.spec_from_file_location("rollmean_inputs", os.path.join(golden_dir, "rollmean_inputs.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    reads = mod.reads()[:12]
    want = json.load(open(os.path.join(golden_dir, "rollmean_golden.json")))["segments"][:12]
    path = tmp_path / "sig.tsv"
    with open(path, "w") as fh:
        for i, r in enumerate(reads):
            vals = [str(int(v)) for v in r]
            if i == 5 and not 0 < int(vals[7]) < 1200:
                vals[7] = "40000"                               # an outlier beyond int16: the line is flagged, parsed with int() and clipped
            elif i == 5:
                vals.insert(7, "40000")                         # (an extra outlier sample does not change the kept samples)
            fh.write("\t".join([f"f{i}.fast5", f"read{i}", "a", "b"] + vals) + "\n")

    class Ctx(OracleContext):
        def rollmean(self, signals, offsets, cfg, **kw):
            self.calls += 1
            ocfg = oracle.RollmeanCfg(cfg.w, cfg.seg_dist, cfg.lo_thresh, cfg.hi_thresh, cfg.shift, cfg.std_factor)
            return oracle.rollmean_batch(np.asarray(signals), np.asarray(offsets), ocfg, lim_lo=cfg.lim_low, lim_hi=cfg.lim_hi)

    monkeypatch.setattr(cli_drna_segmenter, "BATCH_READS", 4)
    args = cli_drna_segmenter.build_parser().parse_args(["-s", str(path)])
    out = io.StringIO()
    ctx = Ctx()
    cli_drna_segmenter.run_signal_file(ctx, args, sqk.RollmeanConfig(w=args.window), out)
    assert ctx.calls >= 3
    exp = [f"f{i}.fast5\tread{i}\t{w[0]}\t{w[1]}" for i, w in enumerate(want) if w is not None]
    assert [l for l in out.getvalue().split("\n") if l] == exp
