"""Drop-in command line for the reference's MotifSeq.py: same flags (MotifSeq.py:86-126), same stderr
banner (:147-151), same TSV header and rows (:160-163, :446-449), same `.model` formats.  The per-read
work (outlier removal, normalisation, subsequence DTW) runs in libsqk on the GPU, batched over reads;
the experimental score columns (:441-445) are the same scipy/numpy expressions on the host.

Deliberate differences, all loud:
  * ``-m`` accepts the scrappie `.model` the reference ships (its own ``read_bait_model`` crashes on it and
    never searches) as well as the bait format.
  * ``-i`` (scrappie simulation of a fasta) needs the scrappie neural network: not available -> error.
  * ``-v`` / ``--save`` plotting is out of scope -> warning, rows are still printed.
  * ``-s`` files with float (pA) columns run through the float64 front end (same results as the reference's
    float path); ``-x`` is only available for raw integer signal.
  * reads that are empty after outlier removal, or whose MAD is 0 under medmad, are reported on stderr
    and skipped (the reference raises / prints NaN rows for them).
"""
from __future__ import annotations

import argparse
import gzip
import os
import sys

import numpy as np

VERSION = "1.3.0"
BATCH_SAMPLES = 48 << 20
BATCH_READS = 16384


class MyParser(argparse.ArgumentParser):
    def error(self, message):
        sys.stderr.write('error: %s\n' % message)
        self.print_help()
        sys.exit(2)


def build_parser():
    parser = MyParser(description="MotifSeq - the Ctrl+f for signal. Signal-level local alignment of sequence motifs")
    group = parser.add_mutually_exclusive_group()
    mods = parser.add_mutually_exclusive_group()
    group.add_argument("-f", "--f5f", help="File list of fast5 paths")
    group.add_argument("-p", "--f5_path", help="Fast5 top dir")
    group.add_argument("-s", "--signal", help="Extracted signal file from SquigglePull")
    parser.add_argument("-l", "--scale", default="medmad", choices=["zscale", "medmad"],
                        help="scaling/normalisation factor to use")
    mods.add_argument("-i", "--fasta_input", help="fasta file to be converted to simulated signal by scrappy")
    parser.add_argument("--scrappie_model", default="squiggle_r94",
                        choices=['squiggle_r94', 'squiggle_r94_rna', 'squiggle_r10'],
                        help="model to use with fasta_input for conversion")
    mods.add_argument("-m", "--model",
                      help="custom multiline .tsv of signal to search for - see docs - name{tab}60{tab}435...")
    parser.add_argument("-x", "--sig_extract", action="store_true", help="Extract signal of match")
    parser.add_argument("--slope", type=float, default=2.90, help="[Experimental] slope")
    parser.add_argument("--intercept", type=float, default=-9.6, help="[Experimental] intercept")
    parser.add_argument("--std_const", type=float, default=0.08468, help="[Experimental] standard deviation constant")
    parser.add_argument("-v", "--view", action="store_true", help="view each output")
    parser.add_argument("--save", help="save path for images")
    parser.add_argument("--img", default="png", help="Type of image to save. png, jpeg, pdf, svg, etc. (default: png)")
    parser.add_argument("-scale_hi", "--scale_hi", type=int, default=1200, help="Upper limit for signal outlier scaling")
    parser.add_argument("-scale_low", "--scale_low", type=int, default=0, help="Lower limit for signal outlier scaling")
    parser.add_argument("-V", "--version", action="store_true", help="Print version information")
    parser.add_argument("--verbose", action="store_true", help="engage higher level of verbosity for troubleshooting")
    # additions (not in the reference)
    parser.add_argument("--device", type=int, default=0, help="[sqk] CUDA device index")
    parser.add_argument("--precision", default="fp64", choices=["fp64", "fp32"],
                        help="[sqk] fp64 = bit-exact with the reference's float64 DTW (default); fp32 = fast mode")
    group.add_argument("--slow5", help="[sqk] BLOW5 file (binary SLOW5) instead of fast5 / TSV input")
    parser.add_argument("--start_col", type=int, default=8, help="[sqk] first signal column of a -s file (reference: 8)")
    return parser


def format_row(fast5, read_id, name, start, end, dist, m, b, std, L, extract=None):
    """The row get_region_multi prints (MotifSeq.py:441-449), same expressions, same formatting."""
    from .tsv import ndtr                   # scipy.stats.norm.cdf bit for bit (scipy's ndtr restated in libsqk), without the import
    mod_mean = (m * L) + b
    mod_stdev = mod_mean * std
    Z = (dist - mod_mean) / mod_stdev
    p_value = ndtr(np.float64(Z))[()]       # numpy scalar, as norm.cdf returns
    hit_P = (1 - p_value) * 100
    cols = [fast5, read_id, name, start, end, end - start, dist, mod_mean, mod_stdev, Z, p_value, hit_P]
    if extract is not None:
        cols.append('\t'.join([str(i) for i in extract]))
    return "\t".join("{}".format(c) for c in cols)


def format_rows(heads, names, hits, m, b, std, L):
    """The same rows for a whole batch: the per-model constants once, Z / p-value / probability as arrays (the same
    IEEE operations and scipy's ndtr, restated in libsqk, as the scalar expressions), one join per row.  heads: [(fast5, readID)];
    hits: structured [n_reads, n_models].  Reads with a status hit (start < 0) are skipped: -> (rows, skipped indices)."""
    from .tsv import ndtr                   # scipy.special.ndtr bit for bit (what scipy.stats.norm.cdf evaluates), without the import
    rows, skipped = [], []
    n = len(heads)
    cols = []
    for c, name in enumerate(names):
        mod_mean = (m * L[c]) + b
        mod_stdev = mod_mean * std
        dist = hits["dist"][:, c].astype(np.float64)
        Z = (dist - mod_mean) / mod_stdev
        p_value = ndtr(Z)
        hit_P = (1 - p_value) * 100
        cols.append((name, hits["start"][:, c].tolist(), hits["end"][:, c].tolist(), dist.tolist(), "{}".format(mod_mean),
                     "{}".format(mod_stdev), Z.tolist(), p_value.tolist(), hit_P.tolist()))
    for r in range(n):
        f5, rid = heads[r]
        for name, start, end, dist, mm, ms, Z, pv, hp in cols:
            if start[r] < 0:
                skipped.append((r, start[r]))
                continue
            rows.append(f"{f5}\t{rid}\t{name}\t{start[r]}\t{end[r]}\t{end[r] - start[r]}\t{dist[r]}\t{mm}\t{ms}\t{Z[r]}\t{pv[r]}\t{hp[r]}")
    return rows, skipped


def format_rows_bytes(heads_bytes, names, hits, m, b, std, L):
    """format_rows for a batch whose head columns are still text (tsv.Batch.heads_bytes): the per-model constants and the
    score arrays are computed here exactly as above, the text is written by libsqk.  -> (bytes, skipped [(read, code)])."""
    from . import tsv
    consts, means, stdevs = [], [], []
    for c, name in enumerate(names):
        mod_mean = (m * L[c]) + b
        mod_stdev = mod_mean * std
        consts.append("{}\t{}".format(mod_mean, mod_stdev))
        means.append(mod_mean); stdevs.append(mod_stdev)
    zs, ps, hps = tsv.score_hits(hits, means, stdevs)      # (dist - mod_mean) / mod_stdev, norm.cdf, (1 - p) * 100
    text = tsv.format_hit_rows(heads_bytes, hits, names, consts, zs, ps, hps)
    bad = np.argwhere(hits["start"] < 0)
    return text, [(int(r), int(hits["start"][r, c])) for r, c in bad]


def _opener(path):
    return gzip.open if path.endswith('.gz') else open


def iter_reads(args):
    """Yield (fast5_name, read_id, int16 signal) in the reference's iteration order."""
    from . import fast5 as f5
    if args.f5f:
        with _opener(args.f5f)(args.f5f, 'rt') as s:
            for l in s:
                path = l.strip('\n').split('\t')[0]
                if not path:
                    continue
                yield from _one_fast5(f5, path, path.split('/')[-1], "Failed to extract signal: {} {}\n".format(path, path.split('/')[-1]))
    elif args.f5_path:
        for dirpath, dirnames, files in os.walk(args.f5_path):
            for fast5 in files:
                if fast5.endswith('.fast5'):
                    fast5_file = os.path.join(dirpath, fast5)
                    yield from _one_fast5(f5, fast5_file, fast5,
                                          "main():data not extracted. Moving to next file - {}\n".format(fast5_file))
    elif args.slow5:
        from . import slow5
        name = os.path.basename(args.slow5)
        for rec in slow5.read_blow5(args.slow5):
            yield name, rec["read_id"], rec["signal"]
    elif args.signal:
        with _opener(args.signal)(args.signal, 'rt') as s:
            for line in s:
                line = line.rstrip('\n')
                head, tail = split_signal_columns(line, args.start_col)
                if tail is None:
                    sys.stderr.write("No Signal found - please check signal format\n")
                    continue
                vals = np.fromstring(tail, dtype=np.float64, sep='\t')     # C strtod loop: ~20x the reference's float(i) list
                if not vals.any():
                    sys.stderr.write("No Signal found - please check signal format\n")
                    continue
                ints = np.rint(vals)
                if np.array_equal(ints, vals) and ints.min() >= -32768 and ints.max() <= 32767:
                    sig = ints.astype(np.int16)          # raw DAC values: the int16 kernels (2 bytes/sample)
                else:
                    sig = vals                           # pA or any other float signal: the float64 front end
                yield head[0], head[1] if len(head) > 1 else "", sig


def split_signal_columns(line, start_col):
    """'a<TAB>b<TAB>...<TAB>s0<TAB>s1...' -> (['a', 'b', ...first start_col fields], 's0<TAB>s1...') without splitting the
    (long) signal part; (fields, None) if the line has no signal columns."""
    pos = -1
    for _ in range(start_col):
        pos = line.find('\t', pos + 1)
        if pos < 0:
            return line.split('\t'), None
    tail = line[pos + 1:]
    return line[:pos].split('\t'), (tail if tail else None)


def _one_fast5(f5, path, name, fail_msg):
    try:
        f = f5.Fast5File(path)
        if f5.is_multi_fast5(f):
            for rname, rec in f5.read_multi_fast5(path).items():
                yield name, rec["read_id"], rec["signal"]
        else:
            rec = f5.read_single_fast5(path)
            # h5py hands the reference a bytes object for fixed-length string attributes and it prints b'...'
            # (MotifSeq.py:346, TODO at :46); reproduce that so downstream parsers see the same column
            rid = "b'{}'".format(rec["read_id"]) if rec.get("read_id_is_bytes") else rec["read_id"]
            yield name, rid, rec["signal"]
    except Exception as e:          # the reference prints a traceback and moves on (MotifSeq.py:334-351)
        sys.stderr.write("process_fast5():failed to extract events or fastq from: {} ({}: {})\n".format(path, type(e).__name__, e))
        sys.stderr.write(fail_msg)


def flush(ctx, args, batch, model, m_order, L, out):
    if not batch:
        return
    sigs = [b[2] for b in batch]
    offsets = np.zeros(len(sigs) + 1, dtype=np.int64)
    np.cumsum([s.size for s in sigs], out=offsets[1:])
    if any(x.dtype.kind == "f" for x in sigs):
        sigs = [x.astype(np.float64) for x in sigs]      # a batch with float reads goes through the float64 path as a whole
    signals = np.concatenate(sigs) if sigs else np.zeros(0, dtype=np.int16)
    hits, kept = ctx.motifseq(signals, offsets, [model[n] for n in m_order], scale=args.scale,
                              scale_low=args.scale_low, scale_hi=args.scale_hi, precision=args.precision)
    for r, (fast5, read_id, sig) in enumerate(batch):
        for c, name in enumerate(m_order):
            h = hits[r, c]
            start, end, dist = int(h["start"]), int(h["end"]), h["dist"]
            if start < 0:
                why = "no samples left after outlier removal" if start == -1 else "MAD is 0: med-MAD scaling undefined"
                sys.stderr.write("{} {}: {} - skipped\n".format(fast5, read_id, why))
                continue
            extract = None
            if args.sig_extract and sig.dtype.kind == "f":
                sys.stderr.write("{} {}: -x is not available for float (pA) signal input - row printed without the signal\n".format(fast5, read_id))
            elif args.sig_extract:
                _, norm, _ = ctx.motifseq_trace(sig, model[name], scale=args.scale, scale_low=args.scale_low,
                                                scale_hi=args.scale_hi, precision=args.precision)
                extract = norm[start:end]
            out.write(format_row(fast5, read_id, name, start, end, dist, args.slope, args.intercept, args.std_const,
                                 L[c], extract) + "\n")
    batch.clear()


def run_signal_file(ctx, args, model, m_order, L, out):
    """-s input through the batched text reader (squigglekit_b200.tsv): a batch of plain int16 lines goes from the parsed
    pinned buffer straight into one libsqk call and comes back as vectorised rows; a batch with anything else in it
    (pA values, empty lines, all-zero reads) or -x takes the per-line path with the reference's messages.
    SQK_CLI_PROFILE=1 prints where the wall-clock time went (stderr)."""
    import time
    from . import tsv
    prof = os.environ.get("SQK_CLI_PROFILE")
    tm = {"parse": 0.0, "heads": 0.0, "gpu": 0.0, "format": 0.0, "write": 0.0}
    models = [model[n] for n in m_order]
    t_last = time.perf_counter()
    with tsv.Reader(args.signal, args.start_col, max_lines=BATCH_READS, max_samples=BATCH_SAMPLES) as rd:
        for b in rd:
            t0 = time.perf_counter(); tm["parse"] += t0 - t_last
            if not b.status.any() and not args.sig_extract:
                heads = b.heads_bytes(2)
                t1 = time.perf_counter(); tm["heads"] += t1 - t0
                hits, _ = ctx.motifseq(b.signals[:int(b.offsets[b.n])], b.offsets, models, scale=args.scale,
                                       scale_low=args.scale_low, scale_hi=args.scale_hi, precision=args.precision, want_kept=False)
                t2 = time.perf_counter(); tm["gpu"] += t2 - t1
                text, skipped = format_rows_bytes(heads, m_order, hits, args.slope, args.intercept, args.std_const, L)
                t3 = time.perf_counter(); tm["format"] += t3 - t2
                if skipped:
                    hl = heads.decode("utf-8", "replace").split("\n")
                    for r, code in skipped:
                        why = "no samples left after outlier removal" if code == -1 else "MAD is 0: med-MAD scaling undefined"
                        sys.stderr.write("{}: {} - skipped\n".format(hl[r].replace("\t", " "), why))
                if text:
                    out.flush()
                    (out.buffer if hasattr(out, "buffer") else out).write(text if hasattr(out, "buffer") else text.decode("utf-8", "replace"))
                t_last = time.perf_counter(); tm["write"] += t_last - t3
                continue
            batch = []
            for i in range(b.n):
                h = b.head(i)
                if b.status[i] & tsv.NO_SIGNAL:
                    sys.stderr.write("No Signal found - please check signal format\n")
                    continue
                if b.status[i] == 0:
                    batch.append((h[0], h[1] if len(h) > 1 else "", b.sig(i).copy()))
                    continue
                vals = np.fromstring(b.tail_text(i), dtype=np.float64, sep='\t')
                if not vals.any():
                    sys.stderr.write("No Signal found - please check signal format\n")
                    continue
                ints = np.rint(vals)
                if np.array_equal(ints, vals) and ints.min() >= -32768 and ints.max() <= 32767:
                    sig = ints.astype(np.int16)
                else:
                    sig = vals
                batch.append((h[0], h[1] if len(h) > 1 else "", sig))
            flush(ctx, args, batch, model, m_order, L, out)
            t_last = time.perf_counter()
    if prof:
        sys.stderr.write("sqk profile (s): " + ", ".join("{} {:.3f}".format(k, v) for k, v in tm.items()) + "\n")


def main(argv=None):
    parser = build_parser()
    raw_args = sys.argv[1:] if argv is None else list(argv)
    args = parser.parse_args(raw_args)
    if not raw_args:
        parser.print_help(sys.stderr)
        sys.exit(1)
    if args.version:
        sys.stderr.write("SquiggleKit MotifSeq: {}\n".format(VERSION))
        sys.exit(1)
    if args.verbose:
        sys.stderr.write("Verbose mode active - dumping info to stderr\n")
        sys.stderr.write("SquiggleKit MotifSeq: {}\n".format(VERSION))
        sys.stderr.write("args: {}\n".format(args))

    sys.stderr.write("\n\n**********************************************************\n")
    sys.stderr.write("*  z-score, p-value, probability, etc. are based on      *\n")
    sys.stderr.write("*     preliminary experimental modeling only             *\n")
    sys.stderr.write("*                Use at own risk                         *\n")
    sys.stderr.write("**********************************************************\n\n\n")

    if args.fasta_input:
        sys.stderr.write("-i/--fasta_input needs the scrappie neural-network simulator, which is not available here; "
                         "run `scrappie squiggle` yourself and pass its output with -m\n")
        sys.exit(1)
    if not args.model:
        sys.stderr.write("no model given (-m)\n")
        parser.print_help(sys.stderr)
        sys.exit(1)
    if args.view or args.save:
        sys.stderr.write("warning: -v/--view and --save plotting are not part of the GPU port; printing rows only\n")
    if not (args.f5f or args.f5_path or args.signal or args.slow5):
        sys.stderr.write("Unknown file or path input")
        parser.print_help(sys.stderr)
        sys.exit(1)

    from . import Context, read_model
    model, m_order, L = read_model(args.model)
    if not m_order:
        sys.stderr.write("no motif found in {}\n".format(args.model))
        sys.exit(1)

    out = sys.stdout
    head = ["fast5", "readID", "model", "start", "end", "length", "distance_score", "model_mean", "model_stdev", "Z-score",
            "p-value", "hit_Probability"]
    if args.sig_extract:
        head.append("normalised_signal")
    out.write("\t".join(head) + "\n")

    with Context(args.device) as ctx:
        batch, n_samples = [], 0
        if args.signal:
            run_signal_file(ctx, args, model, m_order, L, out)
        for rec in ([] if args.signal else iter_reads(args)):
            batch.append(rec)
            n_samples += rec[2].size
            if n_samples >= BATCH_SAMPLES or len(batch) >= BATCH_READS:
                flush(ctx, args, batch, model, m_order, L, out)
                n_samples = 0
        flush(ctx, args, batch, model, m_order, L, out)
    out.flush()


if __name__ == '__main__':
    main()
