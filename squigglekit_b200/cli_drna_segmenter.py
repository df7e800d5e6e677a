"""Drop-in command line for the reference's dRNA_segmenter.py.

``-f/--slow5 file.blow5``: ``readID<TAB>start<TAB>end`` of the adapter segment (dRNA_segmenter.py:86-176: outlier
removal, threshold statistics over samples [1000, 5000), one-sided run detector).
``-s/--signal file.tsv``: ``fast5<TAB>readID<TAB>start<TAB>end`` from the rolling-mean detector (dRNA_segmenter.py:272-326:
``rolling(window=w).mean()``, ``bot = mean - 0.5 std``, runs below ``bot``).  As shipped that branch raises NameError
(``w`` only exists in the comment ``# w = 2000``, :81); here ``w`` defaults to 2000 and ``--window`` changes it.
Reads without a segment print nothing, as in the reference.  The whole per-read computation runs in libsqk on the GPU,
batched.  ``-p`` plotting is out of scope -> warning.
"""
from __future__ import annotations

import argparse
import sys

import numpy as np

BATCH_SAMPLES = 48 << 20
BATCH_READS = 16384


class MyParser(argparse.ArgumentParser):
    def error(self, message):
        sys.stderr.write('error: %s\n' % message)
        self.print_help()
        sys.exit(2)


def build_parser():
    parser = MyParser(description="dRNA_segmenter - cut out adapter region of dRNA signal")
    parser.add_argument("-s", "--signal", help="Signal file")
    parser.add_argument("-f", "--slow5", help="slow5 file")
    parser.add_argument("-c", "--start_col", type=int, default="4", help="start column for signal")
    parser.add_argument("-p", "--plot", action="store_true", help="Live plot each segment")
    parser.add_argument("--window", type=int, default=2000,
                        help="rolling-mean window of the -s branch (the reference's `# w = 2000`)")
    return parser


def flush(ctx, names, sigs, out):
    if not names:
        return
    offsets = np.zeros(len(sigs) + 1, dtype=np.int64)
    np.cumsum([s.size for s in sigs], out=offsets[1:])
    segs, found = ctx.adapter(np.concatenate(sigs) if offsets[-1] else np.zeros(0, np.int16), offsets)
    for r, name in enumerate(names):
        if found[r] > 0:
            out.write("{}\t{}\t{}\n".format(name, int(segs[r, 0]), int(segs[r, 1])))
    names.clear(); sigs.clear()


def flush_tsv(ctx, cfg, names, sigs, out):
    if not names:
        return
    offsets = np.zeros(len(sigs) + 1, dtype=np.int64)
    np.cumsum([s.size for s in sigs], out=offsets[1:])
    segs, found = ctx.rollmean(np.concatenate(sigs) if offsets[-1] else np.zeros(0, np.int16), offsets, cfg)
    for r, (f5, rid) in enumerate(names):
        if found[r] > 0:
            out.write("{}\t{}\t{}\t{}\n".format(f5, rid, int(segs[r, 0]), int(segs[r, 1])))
    names.clear(); sigs.clear()


def run_signal_file(ctx, args, cfg, out):
    """-s input through the batched text reader (squigglekit_b200.tsv): a batch of plain int16 lines goes from the parsed
    buffer straight into one sqk_rollmean call; a batch with anything else in it takes the per-line path, int() per field as
    the reference does (:278; values beyond int16 are outliers either way)."""
    from . import tsv
    with tsv.Reader(args.signal, args.start_col, max_lines=BATCH_READS, max_samples=BATCH_SAMPLES) as rd:
        for b in rd:
            if not b.status.any():
                segs, found = ctx.rollmean(b.signals[:int(b.offsets[b.n])], b.offsets, cfg)
                heads = b.heads(2)
                for r in np.flatnonzero(np.asarray(found) > 0):
                    h = heads[r]
                    out.write("{}\t{}\t{}\t{}\n".format(h[0], h[1] if len(h) > 1 else "", int(segs[r, 0]), int(segs[r, 1])))
                continue
            names, sigs = [], []
            for i in range(b.n):
                h = b.head(i)
                if b.status[i] & tsv.NO_SIGNAL:
                    sig = np.zeros(0, dtype=np.int16)            # (the reference would look at an empty list: no segment)
                elif b.status[i] & tsv.NOT_INT16:
                    sig = np.clip(np.array([int(v) for v in b.tail_text(i).split("\t") if v != ""], dtype=np.int64),
                                  -32768, 32767).astype(np.int16)
                else:
                    sig = b.sig(i).copy()
                names.append((h[0], h[1] if len(h) > 1 else "")); sigs.append(sig)
            flush_tsv(ctx, cfg, names, sigs, out)


def main(argv=None, out=sys.stdout):
    parser = build_parser()
    args = parser.parse_args(argv)
    if argv is None and len(sys.argv) == 1:
        parser.print_help(sys.stderr)
        sys.exit(1)
    if args.plot:
        sys.stderr.write("warning: -p plotting is not part of the B200 build; continuing without it\n")
    import squigglekit_b200 as sqk
    from . import slow5
    names, sigs, pending = [], [], 0
    if not args.slow5:
        if not args.signal:
            parser.print_help(sys.stderr)
            sys.exit(1)
        cfg = sqk.RollmeanConfig(w=args.window)
        with sqk.Context(0) as ctx:
            run_signal_file(ctx, args, cfg, out)
        return
    with sqk.Context(0) as ctx:
        for rec in slow5.read_blow5(args.slow5):
            names.append(rec["read_id"]); sigs.append(np.ascontiguousarray(rec["signal"], dtype=np.int16))
            pending += sigs[-1].size
            if pending >= BATCH_SAMPLES or len(names) >= BATCH_READS:
                flush(ctx, names, sigs, out); pending = 0
        flush(ctx, names, sigs, out)
