"""Drop-in command line for the slow5 branch of the reference's dRNA_segmenter.py: ``-f/--slow5 file.blow5`` in,
``readID<TAB>start<TAB>end`` of the adapter segment out (dRNA_segmenter.py:86-176); reads without a segment print
nothing, as in the reference.  The whole per-read computation (outlier removal, threshold statistics over samples
[1000, 5000), the run detector) runs in libsqk on the GPU, batched.

Differences, all loud:
  * ``-s/--signal`` (the TSV branch, dRNA_segmenter.py:272-326) cannot run in the reference as shipped (``w`` is
    undefined, NameError) and is not provided -> error message.
  * ``-p`` plotting is out of scope -> warning.
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


def main(argv=None, out=sys.stdout):
    parser = build_parser()
    args = parser.parse_args(argv)
    if argv is None and len(sys.argv) == 1:
        parser.print_help(sys.stderr)
        sys.exit(1)
    if args.plot:
        sys.stderr.write("warning: -p plotting is not part of the B200 build; continuing without it\n")
    if not args.slow5:
        sys.stderr.write("error: only -f/--slow5 input is supported: the reference's -s branch uses an undefined "
                         "window `w` (dRNA_segmenter.py:281) and cannot run as shipped\n")
        sys.exit(2)
    import squigglekit_b200 as sqk
    from . import slow5
    names, sigs, pending = [], [], 0
    with sqk.Context(0) as ctx:
        for rec in slow5.read_blow5(args.slow5):
            names.append(rec["read_id"]); sigs.append(np.ascontiguousarray(rec["signal"], dtype=np.int16))
            pending += sigs[-1].size
            if pending >= BATCH_SAMPLES or len(names) >= BATCH_READS:
                flush(ctx, names, sigs, out); pending = 0
        flush(ctx, names, sigs, out)
