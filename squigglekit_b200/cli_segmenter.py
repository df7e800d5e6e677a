"""Drop-in command line for the reference's segmenter.py: same flags (segmenter.py:53-97), same
``name<TAB>s1,e1,s2,e2,...`` output (:138-143), same stderr messages.  Per read, the truncation
``sig[:Num]`` (with the reference's Num=0 -> -1 quirk that drops the last sample, :104-105), the outlier
removal (:311-318) and get_segs (:399-470) run in libsqk on the GPU, batched; test_segs (:473-494) is the
same two comparisons on the host.

fast5 input follows the reference: by default each read is converted to pA rounded to two decimals before
segmentation (:344-349, :366-370) -- done on the GPU from the raw samples and the per-read calibration --
and ``--raw_signal`` segments the raw integers.

``-s`` files may hold raw integers (int16 kernels) or floats such as SquigglePull's pA output (float64 front
end, same results as the reference's float branch).

Deliberate differences, all loud:
  * ``-v`` plotting is out of scope -> warning.
"""
from __future__ import annotations

import argparse
import gzip
import os
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
    parser = MyParser(description="segmenter - script to find obvious regions in squiggle data")
    group = parser.add_mutually_exclusive_group()
    group.add_argument("-i", "--ind", nargs='+', help="Individual fast5 file/s")
    group.add_argument("-p", "--f5_path", help="Fast5 top dir")
    group.add_argument("-s", "--signal", help="Extracted signal file from squigglePull")
    group.add_argument("--slow5", help="[sqk] BLOW5 file (binary SLOW5) instead of fast5 / TSV input")
    parser.add_argument("--single", action="store_true", help="single fast5 files")
    parser.add_argument("-n", "--Num", type=int, default=0, help="Section of signal to look at - default 0=all")
    parser.add_argument("-e", "--error", type=int, default=5, help="Allowable error in segment algorithm")
    parser.add_argument("-c", "--corrector", type=int, default=50,
                        help="Window size for increasing total error correction - better long segment detection")
    parser.add_argument("-w", "--window", type=int, default=150, help="Minimum segment window size to be detected")
    parser.add_argument("-d", "--seg_dist", type=int, default=50, help="Maximum distance between 2 segments to be merged into 1")
    parser.add_argument("-t", "--std_scale", type=float, default=0.75, help="Scale factor of STDev about median")
    parser.add_argument("-v", "--view", action="store_true", help="view each output")
    parser.add_argument("-g", "--gap", action="store_true", help="Turn on gap distance for stall to polyTAil")
    parser.add_argument("-b", "--gap_dist", type=int, default=3000,
                        help="Maximum distance between stall and polyTAil segment - for 10X/dRNA")
    parser.add_argument("-k", "--stall", action="store_true", help="Turn on stall detection - must be present")
    parser.add_argument("-u", "--test", action="store_true", help="Run Tests")
    parser.add_argument("-l", "--stall_len", type=float, default=0.25,
                        help="Minimum percentage of minimum window segment for initial stall segment")
    parser.add_argument("-j", "--stall_start", type=int, default=300,
                        help="Maximum distance for start of stall segment to be detected")
    parser.add_argument("-lim_hi", "--lim_hi", type=int, default=900, help="Upper limit for signal outlier scaling")
    parser.add_argument("-lim_low", "--lim_low", type=int, default=0, help="Lower limit for signal outlier scaling")
    parser.add_argument("--raw_signal", action="store_true", help="Plot raw signal instead of converting to pA")
    # additions (not in the reference)
    parser.add_argument("--device", type=int, default=0, help="[sqk] CUDA device index")
    parser.add_argument("--max_segs", type=int, default=64, help="[sqk] capacity of the per-read segment list")
    parser.add_argument("--start_col", type=int, default=4, help="[sqk] first signal column of a -s file (reference: 4)")
    return parser


def _opener(path):
    return gzip.open if path.endswith('.gz') else open


def _fast5_reads(args, path, label):
    """Yield (name_to_print, int16 signal, pa_offset, pa_scale) the way the reference's branches do: multi-read
    files print the read group name (segmenter.py:143), single-read files print the file name (:174 / :288).
    pa_scale = float("{0:.2f}".format(range)) / digitisation (:344, :389, :515-517)."""
    from . import fast5 as f5

    def cal(rec):
        return float(rec["offset"]), float("{0:.2f}".format(rec["range"])) / float(rec["digitisation"])
    try:
        if not args.single:
            for rname, rec in f5.read_multi_fast5(path).items():
                yield (rname, rec["signal"]) + cal(rec)
        else:
            rec = f5.read_single_fast5(path)
            yield (label, rec["signal"]) + cal(rec)
    except Exception as e:
        sys.stderr.write('process_fast5():failed to extract events or fastq from: {} ({}: {})'.format(path, type(e).__name__, e))
        sys.stderr.write("main():data not extracted. Moving to next file: {}".format(label))


def iter_reads(args):
    if args.f5_path:
        for dirpath, dirnames, files in os.walk(args.f5_path):
            for fast5 in files:
                if fast5.endswith('.fast5'):
                    yield from _fast5_reads(args, os.path.join(dirpath, fast5), fast5)
    elif args.ind:
        for fast5_file in args.ind:
            yield from _fast5_reads(args, fast5_file, fast5_file)
    elif args.slow5:
        from . import slow5
        for rec in slow5.read_blow5(args.slow5):
            yield (rec["read_id"], rec["signal"], float(rec["offset"]),
                   float("{0:.2f}".format(rec["range"])) / float(rec["digitisation"]))
    elif args.signal:
        with _opener(args.signal)(args.signal, 'rt') as s:
            from .cli_motifseq import split_signal_columns
            for line in s:
                line = line.rstrip('\n')
                head, tail = split_signal_columns(line, args.start_col)
                fast5 = head[0]
                if tail is None:
                    sys.stderr.write("No signal found in file: {} {}".format(args.signal, fast5))
                    continue
                first = tail.split('\t', 1)[0]
                if "." in first:                      # the reference switches to float parsing on this test (:198)
                    sig = np.fromstring(tail, dtype=np.float64, sep='\t')
                    if not sig.any():
                        sys.stderr.write("No signal found in file: {} {}".format(args.signal, fast5))
                        continue
                    yield fast5, sig, 0.0, 1.0        # float64 front end
                    continue
                sig = np.fromstring(tail, dtype=np.int64, sep='\t')
                if not sig.any():
                    sys.stderr.write("No signal found in file: {} {}".format(args.signal, fast5))
                    continue
                if sig.min() < -32768 or sig.max() > 32767:
                    yield fast5, sig.astype(np.float64), 0.0, 1.0
                    continue
                yield fast5, sig.astype(np.int16), 0.0, 1.0


def segment_batch(ctx, signals, offsets, cfg, **kw):
    """ctx.segmenter with room for every segment: the reference has no cap on the segments of a read, so when a read
    overflows the row (n_segs > max_segs) the batch is run again with rows as long as its longest list."""
    import dataclasses
    segs, nsegs = ctx.segmenter(signals, offsets, cfg, **kw)
    most = int(nsegs.max()) if nsegs.size else 0
    if most > cfg.max_segs:
        segs, nsegs = ctx.segmenter(signals, offsets, dataclasses.replace(cfg, max_segs=most), **kw)
    return segs, nsegs


def flush(ctx, args, cfg, batch, out):
    from . import segs_to_lists, test_segs
    if not batch:
        return
    sigs = [b[1] for b in batch]
    offsets = np.zeros(len(sigs) + 1, dtype=np.int64)
    np.cumsum([s.size for s in sigs], out=offsets[1:])
    if any(x.dtype.kind == "f" for x in sigs):
        sigs = [x.astype(np.float64) for x in sigs]
    if args.signal or args.raw_signal:
        segs, nsegs = segment_batch(ctx, np.concatenate(sigs), offsets, cfg)
    else:
        segs, nsegs = segment_batch(ctx, np.concatenate(sigs), offsets, cfg, pa_offset=[b[2] for b in batch],
                                    pa_scale=[b[3] for b in batch])
    for (name, _, _, _), found in zip(batch, segs_to_lists(segs, nsegs)):
        if not found:
            sys.stderr.write("no segments found: {}".format(name))
            continue
        if args.test:
            if args.stall and found[0][0] > args.stall_start:
                sys.stderr.write("start seg too late!")
            elif args.gap and len(found) > 1 and found[1][0] > found[0][1] + args.gap_dist:
                sys.stderr.write("second seg too far!")
            found = test_segs(found, cfg)
            if not found:
                continue
        flat = []
        for i, j in found:
            flat.append(str(i))
            flat.append(str(j))
        out.write("\t".join([name, ",".join(flat)]) + "\n")
    batch.clear()


def emit(args, cfg, names, segs, nsegs, out):
    """The printing half of flush() for a batch whose segments are already computed."""
    from . import segs_to_lists, test_segs
    lines = []
    for name, found in zip(names, segs_to_lists(segs, nsegs)):
        if not found:
            sys.stderr.write("no segments found: {}".format(name))
            continue
        if args.test:
            if args.stall and found[0][0] > args.stall_start:
                sys.stderr.write("start seg too late!")
            elif args.gap and len(found) > 1 and found[1][0] > found[0][1] + args.gap_dist:
                sys.stderr.write("second seg too far!")
            found = test_segs(found, cfg)
            if not found:
                continue
        lines.append(name + "\t" + ",".join(str(v) for ij in found for v in ij))
    if lines:
        out.write("\n".join(lines) + "\n")


def emit_batch(args, cfg, heads_bytes, segs, nsegs, out):
    """emit() for a batch whose names are still text (tsv.Batch.heads_bytes(1)): the -k / -g tests of test_segs
    (segmenter.py:473-494) on arrays, the rows written by libsqk; the reference's stderr notes for the reads it drops."""
    from . import tsv
    n = nsegs.shape[0]
    has = nsegs > 0
    first_start = segs[:, 0, 0]
    late = has & (first_start > args.stall_start) if cfg.stall else np.zeros(n, dtype=bool)
    far = np.zeros(n, dtype=bool)
    if cfg.gap and segs.shape[1] > 1:
        far = (nsegs > 1) & (segs[:, 1, 0] > segs[:, 0, 1] + args.gap_dist)
    keep = has.copy()
    if args.test:
        keep &= ~late & ~far
    if not has.all() or (args.test and (late.any() or far.any())):
        names = heads_bytes.decode("utf-8", "replace").split("\n")
        for r in np.flatnonzero(~has):
            sys.stderr.write("no segments found: {}".format(names[r]))
        if args.test:
            for r in np.flatnonzero(has & (late | far)):
                sys.stderr.write("start seg too late!" if (cfg.stall and late[r]) else "second seg too far!")
    text = tsv.format_seg_rows(heads_bytes, segs, nsegs, keep)
    if text:
        out.flush()
        if hasattr(out, "buffer"):
            out.buffer.write(text)
        else:
            out.write(text.decode("utf-8", "replace"))


def run_signal_file(ctx, args, cfg, out):
    """-s input through the batched text reader (squigglekit_b200.tsv): batches of plain int16 lines go from the parsed
    pinned buffer straight into one libsqk call; anything else in a batch takes the per-line path.
    SQK_CLI_PROFILE=1 prints where the wall-clock time went (stderr)."""
    import time
    from . import tsv
    tm = {"parse": 0.0, "heads": 0.0, "gpu": 0.0, "format": 0.0}
    t_last = time.perf_counter()
    with tsv.Reader(args.signal, args.start_col, max_lines=BATCH_READS, max_samples=BATCH_SAMPLES) as rd:
        for b in rd:
            t0 = time.perf_counter(); tm["parse"] += t0 - t_last
            if not b.status.any():
                heads = b.heads_bytes(1)
                t1 = time.perf_counter(); tm["heads"] += t1 - t0
                segs, nsegs = segment_batch(ctx, b.signals[:int(b.offsets[b.n])], b.offsets, cfg)
                t2 = time.perf_counter(); tm["gpu"] += t2 - t1
                emit_batch(args, cfg, heads, segs, nsegs, out)
                t_last = time.perf_counter(); tm["format"] += t_last - t2
                continue
            batch = []
            for i in range(b.n):
                fast5 = b.head(i)[0]
                if b.status[i] & tsv.NO_SIGNAL:
                    sys.stderr.write("No signal found in file: {} {}".format(args.signal, fast5))
                    continue
                if b.status[i] == 0:
                    batch.append((fast5, b.sig(i).copy(), 0.0, 1.0))
                    continue
                tail = b.tail_text(i)
                first = tail.split('\t', 1)[0]
                if "." in first:                      # the reference switches to float parsing on this test (:198)
                    sig = np.fromstring(tail, dtype=np.float64, sep='\t')
                else:
                    sig = np.fromstring(tail, dtype=np.int64, sep='\t')
                if not sig.any():
                    sys.stderr.write("No signal found in file: {} {}".format(args.signal, fast5))
                    continue
                if sig.dtype.kind == "i" and sig.min() >= -32768 and sig.max() <= 32767:
                    batch.append((fast5, sig.astype(np.int16), 0.0, 1.0))
                else:
                    batch.append((fast5, sig.astype(np.float64), 0.0, 1.0))
            flush(ctx, args, cfg, batch, out)
            t_last = time.perf_counter()
    if os.environ.get("SQK_CLI_PROFILE"):
        sys.stderr.write("\nsqk profile (s): " + ", ".join("{} {:.3f}".format(k, v) for k, v in tm.items()) + "\n")


def main(argv=None):
    parser = build_parser()
    raw_args = sys.argv[1:] if argv is None else list(argv)
    args = parser.parse_args(raw_args)
    if not raw_args:
        parser.print_help(sys.stderr)
        sys.exit(1)
    if not (args.f5_path or args.ind or args.signal or args.slow5):
        sys.stderr.write("Unknown file or path input")
        parser.print_help(sys.stderr)
        sys.exit(1)
    if args.view:
        sys.stderr.write("warning: -v/--view plotting is not part of the GPU port; printing segments only\n")

    from . import Context, SegConfig
    cfg = SegConfig(error=args.error, corrector=args.corrector, window=args.window, seg_dist=args.seg_dist,
                    std_scale=args.std_scale, stall_len=args.stall_len, lim_low=args.lim_low, lim_hi=args.lim_hi,
                    Num=args.Num, stall=args.stall, stall_start=args.stall_start, gap=args.gap, gap_dist=args.gap_dist,
                    max_segs=args.max_segs)
    out = sys.stdout
    with Context(args.device) as ctx:
        batch, n_samples = [], 0
        if args.signal:
            run_signal_file(ctx, args, cfg, out)
        for rec in ([] if args.signal else iter_reads(args)):
            batch.append(rec)
            n_samples += rec[1].size
            if n_samples >= BATCH_SAMPLES or len(batch) >= BATCH_READS:
                flush(ctx, args, cfg, batch, out)
                n_samples = 0
        flush(ctx, args, cfg, batch, out)
    out.flush()
    sys.stderr.write("Done")


if __name__ == '__main__':
    main()
