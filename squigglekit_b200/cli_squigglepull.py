"""Drop-in command line for the reference's SquigglePull.py -- the feeder that writes the signal TSV both hot paths read
(``-s``): one line per read, ``fast5 <TAB> readID [<TAB> digitisation <TAB> offset <TAB> range <TAB> sampling_rate] <TAB> s0 ...``
(SquigglePull.py:243-253), raw integers with ``-r`` or picoamperes ``np.round((raw + offset) * (range / digitisation), 2)``
with ``range`` first cut to two decimals (SquigglePull.py:186-193, 236-239).  Same flags.  Host-side only: the fast5 files are
read by squigglekit_b200/fast5.py (no h5py; deflate and VBZ chunks), raw rows are written by libsqk's batched text writer
(``sqk_tsv_format``), pA rows field by field with ``str()`` as the reference does.

Auto-detection of multi-read files keeps the reference's rule -- the SECOND key of the file's root group contains "read"
(SquigglePull.py:143-146) -- except that a file with a single root key is looked at by that key instead of raising.
"""
from __future__ import annotations

import argparse
import os
import sys
import time
import traceback

import numpy as np

BATCH_READS = 256


class MyParser(argparse.ArgumentParser):
    def error(self, message):
        sys.stderr.write('error: %s\n' % message)
        self.print_help()
        sys.exit(2)


def build_parser():
    parser = MyParser(description="SquigglePull - extraction and (optional) conversion to pA of raw signal from Oxford Nanopore fast5 files")
    parser.add_argument("-p", "--path", help="Top directory path of fast5 files")
    parser.add_argument("-t", "--type", action="store", default="auto", choices=["auto", "single", "multi"],
                        help="Specify the type of files provided. Default is autodetection which enables a mix of single and multifast5 files.")
    parser.add_argument("-v", "--verbose", action="store_true", help="Engage higher output verbosity")
    parser.add_argument("-r", "--raw_signal", action="store_true", help="No conversion to pA, raw signal is extracted instead")
    parser.add_argument("-i", "--extra_info", action="store_true",
                        help="Print extra information used for signal conversion and in methylation calling - nanopolish/f5c")
    return parser


def _txt(v):
    return v.decode() if isinstance(v, (bytes, bytearray)) else str(v)


def _record(raw, read_id, ch):
    """One read as the reference's f5_dic: raw int16 array + the channel constants (range through "{0:.2f}")."""
    return {"raw": raw, "readID": _txt(read_id), "digitisation": ch["digitisation"], "offset": ch["offset"],
            "range": float("{0:.2f}".format(ch["range"])), "sampling_rate": ch["sampling_rate"]}


def extract_f5_all(filename, args):
    """-> (list of records in file order, multi?) -- SquigglePull.py:130-234 on top of fast5.py."""
    from . import fast5
    f = fast5.Fast5File(filename)
    keys = f.keys()
    multi = False
    if args.type == "auto":
        probe = keys[1] if len(keys) > 1 else (keys[0] if keys else "")
        multi = "read" in probe
        if args.verbose:
            sys.stderr.write("{} detected as a {} fast5 file\n".format(filename, "multi" if multi else "single"))
    elif args.type == "multi":
        multi = True
    recs = []
    if not multi:
        try:
            reads = f["Raw/Reads"]
            rd = reads[reads.keys()[0]]
            recs.append(_record(np.ascontiguousarray(rd["Signal"].read(), dtype=np.int16), rd.attrs["read_id"],
                                f["UniqueGlobalKey/channel_id"].attrs))
        except Exception:
            traceback.print_exc()
            sys.stderr.write("extract_fast5_all():failed to extract raw signal or fastq from {}".format(filename))
            recs = []
        return recs, multi
    for read in keys:
        try:
            g = f[read]
            recs.append(_record(np.ascontiguousarray(g["Raw/Signal"].read(), dtype=np.int16), g["Raw"].attrs["read_id"],
                                g["channel_id"].attrs))
        except Exception:
            traceback.print_exc()
            sys.stderr.write("extract_fast5_all():failed to read readID: {}".format(read))
    return recs, multi


def head_of(rec, args, fast5):
    if args.extra_info:
        return "{}\t{}\t{}\t{}\t{}\t{}".format(fast5, rec["readID"], rec["digitisation"], rec["offset"], rec["range"],
                                               rec["sampling_rate"])
    return "{}\t{}".format(fast5, rec["readID"])


def convert_to_pA(rec):
    """np.round(convert_to_pA_numpy(raw, digitisation, range, offset), 2) -- SquigglePull.py:189-191, 236-239."""
    raw_unit = rec["range"] / rec["digitisation"]
    return np.round((np.array(rec["raw"], dtype=int) + rec["offset"]) * raw_unit, 2)


def emit(pending, args, out):
    """pending: [(fast5, record)].  Raw rows go through libsqk's writer in one call, pA rows field by field."""
    from . import tsv
    if not pending:
        return
    if args.raw_signal:
        offsets = np.zeros(len(pending) + 1, dtype=np.int64)
        np.cumsum([r["raw"].size for _, r in pending], out=offsets[1:])
        signals = np.concatenate([r["raw"] for _, r in pending]) if offsets[-1] else np.zeros(0, np.int16)
        text = tsv.format_reads([head_of(r, args, f5) for f5, r in pending], signals, offsets)
        out.flush()
        (out.buffer if hasattr(out, "buffer") else out).write(text if hasattr(out, "buffer") else text.decode())
    else:
        for f5, r in pending:
            out.write(head_of(r, args, f5) + "\t" + "\t".join(map(str, convert_to_pA(r))) + "\n")
    pending.clear()


def main(argv=None, out=None):
    out = sys.stdout if out is None else out
    parser = build_parser()
    raw_args = sys.argv[1:] if argv is None else list(argv)
    args = parser.parse_args(raw_args)
    if not raw_args:
        parser.print_help(sys.stderr)
        sys.exit(1)
    if args.verbose:
        sys.stderr.write("Verbose mode on. Starting timer.\n")
        start_time = time.time()
    if not args.path or not os.path.isdir(args.path):
        sys.stderr.write("The provided path {} is not an existing directory.\n".format(args.path))
        sys.exit(1)
    pending = []
    for dirpath, dirnames, files in os.walk(args.path):
        for fast5 in files:
            if fast5.endswith('.fast5'):
                fast5_file = os.path.join(dirpath, fast5)
                try:
                    recs, multi = extract_f5_all(fast5_file, args)
                except Exception:
                    traceback.print_exc()
                    recs = []
                if not recs:
                    sys.stderr.write("main():data not extracted from {}. Moving to next file.".format(fast5_file))
                    continue
                pending += [(fast5, r) for r in recs]
                if len(pending) >= BATCH_READS:
                    emit(pending, args, out)
    emit(pending, args, out)
    out.flush()
    if args.verbose:
        sys.stderr.write("Time taken: {}\n".format(time.time() - start_time))


if __name__ == '__main__':
    main()
