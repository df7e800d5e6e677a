"""Golden vectors for the rolling-mean adapter finder (dRNA_segmenter.py, TSV branch, :272-326).

The reference keeps the algorithm inline in main(); as shipped it raises NameError because `w` only exists in a comment
(:81 `# w = 2000`).  This script (build container only: it reads /root/reference at run time, nothing of it is copied
into the repo) cuts the body of the `for read in s:` loop out of the reference file, wraps it UNMODIFIED into a function
that takes `w` as an argument, feeds it the reads of tests/golden/rollmean_inputs.py as SquigglePull-style TSV lines and
captures what it prints.  real pandas (3.0.2) does the rolling mean / mean / std.  Output: rollmean_golden.json (per
read: [x, y] or null), committed.

usage:  python tests/golden/make_rollmean_golden.py
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys
import textwrap
import types
import warnings

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import rollmean_inputs  # noqa: E402
from oracle import refload  # noqa: E402

REF = refload.REFERENCE_ROOT


def reference_loop():
    """-> callable(signal, w) running the reference's loop body on one TSV line, returning what it printed."""
    src = open(os.path.join(REF, "dRNA_segmenter.py")).read().split("\n")
    first = next(i for i, l in enumerate(src) if l.strip() == "for read in s:")
    last = next(i for i, l in enumerate(src) if i > first and l.startswith("def scale_outliers"))
    body = textwrap.dedent("\n".join(src[first + 1:last]))
    scale_src = src[last:]
    scale_src = "\n".join(scale_src[:next(i for i, l in enumerate(scale_src) if l.startswith("if __name__"))])
    code = "def _one(read, args, w):\n" + textwrap.indent(body, "    ") + "\n" + scale_src
    ns = {"np": np, "pd": pd}
    exec(compile(code, "dRNA_segmenter_tsv_loop", "exec"), ns)
    args = types.SimpleNamespace(start_col=4)

    def run(signal, w=2000):
        line = "\t".join(["f.fast5", "rid", "x", "y"] + [str(int(v)) for v in signal]) + "\n"
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf), np.errstate(all="ignore"), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ns["_one"](line, args, w)
        out = buf.getvalue().strip()
        if not out:
            return None
        _, _, a, b = out.split("\n")[0].split("\t")
        return [int(a), int(b)]
    return run


def main():
    run = reference_loop()
    rs = rollmean_inputs.reads()
    want = [run(r) for r in rs]
    # a second parameterisation (w = 700) on the first reads: the window length is a parameter of sqk_rollmean
    want_w700 = [run(r, 700) for r in rs[:12]]
    json.dump({"source": "dRNA_segmenter.py TSV-branch loop body executed from the reference file with w injected; pandas "
                         + pd.__version__, "w": 2000, "segments": want, "w700_first12": want_w700},
              open(os.path.join(HERE, "rollmean_golden.json"), "w"))
    print(f"{len(rs)} reads, {sum(w is not None for w in want)} with a segment; w=700: {sum(w is not None for w in want_w700)} of 12")


if __name__ == "__main__":
    main()
