"""Golden vectors for the dRNA adapter finder (dRNA_segmenter.py, slow5 branch, :86-176).

The reference keeps that algorithm inline in main() -- there is no function to import.  This script (build
container only: it reads /root/reference at run time, nothing of it is copied into the repo) cuts the body of the
`for read in s5.seq_reads():` loop out of the reference file, wraps it UNMODIFIED into a function and runs it on
  * the reads of example/slow5/0.blow5 (the reference's own sample file, parsed by squigglekit_b200.slow5) and
  * synthetic dRNA-like reads (low adapter plateau, then the RNA signal; interruptions, outliers, short reads),
capturing what it prints.  Output: adapter_inputs.npz (signals, offsets) + adapter_golden.json (per read: [start, end]
or null), both committed.

usage:  python tests/golden/make_adapter_golden.py
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys
import textwrap

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refload  # noqa: E402
from squigglekit_b200 import slow5  # noqa: E402

REF = refload.REFERENCE_ROOT


def reference_loop():
    """-> callable(read_dict) running the reference's loop body, returning what it printed."""
    src = open(os.path.join(REF, "dRNA_segmenter.py")).read().split("\n")
    first = next(i for i, l in enumerate(src) if l.strip() == "for read in s5.seq_reads():")
    last = next(i for i, l in enumerate(src) if i > first and l.strip().startswith("# for read in s5.seq_reads():"))
    body = textwrap.dedent("\n".join(src[first + 1:last]))
    scale_src = src[next(i for i, l in enumerate(src) if l.startswith("def scale_outliers")):]
    scale_src = "\n".join(scale_src[:next(i for i, l in enumerate(scale_src) if l.startswith("if __name__"))])
    code = "def _one(read, t_start, t_end):\n" + textwrap.indent(body, "    ") + "\n" + scale_src
    ns = {"np": np}
    exec(compile(code, "dRNA_segmenter_loop", "exec"), ns)

    def run(signal):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf), np.errstate(all="ignore"):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                ns["_one"]({"read_id": "r", "signal": np.asarray(signal, dtype=np.int16)}, 1000, 5000)
        out = buf.getvalue().strip()
        if not out:
            return None
        _, a, b = out.split("\n")[0].split("\t")
        return [int(a), int(b)]
    return run


def synthetic_reads(rng):
    reads = []
    for i in range(40):
        n = int(rng.integers(6000, 16000))
        adapter = int(rng.integers(1500, 7000))
        lo_lvl, hi_lvl = rng.uniform(380, 480), rng.uniform(560, 700)
        sig = np.where(np.arange(n) < adapter, lo_lvl, hi_lvl) + rng.normal(0, rng.uniform(5, 25), n)
        # dwell structure on the RNA part, dips back below the threshold, spikes inside the adapter
        lv = np.repeat(rng.normal(0, 35, n // 12 + 1), 12)[:n]
        sig[adapter:] += lv[adapter:]
        for _ in range(int(rng.integers(0, 6))):
            at = int(rng.integers(0, n - 50)); ln = int(rng.integers(1, 40))
            sig[at:at + ln] += rng.choice([-1, 1]) * rng.uniform(80, 250)
        if i % 5 == 0:                      # a second low stretch after the adapter (merge / break logic)
            at = min(n - 10, adapter + int(rng.integers(200, 3000))); ln = int(rng.integers(80, 2500))
            sig[at:at + ln] = lo_lvl + rng.normal(0, 10, min(ln, n - at))
        if i % 7 == 0:
            sig[rng.integers(0, n, 30)] = rng.choice([-5, 0, 1200, 1500, 3000], 30)   # outliers
        reads.append(np.clip(np.rint(sig), -32768, 32767).astype(np.int16))
    reads.append(rng.integers(300, 700, 500).astype(np.int16))       # shorter than t_start: NaN threshold
    reads.append(rng.integers(300, 700, 1001).astype(np.int16))      # one-sample statistics window
    reads.append(np.full(8000, 500, np.int16))                       # constant: nothing is < top
    reads.append(np.zeros(0, np.int16))
    reads.append(np.full(3000, 2000, np.int16))                      # all outliers
    reads.append(np.r_[np.full(4000, 400), np.full(4000, 650)].astype(np.int16))   # noiseless step
    return reads


def main():
    run = reference_loop()
    rng = np.random.default_rng(20240601)
    reads = [r["signal"] for r in slow5.read_blow5(os.path.join(REF, "example", "slow5", "0.blow5"))]
    n_real = len(reads)
    reads += synthetic_reads(rng)
    want = [run(r) for r in reads]
    offsets = np.zeros(len(reads) + 1, np.int64)
    np.cumsum([r.size for r in reads], out=offsets[1:])
    np.savez_compressed(os.path.join(HERE, "adapter_inputs.npz"), signals=np.concatenate(reads), offsets=offsets)
    json.dump({"source": "dRNA_segmenter.py slow5-branch loop body executed from the reference file",
               "n_real_reads": n_real, "segments": want}, open(os.path.join(HERE, "adapter_golden.json"), "w"))
    print(f"{len(reads)} reads ({n_real} from example/slow5/0.blow5), {sum(w is not None for w in want)} with a segment")


if __name__ == "__main__":
    main()
