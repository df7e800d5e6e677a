"""Wall-clock throughput of the drop-in command lines on a generated SquigglePull-style signal file (bench.py's
``cli_e2e`` block).  The file holds synthetic int16 reads of the benchmark shape behind eight leading columns (what
MotifSeq.py's ``l[8:]`` expects, MotifSeq.py:252-298), written with the library's own text writer
(squigglekit_b200.tsv, SquigglePull.py:243-253's format).  Each command line is run as a user would run it -- a fresh
``python MotifSeq.py -s ... -m ...`` process, stdout to /dev/null -- so interpreter start-up, CUDA context creation, text
parsing, the GPU path and the row formatting are all inside the number.  Nothing here touches the CPU oracle; the
reference-side rates are measured by bench.py.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEAD_COLS = 8


def _scratch_dir(need_bytes: int) -> str:
    for d in ("/dev/shm", tempfile.gettempdir()):
        try:
            if shutil.disk_usage(d).free > need_bytes * 1.3:
                return d
        except Exception:
            continue
    return tempfile.gettempdir()


def write_model(path: str, motif: np.ndarray, name: str = "bench_motif"):
    """The bait `.model` layout (MotifSeq.py:408-428): name, k-mer length, spare column, then the signal."""
    with open(path, "w") as fh:
        fh.write("\t".join([name, str(max(1, motif.size // 8)), "x"] + [repr(float(v)) for v in motif]) + "\n")


def generate(n_reads: int, n_samples: int, motif: np.ndarray, device=None, chunk: int = 8192):
    """-> (signal file path, model file path, seconds).  Reads come from synth.motifseq_reads_* (same generator as the
    kernel benchmark)."""
    from . import synth, tsv
    d = tempfile.mkdtemp(prefix="sqk_cli_", dir=_scratch_dir(n_reads * n_samples * 5))
    sig_path, model_path = os.path.join(d, "signal.tsv"), os.path.join(d, "motif.model")
    write_model(model_path, motif)
    t0 = time.perf_counter()
    with open(sig_path, "wb") as fh:
        for r0 in range(0, n_reads, chunk):
            n = min(chunk, n_reads - r0)
            if device is not None:
                sig = synth.motifseq_reads_torch(n, n_samples, motif, device, seed=synth.BASE_SEED + 1000 + r0).cpu().numpy().reshape(-1)
            else:
                sig = synth.motifseq_reads_np(n, n_samples, motif, seed=synth.BASE_SEED + 1000 + r0)[0]
            off = np.arange(n + 1, dtype=np.int64) * n_samples
            heads = ["\t".join([f"batch_{(r0 + i) // 4000}.fast5", f"read_{r0 + i:07d}", "8192.0", "6.0", "1467.61", "4000.0", "x", "y"])
                     for i in range(n)]
            tsv.write_reads(fh, heads, sig, off)
    return sig_path, model_path, time.perf_counter() - t0


def time_cli(script: str, argv, n_reads: int, timeout: float = 900.0):
    """Run `python <script> argv...` with stdout captured -> dict(reads/s, seconds, rows).  With SQK_CLI_PROFILE set the
    command line's own breakdown (its last stderr line) is returned as well."""
    t0 = time.perf_counter()
    p = subprocess.run([sys.executable, os.path.join(ROOT, script), *argv], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=timeout)
    dt = time.perf_counter() - t0
    if p.returncode != 0:
        raise RuntimeError(f"{script} exited with {p.returncode}: {p.stderr[-300:]!r}")
    rows = p.stdout.count(b"\n")
    out = {"value": n_reads / dt, "unit": "reads/s", "seconds": dt, "rows_printed": rows}
    prof = [ln for ln in p.stderr.decode("utf-8", "replace").split("\n") if ln.startswith("sqk profile")]
    if prof:
        out["profile"] = prof[-1]
    return out


def cleanup(sig_path: str):
    shutil.rmtree(os.path.dirname(sig_path), ignore_errors=True)
