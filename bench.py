#!/usr/bin/env python
"""bench.py -- MotifSeq reads/s (4096-sample int16 reads x 80-point motif, subsequence DTW, fp64 exact
results) on N B200s of one node, the HBM / issue rooflines of the dominant kernel, the CPU path timed beside
it, and (1 GPU, default workload) a block each for the segmenter hot path and the two command lines.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload configs2|configs3|configs4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

A step = one pass of the MotifSeq hot path (outlier removal -> z-score -> subsequence DTW -> hit record) over
one batch of synthetic reads per GPU (default: 100 000 x 4096, BASELINE.json configs[2]; weak scaling: every
rank gets its own batch; the only exchange is the gather of the 16-byte hit records, once per step).  `value`
is measured with the batch already resident in HBM; `e2e` goes through the C ABI with HOST buffers, H2D and D2H
inside the timed region.  Prints ONE JSON line on rank 0.  Exits non-zero if the parity guard fails.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_READS = 100_000          # per GPU
N_SAMPLES = 4096
N_MOTIF = 80
SCALE = "zscale"
METRIC = "MotifSeq reads/sec (4k-sample x 80-event subsequence DTW)"
BYTES_PER_READ = 2 * N_SAMPLES + 16         # int16 samples read once + one 16-byte hit record
CELLS_PER_READ = N_SAMPLES * N_MOTIF        # nominal (outlier removal drops ~0.05 % of the columns)
SEG_MAX_SEGS = 16
SEG_M = 4096
SEG_BYTES_PER_READ = 2 * SEG_M + 4 * (1 + 2 * SEG_MAX_SEGS)

# BASELINE.json configs (per GPU): name -> (reads, samples, host-buffer leg?)
WORKLOADS = {
    "configs2": (100_000, 4096, True),        # the metric's configuration (default)
    "configs3": (1_000_000, 20_000, False),   # HBM roofline run: 40 GB resident
    "configs4": (1_250_000, 50_000, False),   # per-GPU shard of 10 M x 50 k on 8 GPUs: 125 GB resident
}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def host_threads() -> int:
    """Hardware threads this process may run on -- NOT OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons, one sample every 50 ms.  The sampler runs from before the warm-up
    (nvidia-smi needs a moment to start); mark()/unmark() bracket a timed region and summary() only uses
    the samples that arrived inside it."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, enabled: bool = True):
        self.index, self.rows, self.proc, self.enabled = index, [], None, enabled
        self.t0 = self.t1 = None

    def __enter__(self):
        if not self.enabled:                                     # (only the rank that prints the line samples its GPU)
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def wait_first(self, timeout: float = 15.0):
        """Block until the first sample has arrived (nvidia-smi takes seconds to start on an 8-GPU box; a timed region
        of a few steps would otherwise be over before the sampler has said anything).  Called outside the timed region."""
        t_end = time.perf_counter() + timeout
        while self.proc is not None and not self.rows and self.proc.poll() is None and time.perf_counter() < t_end:
            time.sleep(0.01)

    def mark(self):
        self.t0, self.t1 = time.perf_counter(), None

    def unmark(self):
        self.t1 = time.perf_counter()

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def digest(rows):
            sm, mx, pw, reasons = [], [], [], set()
            for _, r in rows:
                try:
                    sm.append(float(r[0])); mx.append(float(r[1]))
                except Exception:
                    continue
                try:
                    pw.append(float(r[2]))
                except Exception:
                    pass
                for n, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            return sm, mx, pw, reasons

        inside = [x for x in self.rows if self.t0 is not None and self.t0 <= x[0] <= (self.t1 or 1e30)]
        sm, mx, pw, reasons = digest(inside)
        sm_all, mx_all, pw_all, reasons_all = digest(self.rows)
        if not sm_all:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        use = sm if sm else sm_all
        return {"sm_mhz": statistics.median(use), "sm_max_mhz": max(mx_all), "reasons": sorted(reasons if sm else reasons_all),
                "samples": len(sm), "samples_incl_warmup": len(sm_all), "sm_mhz_min_incl_warmup": min(sm_all),
                "power_w_max": max(pw) if pw else (max(pw_all) if pw_all else None),
                "reasons_incl_warmup": sorted(reasons_all)}


def cpu_reference_rate(signals, offsets, motif, threads: int, target_s: float):
    """The reference's per-read path on host cores: scale_outliers -> z-score -> full N x M matrix DTW +
    back-trace (oracle port of mlpy 3.5.0; the reference's Python cannot run here without mlpy).
    Times a bounded sample sized for ~target_s seconds.  -> (reads/s, n_sample, seconds)."""
    import oracle
    n_all = offsets.size - 1
    probe = min(64 * max(1, threads), n_all)
    t0 = time.perf_counter()
    oracle.motifseq_batch(signals, offsets[:probe + 1], motif, scale=SCALE, full_matrix=True, n_threads=threads)
    dt = max(time.perf_counter() - t0, 1e-6)
    want = int(max(probe, probe / dt * target_s))
    done, t0 = 0, time.perf_counter()
    while done < want:                         # several passes over the sample when the box is fast
        n = min(n_all, want - done)
        oracle.motifseq_batch(signals, offsets[:n + 1], motif, scale=SCALE, full_matrix=True, n_threads=threads)
        done += n
    dt = time.perf_counter() - t0
    return done / dt, done, dt


def bench_motif():
    """Default: 10 levels x dwell 8 = 80 points; other lengths keep dwell 8 and cut to size."""
    from squigglekit_b200 import synth
    return synth.make_motif(n_levels=(N_MOTIF + 7) // 8, dwell=8)[:N_MOTIF]


def run_reference(args):
    """--impl reference: the CPU implementation of the path on the box's host cores, same workload shape.  The thread
    count is --ref-threads or every hardware thread this process may use (never OMP_NUM_THREADS, which the launcher
    sets to 1 under torchrun), so the arm is the same at every --gpus N; the one-thread rate -- how the reference
    really runs -- is printed beside it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle
    from squigglekit_b200 import synth
    threads = args.ref_threads if args.ref_threads > 0 else host_threads()
    motif = bench_motif()
    n_gen = 4096 * max(1, min(threads, 16) // 4)
    sig, off, _ = synth.motifseq_reads_np(min(n_gen, N_READS), N_SAMPLES, motif)
    for _ in range(args.warmup):
        oracle.motifseq_batch(sig, off[:257], motif, scale=SCALE, full_matrix=True, n_threads=threads)
    rates, n_used, secs = [], 0, 0.0
    t_all = time.perf_counter()
    for _ in range(args.steps):
        r, n_used, dt = cpu_reference_rate(sig, off, motif, threads, target_s=min(6.0, 90.0 / max(1, args.steps)))
        rates.append(r); secs += dt
    value = statistics.median(rates)
    one_rate, one_n, one_dt = cpu_reference_rate(sig, off, motif, 1, target_s=5.0)
    sample = (f"{n_used} reads of {N_SAMPLES} samples per step ({N_SAMPLES}x{N_MOTIF} cells each), synthetic, "
              f"{threads} OpenMP threads, full N x M float64 cost matrix + back-trace per read as mlpy does")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"MotifSeq {N_MOTIF}-point motif vs synthetic {N_SAMPLES}-sample int16 reads "
                               f"(BASELINE configs[2] shape), bounded sample per step", "scale": SCALE,
                   "reads_per_step": n_used, "n_samples": N_SAMPLES, "n_motif": N_MOTIF,
                   "same_shape_as_native_arm": True, "same_read_count_as_native_arm": False,
                   "note": "a bounded sample of the same read shape per step (the full 100 000-read batch would take minutes on the CPU); rates are per read, so they compare"},
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": threads, "kind": "port", "sample": sample,
                         "one_thread_value": one_rate,
                         "one_thread_sample": f"{one_n} reads in {one_dt:.1f} s, 1 thread (the reference script is single-threaded)",
                         "threads_source": "--ref-threads" if args.ref_threads > 0 else "os.sched_getaffinity (OMP_NUM_THREADS ignored)"},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line), flush=True)
    return 0


def ubench_cells_per_s(which="dtw_step", prec="fp64"):
    """Register-only micro-benchmark of the DTW step's instruction stream (squigglekit_b200/sqk_ubench): what the same
    instructions reach without memory, refills or hand-overs.  Reported as `ubench_frac`, NOT as the roofline: the
    roofline is the hardware's issue rate.  Falls back to the committed measurement in profiles/."""
    exe = os.path.join(ROOT, "squigglekit_b200", "sqk_ubench")
    best, src = None, None
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=60).stdout
        src = "sqk_ubench run live"
    except Exception:
        out = ""
    if not out.strip():
        for name in sorted(os.listdir(os.path.join(ROOT, "profiles")), reverse=True):
            if name.startswith("ubench") and name.endswith(".jsonl"):
                out = open(os.path.join(ROOT, "profiles", name)).read()
                src = f"profiles/{name}"
                break
    for ln in out.splitlines():
        try:
            d = json.loads(ln)
        except Exception:
            continue
        if d.get("bench") in which.split("|") and d.get("precision") == prec:
            best = max(best or 0.0, float(d["cells_per_s"]))
    return best, src


def profile_traffic():
    """DRAM bytes per launch from the committed ncu captures (profiles/dtw_traffic.json): not measurable inside a
    timed run, so the line says where each figure comes from."""
    try:
        with open(os.path.join(ROOT, "profiles", "dtw_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


# ------------------------------------------------------------------------------------------------
# segmenter block (hot path B, BASELINE configs[1]) -- 1 GPU, default workload only
# ------------------------------------------------------------------------------------------------
def segmenter_block(ctx, dev, peak, steps):
    import torch

    import oracle
    import squigglekit_b200 as sqk
    from squigglekit_b200 import synth

    cfg = sqk.SegConfig(stall=True, max_segs=SEG_MAX_SEGS)          # -ku: stall detection, test_segs on the host
    out = {"metric": "segmenter reads/sec (4096-sample int16 reads, get_segs -ku)", "unit": "reads/s", "runs": []}
    for R, e2e_reads in ((10_000, 10_000), (1_000_000, 250_000)):
        sig = synth.segmenter_reads_torch(R, SEG_M, dev).view(-1)
        off = torch.arange(R + 1, dtype=torch.int64, device=dev) * SEG_M
        for _ in range(3):
            segs, nsegs = ctx.segmenter(sig, off, cfg, max_read_len=SEG_M)
        torch.cuda.synchronize()
        ctx.enable_timing(True); ctx.timing(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            segs, nsegs = ctx.segmenter(sig, off, cfg, max_read_len=SEG_M)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        kt = ctx.timing(reset=True); ctx.enable_timing(False)
        # parity on a sub-sample against the CPU oracle (pinned on the reference's own get_segs, tests/test_oracle.py)
        idx = np.arange(0, R, max(1, R // 512))[:512]
        sub = sig.view(R, SEG_M)[torch.from_numpy(idx).to(dev)].cpu().numpy().reshape(-1)
        suboff = np.arange(idx.size + 1, dtype=np.int64) * SEG_M
        want, wn = oracle.segmenter_batch(sub, suboff, oracle.SegCfg(), 0, 900, 0, SEG_MAX_SEGS)
        got, gn = segs.cpu().numpy()[idx], nsegs.cpu().numpy()[idx]
        m = np.arange(SEG_MAX_SEGS)[None, :, None] < wn[:, None, None]
        parity = bool(np.array_equal(gn, wn) and np.array_equal(np.where(m, got, 0), np.where(m, want, 0)))
        # e2e: pinned host buffers through the C ABI, H2D + D2H inside the timed region
        h_sig = sqk.pinned_empty(e2e_reads * SEG_M, np.int16)
        h_sig[:] = sig[: e2e_reads * SEG_M].cpu().numpy()
        h_off = off[: e2e_reads + 1].cpu().numpy()
        for _ in range(2):
            ctx.segmenter(h_sig, h_off, cfg, max_read_len=SEG_M)
        n_e2e = 3
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            ctx.segmenter(h_sig, h_off, cfg, max_read_len=SEG_M)
        e2e_s = (time.perf_counter() - t0) / n_e2e
        sqk.pinned_free(h_sig)
        stats_ms = kt["stats"]["ms"] / max(1, kt["stats"]["launches"])
        fsm_ms = kt["seg_fsm"]["ms"] / max(1, kt["seg_fsm"]["launches"])
        run = {"reads": R, "value": R / (ms * 1e-3), "ms_per_step": ms,
               "kernels_ms": {"stats (sqk_stats3_kernel + redo list)": stats_ms, "state machine (sqk_fsm_mask_kernel + redo list)": fsm_ms},
               "roofline": {"bound": "hbm", "unit": "GB/s", "peak": peak, "algorithmic_bytes_per_read": SEG_BYTES_PER_READ,
                            "achieved_step": R * SEG_BYTES_PER_READ / (ms * 1e-3) / 1e9,
                            "frac_step": R * SEG_BYTES_PER_READ / (ms * 1e-3) / 1e9 / peak,
                            "frac_stats_kernel": R * 2 * SEG_M / (stats_ms * 1e-3) / 1e9 / peak if stats_ms > 0 else None},
               "e2e": {"value": e2e_reads / e2e_s, "unit": "reads/s", "reads": e2e_reads,
                       "h2d_bytes_per_step": e2e_reads * SEG_M * 2 + (e2e_reads + 1) * 8,
                       "d2h_bytes_per_step": e2e_reads * (SEG_MAX_SEGS * 8 + 4)},
               "parity_subsample_bit_exact": parity, "parity_reads_checked": int(idx.size),
               "segments_found_mean": float(nsegs.float().mean().item())}
        out["runs"].append(run)
        if R == 10_000:
            # CPU baseline on the same reads: the reference's REAL pure-Python get_segs when its tree is present
            # (build container), else the C port of it (GPU box)
            from oracle import refload
            sub_n = 96
            kind = "port"
            if refload.available():
                try:
                    import types
                    ref = refload.load("segmenter")
                    a = types.SimpleNamespace(error=5, corrector=50, window=150, seg_dist=50, std_scale=0.75, stall_len=0.25,
                                              lim_hi=900, lim_low=0)
                    t0 = time.perf_counter()
                    for r in range(sub_n):
                        s = sub[r * SEG_M:(r + 1) * SEG_M - 1].astype(int)      # sig[:Num] with Num = -1
                        s = ref.scale_outliers(s, a)
                        ref.get_segs(s, a)
                    dt = time.perf_counter() - t0
                    out["cpu_baseline"] = {"value": sub_n / dt, "unit": "reads/s", "cores": 1, "kind": "reference",
                                           "sample": f"{sub_n} reads through the reference's own scale_outliers + get_segs (pure Python), {dt:.1f} s"}
                    kind = "reference"
                except Exception:
                    kind = "port"
            if kind == "port":
                t0 = time.perf_counter()
                reps = 0
                while time.perf_counter() - t0 < 3.0:
                    oracle.segmenter_batch(sub, suboff, oracle.SegCfg(), 0, 900, 0, SEG_MAX_SEGS, n_threads=1)
                    reps += 1
                dt = time.perf_counter() - t0
                out["cpu_baseline"] = {"value": reps * idx.size / dt, "unit": "reads/s", "cores": 1, "kind": "port",
                                       "sample": f"{reps * idx.size} reads, C restatement of scale_outliers + get_segs on 1 core, {dt:.1f} s "
                                                 "(the reference's pure-Python loop runs ~640 reads/s, SURVEY §6)"}
        del sig, off, segs, nsegs
        torch.cuda.empty_cache()
    big = out["runs"][-1]
    out["value"] = big["value"]
    out["note"] = "value = the 1 M-read run (configs[1]'s 10 k reads are 83 MB: launch-latency bound, listed in runs[0])"
    return out


def cli_block(dev, motif):
    """Wall-clock reads/s of the two drop-in command lines (fresh process each: interpreter start, CUDA context, text
    parsing, GPU path, row formatting) on a generated SquigglePull-style TSV of the benchmark's read shape (page cache
    warm), next to the reference's per-line path on the first lines of the same file."""
    import contextlib
    import io
    import types

    import oracle
    from oracle import refload
    from squigglekit_b200 import cli_bench
    n = int(os.environ.get("SQK_CLI_BENCH_READS", "100000"))
    out = {"reads": n, "n_samples": 4096, "file": "fast5, readID, 6 spare columns, 4096 int16 samples per line (SquigglePull.py:243-253 layout)"}
    sig_path, model_path, gen_s = cli_bench.generate(n, 4096, motif, device=dev)
    try:
        out["file_bytes"] = os.path.getsize(sig_path)
        out["generate_seconds"] = gen_s
        # each command line twice: the first run of a fresh box pays the page-cache misses of the interpreter's imports and of
        # the CUDA libraries (seconds, nothing to do with this code); the second one is reported, the first one kept beside it
        os.environ["SQK_CLI_PROFILE"] = "1"
        for key, script, argv in (("motifseq", "MotifSeq.py", ["-s", sig_path, "-m", model_path, "--scale", "zscale"]),
                                  ("segmenter", "segmenter.py", ["-s", sig_path, "--start_col", "8", "-k", "-u"])):
            first = cli_bench.time_cli(script, argv, n)
            out[key] = cli_bench.time_cli(script, argv, n)
            out[key]["first_run_seconds"] = first["seconds"]
            prof = out[key].get("profile")
            if prof:
                try:
                    inside = sum(float(tok) for tok in prof.replace(",", " ").split() if tok.replace(".", "", 1).isdigit())
                    out[key]["steady_state"] = {"value": n / inside, "unit": "reads/s", "seconds": inside,
                                                "what": "reads / the time inside the read loop (parse, GPU, format, write): what a long file converges to once the ~1-2 s of process start, imports and CUDA context creation are amortised"}
                except Exception:
                    pass
        # ---- the reference's per-line path on the first lines of the same file, one core ------------------------------
        n_ref = 200
        with open(sig_path, "rt") as fh:
            lines = [next(fh) for _ in range(n_ref)]
        import scipy.stats as st
        L = max(1, motif.size // 8)
        t0 = time.perf_counter()
        sink = io.StringIO()
        for line in lines:                       # MotifSeq.py:267-298 + get_region_multi :431-449, mlpy -> the oracle's C port
            l = line.strip("\n").split("\t")
            sig = np.array([float(i) for i in l[8:]], dtype=float)
            sig = sig[(sig > 0) & (sig < 1200)]
            sig = (sig - sig.mean()) / sig.std()
            dist, cost, path = oracle.dtw_subsequence(motif, sig)            # MotifSeq.py:437-439
            start, end = path[1][0], path[1][-1]
            mod_mean = (2.90 * L) + -9.6
            mod_stdev = mod_mean * 0.08468
            Z = (dist - mod_mean) / mod_stdev
            p_value = st.norm.cdf(Z)
            sink.write("{}\t{}\t{}\t{}\t{}\t{}\t{}\t{}\t{}\t{}\t{}\t{}\n".format(l[0], l[1], "m", start, end, end - start, dist, mod_mean, mod_stdev, Z, p_value, (1 - p_value) * 100))
        dt = time.perf_counter() - t0
        out["motifseq_reference"] = {"value": n_ref / dt, "unit": "reads/s", "cores": 1, "kind": "port",
                                     "sample": f"first {n_ref} lines: the reference's per-line Python (float() per field, outlier cut, z-score, row formatting) with mlpy.dtw_subsequence replaced by the oracle's C port (full matrix + back-trace), {dt:.1f} s"}
        kind = "port"
        if refload.available():
            try:
                ref = refload.load("segmenter")
                a = types.SimpleNamespace(error=5, corrector=50, window=150, seg_dist=50, std_scale=0.75, stall_len=0.25,
                                          lim_hi=900, lim_low=0, stall=True, stall_start=300, gap=False, gap_dist=3000)
                t0 = time.perf_counter()
                with contextlib.redirect_stdout(sink), contextlib.redirect_stderr(sink):
                    for line in lines:           # segmenter.py:194-230
                        l = line.strip("\n").split("\t")
                        sig = np.array([int(i) for i in l[8:]], dtype=int)
                        sig = ref.scale_outliers(sig[:-1], a)
                        ref.get_segs(sig, a)
                dt = time.perf_counter() - t0
                kind = "reference"
                out["segmenter_reference"] = {"value": n_ref / dt, "unit": "reads/s", "cores": 1, "kind": "reference",
                                              "sample": f"first {n_ref} lines through the reference's own parsing, scale_outliers and get_segs (pure Python), {dt:.1f} s"}
            except Exception:
                kind = "port"
        if kind == "port":
            t0 = time.perf_counter()
            for line in lines:
                l = line.strip("\n").split("\t")
                sig = np.array([int(i) for i in l[8:]], dtype=int)[:-1]
                sig = sig[(sig > 0) & (sig < 900)]
                oracle.get_segs(sig)
            dt = time.perf_counter() - t0
            out["segmenter_reference"] = {"value": n_ref / dt, "unit": "reads/s", "cores": 1, "kind": "port",
                                          "sample": f"first {n_ref} lines: the reference's per-line Python parsing with get_segs replaced by its C restatement, {dt:.1f} s (the pure-Python get_segs runs ~640 reads/s, SURVEY §6)"}
        out["motifseq"]["vs_reference"] = out["motifseq"]["value"] / out["motifseq_reference"]["value"]
        out["segmenter"]["vs_reference"] = out["segmenter"]["value"] / out["segmenter_reference"]["value"]
    finally:
        cli_bench.cleanup(sig_path)
    return out


def main():
    global N_SAMPLES, N_MOTIF, BYTES_PER_READ, CELLS_PER_READ, SCALE
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="configs2", choices=sorted(WORKLOADS),
                    help="BASELINE.json configs[] entry: configs2 = 100k x 4096 (default, the metric), configs3 = 1M x 20000, "
                         "configs4 = 1.25M x 50000 per GPU (8 GPUs = the 10M-read job)")
    ap.add_argument("--reads", type=int, default=0, help="reads per GPU per step (overrides the workload)")
    ap.add_argument("--samples", type=int, default=0, help="samples per read (overrides the workload)")
    ap.add_argument("--motif-len", type=int, default=80, help="motif points (default 80; 163 = the reference's example model)")
    ap.add_argument("--scale", default="zscale", choices=["zscale", "medmad"], help="normalisation (BASELINE: zscale; the reference's default is medmad)")
    ap.add_argument("--lanes", type=int, default=0, help="force lanes-per-read of the DTW kernel (experiments)")
    ap.add_argument("--precision", default="fp64", choices=["fp64", "fp32"])
    ap.add_argument("--plan", default="auto", choices=["auto", "single_pass", "two_pass"],
                    help="how exact (fp64) requests run: float64 recurrence over every column, or float32 lower-bound scan + float64 windows (same bits)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl", "none"],
                    help="multi-GPU gather of the hit records: p2p = the producing kernels store every record into every peer's "
                         "buffer over NVLink (no collective kernel); nccl = all_gather_into_tensor; none = no exchange (experiments)")
    ap.add_argument("--ref-threads", type=int, default=0, help="--impl reference: host threads (default: all hardware threads, ignoring OMP_NUM_THREADS)")
    ap.add_argument("--seed-offset", type=int, default=0, help="added to the per-rank data seed (reproduce another rank's batch on one GPU)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer legs (very large batches: they need the batch in host memory)")
    ap.add_argument("--no-extras", action="store_true", help="skip the sustained twin, the segmenter block and the CLI block")
    args = ap.parse_args()
    SCALE = args.scale
    w_reads, w_samples, w_e2e = WORKLOADS[args.workload]
    R = args.reads or w_reads
    N_SAMPLES, N_MOTIF = (args.samples or w_samples), args.motif_len
    BYTES_PER_READ = 2 * N_SAMPLES + 16
    CELLS_PER_READ = N_SAMPLES * N_MOTIF
    if not w_e2e and not args.samples and not args.reads:
        args.no_e2e = True
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import squigglekit_b200 as sqk
    from squigglekit_b200 import synth
    from squigglekit_b200.dist import GatherPipeline, PeerGather, env_rank_world

    rank, world, local_rank = env_rank_world()
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)

    M = N_SAMPLES
    default_shape = (args.workload == "configs2" and R == 100_000 and M == 4096 and N_MOTIF == 80 and SCALE == "zscale"
                     and args.precision == "fp64")
    motif = bench_motif()
    ctx = sqk.Context(local_rank)
    if args.lanes:
        ctx.set_dtw_lanes(args.lanes)
    ctx.set_dtw_plan(args.plan)
    sig = synth.motifseq_reads_torch(R, M, motif, dev, seed=synth.BASE_SEED + rank + args.seed_offset).view(-1)
    off = torch.arange(R + 1, dtype=torch.int64, device=dev) * M
    exchange = args.exchange
    if exchange == "auto":
        exchange = "p2p" if world > 1 else "none"
    if world == 1:
        exchange = "none"
    # The gather of the 16-byte hit records, once per step, never makes the ranks march in lockstep:
    #   p2p : the kernels that produce a record store it into every peer's gathered buffer through P2P-mapped pointers
    #         (fused compute + all-gather over NVLink, no collective kernel on the SMs); one flag per step per peer
    #   nccl: an asynchronous all_gather_into_tensor per step, double-buffered
    if exchange == "p2p":
        pipe = PeerGather(ctx, R, 1, dev, depth=2)
    else:
        pipe = GatherPipeline((R, 1, 16), torch.uint8, dev, depth=2, enabled=(exchange == "nccl"))
    hits = None

    def step():
        nonlocal hits
        hits = pipe.local_buffer()
        ctx.motifseq(sig, off, motif, scale=SCALE, precision=args.precision, max_read_len=M, out=hits, want_kept=False)
        return pipe.submit()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sustained = None
    with ClockSampler(local_rank, enabled=(rank == 0)) as clocks:
        for _ in range(args.warmup):
            gathered = step()
        pipe.drain()
        barrier()
        ctx.enable_timing(True)
        ctx.timing(reset=True)
        ctx.launches(reset=True)
        clocks.wait_first()
        barrier()
        clocks.mark()
        ev0.record()
        for _ in range(args.steps):
            gathered = step()
        pipe.drain()
        ev1.record()
        barrier()
        clocks.unmark()
        ms = ev0.elapsed_time(ev1)
        kt = ctx.timing(reset=True)
        n_launches = ctx.launches(reset=True)
        ctx.enable_timing(False)
        clock_summary = clocks.summary()
        if default_shape and not args.no_extras:
            # the sustained twin of the burst number: the same step repeated for >= 3 s, clocks recorded over it
            sev0, sev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n_sus = max(args.steps, int(3.2e3 / max(ms / args.steps, 1e-3)))
            barrier()
            clocks.mark()
            sev0.record()
            for _ in range(n_sus):
                gathered = step()
            pipe.drain()
            sev1.record()
            barrier()
            clocks.unmark()
            sus_ms = sev0.elapsed_time(sev1)
            ts = torch.tensor([sus_ms], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ts, op=dist.ReduceOp.MAX)
            sustained = {"value": world * R * n_sus / (float(ts.item()) * 1e-3), "unit": "reads/s", "steps": n_sus,
                         "seconds": float(ts.item()) * 1e-3, "ms_per_step": float(ts.item()) / n_sus, "clocks": clocks.summary()}
    plan_counters = ctx.plan_counters()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * R * args.steps / (ms_max * 1e-3)

    # per-rank kernel times (ms per launch) and step times: what limits the scaling is visible in the line
    per_rank = None
    mine = torch.tensor([kt[k]["ms"] / max(1, kt[k]["launches"]) for k in ("stats", "dtw_lb", "dtw_win", "dtw")] + [ms / args.steps],
                        dtype=torch.float64, device=dev)
    if world > 1:
        allr = torch.empty((world, mine.numel()), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allr, mine)
        a = allr.cpu().numpy()
        per_rank = {"stats_ms": [round(float(x), 4) for x in a[:, 0]], "dtw_lb_ms": [round(float(x), 4) for x in a[:, 1]],
                    "dtw_win_ms": [round(float(x), 4) for x in a[:, 2]], "dtw_ms": [round(float(x), 4) for x in a[:, 3]],
                    "step_ms": [round(float(x), 4) for x in a[:, 4]]}

    # ---- parity guard on what was just timed: a fixed sub-sample against the CPU oracle -------------
    parity = None
    parity_ok = True
    if rank == 0:
        import oracle
        idx = np.arange(0, R, max(1, R // 256))[:256]
        sub = sig.view(R, M)[torch.from_numpy(idx).to(dev)].cpu().numpy().reshape(-1)
        suboff = np.arange(idx.size + 1, dtype=np.int64) * M
        want, _ = oracle.motifseq_batch(sub, suboff, motif, scale=SCALE, full_matrix=False)
        got = sqk.hits_from_torch(hits)[idx, 0]
        idx_ok = bool(np.array_equal(got["start"], want["start"]) and np.array_equal(got["end"], want["end"]))
        if args.precision == "fp64":
            parity = {"reads_checked": int(idx.size), "indices_bit_exact": idx_ok,
                      "dist_bit_exact": bool(np.array_equal(got["dist"], want["dist"]))}
            parity_ok = idx_ok and parity["dist_bit_exact"]
        else:
            parity = {"reads_checked": int(idx.size),
                      "index_mismatch_rate": float(np.mean((got["start"] != want["start"]) | (got["end"] != want["end"]))),
                      "dist_max_rel_err": float(np.max(np.abs(got["dist"] - want["dist"]) / np.maximum(want["dist"], 1e-9)))}
        if world > 1 and exchange != "none":
            # What crossed NVLink, checked: rank 0 regenerates the first reads of OTHER ranks' batches (read r of rank q
            # is a pure function of (seed + q, r)), runs the oracle on them and compares with that rank's block of the
            # gathered records, bytes and all.
            g = gathered.view(world, R, 16)
            ok_all, checked = True, []
            n_chk = min(R, 256)
            first = min(R, 16384)          # the generator's first chunk (synth.motifseq_reads_torch)
            for q in sorted({1, world - 1}):
                rs = synth.motifseq_reads_torch(first, M, motif, dev, seed=synth.BASE_SEED + q)[:n_chk].cpu().numpy().reshape(-1)
                w2, _ = oracle.motifseq_batch(rs, np.arange(n_chk + 1, dtype=np.int64) * M, motif, scale=SCALE, full_matrix=False)
                got_q = g[q, :n_chk].cpu().numpy().reshape(n_chk, 16).view(sqk.HIT_DTYPE).reshape(n_chk)
                ok = bool(np.array_equal(got_q["start"], w2["start"]) and np.array_equal(got_q["end"], w2["end"])
                          and (args.precision != "fp64" or np.array_equal(got_q["dist"], w2["dist"])))
                ok_all &= ok
                checked.append(q)
            own = bool(torch.equal(g[0], hits.view(R, 16)))
            parity["gathered_remote_bit_exact"] = ok_all
            parity["gathered_remote_ranks_checked"] = checked
            parity["gathered_remote_reads_per_rank"] = n_chk
            parity["gathered_own_block_matches"] = own
            parity_ok = parity_ok and ok_all and own

    e2e_value, e2e_steps, h_sig, h_hits, e2e_pageable = None, 0, None, None, None
    h_off = off.cpu().numpy()
    if not args.no_e2e:
        # ---- e2e: host buffers through the C ABI, H2D + D2H inside the timed region ---------------------
        h_sig = sqk.pinned_empty(R * M, np.int16)
        h_sig[:] = sig.cpu().numpy()
        h_hits = sqk.pinned_empty((R, 1), sqk.HIT_DTYPE)
        e2e_steps = max(2, min(args.steps, 5))
        for _ in range(2):
            ctx.motifseq(h_sig, h_off, motif, scale=SCALE, precision=args.precision, max_read_len=M, out=h_hits, want_kept=False)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            ctx.motifseq(h_sig, h_off, motif, scale=SCALE, precision=args.precision, max_read_len=M, out=h_hits, want_kept=False)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_value = world * R * e2e_steps / float(te.item())
        if rank == 0 and args.precision == "fp64":
            parity["e2e_matches_device_path"] = bool(np.array_equal(h_hits.view(np.uint8).reshape(R, 16), hits.cpu().numpy().reshape(R, 16)))
            parity_ok = parity_ok and parity["e2e_matches_device_path"]
        if world == 1 and not args.no_extras:
            # the same call with PAGEABLE numpy arrays (what a caller of the reference holds)
            p_sig = np.array(h_sig, copy=True)
            p_hits = np.zeros((R, 1), dtype=sqk.HIT_DTYPE)
            ctx.motifseq(p_sig, h_off, motif, scale=SCALE, precision=args.precision, max_read_len=M, out=p_hits, want_kept=False)
            t0 = time.perf_counter()
            for _ in range(3):
                ctx.motifseq(p_sig, h_off, motif, scale=SCALE, precision=args.precision, max_read_len=M, out=p_hits, want_kept=False)
            e2e_pageable = {"value": R * 3 / (time.perf_counter() - t0), "unit": "reads/s",
                            "matches_pinned_result": bool(np.array_equal(p_hits.view(np.uint8), h_hits.view(np.uint8)))}
            del p_sig

    if rank == 0:
        peak, peak_src = measured_peaks()
        two_pass = kt["dtw_lb"]["launches"] > 0
        if two_pass:
            # dominant kernel: the float32 lower-bound scan (every sample of every read goes through it once)
            kname, kkey, ub, ubp, instr_per_cell = "sqk_dtw_lb_kernel", "dtw_lb", "lb_step|lb_step2", "fp32_rd", 3
        else:
            kname, kkey, ub, ubp, instr_per_cell = "sqk_dtw_kernel", "dtw", "dtw_step", ("fp64" if args.precision == "fp64" else "fp32"), 10
        dtw_launches = max(1, kt[kkey]["launches"])
        dtw_ms = kt[kkey]["ms"] / dtw_launches
        achieved = R * BYTES_PER_READ / (dtw_ms * 1e-3) / 1e9
        ub_peak, ub_src = ubench_cells_per_s(ub, ubp)
        cells_s = R * CELLS_PER_READ / (dtw_ms * 1e-3)
        # hardware issue ceiling: 4 schedulers x 32 lanes = 128 thread-instructions per clock per SM, at the SM clock seen
        # during the run, divided by the instructions one cell of the recurrence needs (3 for the lower-bound scan:
        # FADD.RZ, FMNMX3, FADD.RM; ~10 for the float64 recurrence with start pointers)
        sm_mhz = clock_summary.get("sm_mhz") or clock_summary.get("sm_max_mhz") or ctx.clock_khz / 1e3
        issue_ceiling = ctx.n_sms * 128 * sm_mhz * 1e6 / instr_per_cell
        tr = profile_traffic() if default_shape else {}
        stats_ms = kt["stats"]["ms"] / max(1, kt["stats"]["launches"])
        win_ms = (kt["dtw_win"]["ms"] / max(1, kt["dtw_win"]["launches"])) if two_pass else None
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            sub_n = min(8192, R)
            sub_sig = sig[: sub_n * M].cpu().numpy()
            rate, n_used, secs = cpu_reference_rate(sub_sig, h_off[: sub_n + 1], motif, threads=1, target_s=12.0)
            cpu = {"value": rate, "unit": "reads/s", "cores": 1, "kind": "port",
                   "sample": f"first {n_used} reads of the same batch, {secs:.1f} s, 1 thread (the reference is single-threaded): "
                             f"scale_outliers + z-score + full {N_MOTIF}x{M} float64 cost matrix + back-trace per read (oracle port of mlpy 3.5.0)"}
        line = {
            "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if args.precision == "fp64" else "f32", "data": "synthetic",
            "config": {"workload": f"MotifSeq {N_MOTIF}-point motif vs {R} synthetic {M}-sample int16 reads per GPU "
                                   f"(BASELINE {args.workload}{'' if not (args.reads or args.samples) else ', shape overridden'}); {SCALE}; outlier window (0,1200)",
                       "baseline_config": args.workload, "reads_per_gpu": R, "n_samples": M, "n_motif": N_MOTIF, "scale": SCALE,
                       "precision": args.precision,
                       "l2_policy": f"input {R * M * 2 / 1e6:.0f} MB per step > 126 MB L2 (no flush needed)",
                       "timing": "CUDA events on the launching stream around K steps, max over ranks; e2e = wall clock of the synchronous host-buffer C-ABI call",
                       "exchange": {"p2p": "the kernels that produce a hit record store it into every peer's gathered buffer through P2P-mapped pointers (NVLink), one flag per step per peer; drained inside the timed region",
                                    "nccl": "one all_gather_into_tensor of the 16-byte hit records per step, double-buffered; drained inside the timed region",
                                    "none": "none (1 GPU)" if world == 1 else "disabled (--exchange none)"}[exchange],
                       "exchange_kind": exchange,
                       "dtw_lanes_per_read": args.lanes or "auto", "dtw_plan": args.plan},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": tr.get("dram_bytes_per_launch_100k_reads_lb" if two_pass else "dram_bytes_per_launch_100k_reads"),
                         "traffic_source": (tr.get("lb_source") if two_pass else tr.get("source")) if tr else None,
                         "traffic_note": "from a committed ncu capture of the same kernel and shape, not measured in this run" if tr else None,
                         "step_traffic_bytes": tr.get("dram_bytes_per_step_100k_reads_two_pass") if two_pass else None,
                         "step_algorithmic_bytes": R * BYTES_PER_READ,
                         "peak_source": peak_src, "kernel": kname,
                         "kernel_ms_per_launch": dtw_ms, "algorithmic_bytes_per_read": BYTES_PER_READ,
                         "note": "the DTW recurrence is ALU-issue bound (SURVEY F7): see roofline_alu; traffic figures come from the committed ncu captures, not from this run",
                         "stats_kernel_ms_per_launch": stats_ms,
                         "stats_kernel_frac": (R * 2 * M / (stats_ms * 1e-3) / 1e9 / peak) if stats_ms > 0 else None,
                         "exact_windows_ms_per_step": win_ms},
            "plan": {"name": "two_pass" if two_pass else "single_pass",
                     "exact_windows_per_step": plan_counters["windows"] if two_pass else None,
                     "second_attempt_windows_per_step": plan_counters.get("second_attempt_windows") if two_pass else None,
                     "full_length_fallback_reads_per_step": plan_counters["fallback_reads"] if two_pass else None},
            "roofline_alu": {"achieved_cells_per_s": cells_s, "issue_ceiling_cells_per_s": issue_ceiling,
                             "frac": cells_s / issue_ceiling,
                             "ceiling": f"{ctx.n_sms} SMs x 128 thread-instructions/clk x {sm_mhz:.0f} MHz / {instr_per_cell} instructions per cell",
                             "ubench_cells_per_s": ub_peak, "ubench_frac": (cells_s / ub_peak) if ub_peak else None, "ubench_source": ub_src},
            "per_rank": per_rank,
            "cpu_baseline": cpu,
            "clocks": clock_summary,
            "sustained": sustained,
            "e2e": ({"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": int(R * M * 2 + (R + 1) * 8),
                     "d2h_bytes_per_step": int(R * 16), "steps": e2e_steps, "host_buffers": "pinned (sqk_host_alloc)",
                     "pageable": e2e_pageable} if e2e_value is not None else None),
            "gpu_launches": int(n_launches),     # counted by the library: every kernel it launched inside the timed region
            "parity": parity,
        }
        if world == 1 and default_shape and not args.no_extras:
            try:
                line["segmenter"] = segmenter_block(ctx, dev, peak, steps=5)
            except Exception as e:
                line["segmenter"] = {"unavailable": f"{type(e).__name__}: {e}"}
            try:
                line["cli_e2e"] = cli_block(dev, motif)
            except Exception as e:
                line["cli_e2e"] = {"unavailable": f"{type(e).__name__}: {e}"}
        print(json.dumps(line), flush=True)
    if h_sig is not None:
        sqk.pinned_free(h_sig)
        sqk.pinned_free(h_hits)
    pipe.close()
    ctx.close()
    ok = torch.tensor([1 if parity_ok else 0], device=dev)
    if world > 1:
        dist.broadcast(ok, 0)
        dist.destroy_process_group()
    if int(ok.item()) != 1:
        sys.stderr.write("bench.py: PARITY GUARD FAILED -- the numbers above are not valid\n")
        return 3
    return 0


if __name__ == "__main__":
    sys.exit(main())
