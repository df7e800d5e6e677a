#!/usr/bin/env python
"""bench.py -- MotifSeq reads/s (4096-sample int16 reads x 80-point motif, subsequence DTW, fp64 exact
mode) on N B200s of one node, plus the HBM / ALU rooflines of the dominant kernel and the CPU path
timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

A step = one pass of the MotifSeq hot path (outlier removal -> z-score -> subsequence DTW -> hit
record) over one batch of 100 000 synthetic reads per GPU (BASELINE.json configs[2]; weak scaling:
every rank gets its own batch, the only exchange is one all-gather of the 16-byte hit records per
step).  `value` is measured with the batch already resident in HBM; `e2e` goes through the C ABI
with pinned HOST buffers, H2D and D2H inside the timed region.  Prints ONE JSON line on rank 0.
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


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons, one sample every 50 ms.  The sampler runs from before the warm-up
    (nvidia-smi needs a moment to start); mark()/unmark() bracket the timed region and summary() only uses
    the samples that arrived inside it."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def __enter__(self):
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

    def mark(self):
        self.t0 = time.perf_counter()

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
            sm, mx, reasons = [], [], set()
            for _, r in rows:
                try:
                    sm.append(float(r[0])); mx.append(float(r[1]))
                except Exception:
                    continue
                for n, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            return sm, mx, reasons

        inside = [x for x in self.rows if self.t0 is not None and self.t0 <= x[0] <= (self.t1 or 1e30)]
        sm, mx, reasons = digest(inside)
        sm_all, mx_all, reasons_all = digest(self.rows)
        if not sm_all:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        use = sm if sm else sm_all
        return {"sm_mhz": statistics.median(use), "sm_max_mhz": max(mx_all), "reasons": sorted(reasons if sm else reasons_all),
                "samples": len(sm), "samples_incl_warmup": len(sm_all), "sm_mhz_min_incl_warmup": min(sm_all),
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
    """--impl reference: the CPU implementation of the path, all host threads, same workload shape."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle
    from squigglekit_b200 import synth
    threads = oracle.max_threads()
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
    sample = (f"{n_used} reads of {N_SAMPLES} samples per step ({N_SAMPLES}x{N_MOTIF} cells each), synthetic, "
              f"{threads} OpenMP threads, full N x M float64 cost matrix + back-trace per read as mlpy does")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"MotifSeq {N_MOTIF}-point motif vs synthetic {N_SAMPLES}-sample int16 reads "
                               f"(BASELINE configs[2] shape), bounded sample per step", "scale": SCALE,
                   "reads_per_step": n_used, "n_samples": N_SAMPLES, "n_motif": N_MOTIF},
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line), flush=True)
    return 0


def alu_peak_cells_per_s(which="dtw_step", prec="fp64"):
    """Register-only micro-benchmark of the DTW step's instruction mix (squigglekit_b200/sqk_ubench):
    the ALU roofline of the kernel.  Falls back to the committed measurement in profiles/."""
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


def main():
    global N_SAMPLES, N_MOTIF, BYTES_PER_READ, CELLS_PER_READ
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--reads", type=int, default=N_READS, help="reads per GPU per step (default: BASELINE configs[2])")
    ap.add_argument("--lanes", type=int, default=0, help="force lanes-per-read of the DTW kernel (experiments)")
    ap.add_argument("--precision", default="fp64", choices=["fp64", "fp32"])
    ap.add_argument("--plan", default="auto", choices=["auto", "single_pass", "two_pass"],
                    help="how exact (fp64) requests run: float64 recurrence over every column, or float32 lower-bound scan + float64 windows (same bits)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (very large batches: it needs the batch in pinned host memory)")
    ap.add_argument("--samples", type=int, default=4096, help="samples per read (default 4096; 20000 = BASELINE configs[3])")
    ap.add_argument("--motif-len", type=int, default=80, help="motif points (default 80; 163 = the reference's example model)")
    args = ap.parse_args()
    N_SAMPLES, N_MOTIF = args.samples, args.motif_len
    BYTES_PER_READ = 2 * N_SAMPLES + 16
    CELLS_PER_READ = N_SAMPLES * N_MOTIF
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import squigglekit_b200 as sqk
    from squigglekit_b200 import synth
    from squigglekit_b200.dist import GatherPipeline, env_rank_world

    rank, world, local_rank = env_rank_world()
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)

    R, M = args.reads, N_SAMPLES
    motif = bench_motif()
    ctx = sqk.Context(local_rank)
    if args.lanes:
        ctx.set_dtw_lanes(args.lanes)
    ctx.set_dtw_plan(args.plan)
    sig = synth.motifseq_reads_torch(R, M, motif, dev, seed=synth.BASE_SEED + rank).view(-1)
    off = torch.arange(R + 1, dtype=torch.int64, device=dev) * M
    # one all-gather of the 16-byte hit records per step, double-buffered: the gather of step i overlaps the kernels
    # of step i+1, so the ranks are not forced into lockstep at every step (everything is drained inside the timed region)
    pipe = GatherPipeline((R, 1, 16), torch.uint8, dev, depth=2)
    hits = pipe.local[0]

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
    with ClockSampler(local_rank) as clocks:
        for _ in range(args.warmup):
            gathered = step()
        pipe.drain()
        barrier()
        ctx.enable_timing(True)
        ctx.timing(reset=True)
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
    ctx.enable_timing(False)
    plan_counters = ctx.plan_counters()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * R * args.steps / (ms_max * 1e-3)

    # ---- parity guard on what was just timed: a fixed sub-sample against the CPU oracle -------------
    parity = None
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
        else:
            parity = {"reads_checked": int(idx.size),
                      "index_mismatch_rate": float(np.mean((got["start"] != want["start"]) | (got["end"] != want["end"]))),
                      "dist_max_rel_err": float(np.max(np.abs(got["dist"] - want["dist"]) / np.maximum(want["dist"], 1e-9)))}
        if world > 1:
            assert gathered.shape[0] == world * R

    e2e_value, e2e_steps, h_sig, h_hits = None, 0, None, None
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

    if rank == 0:
        peak, peak_src = measured_peaks()
        two_pass = kt["dtw_lb"]["launches"] > 0
        if two_pass:
            # dominant kernel: the float32 lower-bound scan (every sample of every read goes through it once)
            kname, kkey, ub, ubp = "sqk_dtw_lb_kernel", "dtw_lb", "lb_step|lb_step2", "fp32_rd"
        else:
            kname, kkey, ub, ubp = "sqk_dtw_kernel", "dtw", "dtw_step", "fp64" if args.precision == "fp64" else "fp32"
        dtw_launches = max(1, kt[kkey]["launches"])
        dtw_ms = kt[kkey]["ms"] / dtw_launches
        achieved = R * BYTES_PER_READ / (dtw_ms * 1e-3) / 1e9
        alu_peak, alu_src = alu_peak_cells_per_s(ub, ubp)
        cells_s = R * CELLS_PER_READ / (dtw_ms * 1e-3)
        traffic = None      # ncu dram__bytes_read + dram__bytes_write of one launch: only known for the captured shape
        if R == 100000 and M == 4096 and N_MOTIF == 80:
            try:
                with open(os.path.join(ROOT, "profiles", "dtw_traffic.json")) as f:
                    traffic = json.load(f).get("dram_bytes_per_launch_100k_reads_lb" if two_pass else "dram_bytes_per_launch_100k_reads")
            except Exception:
                pass
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            sub_n = 8192
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
                                   f"(BASELINE configs[2] when 100000 x 4096 x 80); zscale; outlier window (0,1200)",
                       "reads_per_gpu": R, "n_samples": M, "n_motif": N_MOTIF, "scale": SCALE, "precision": args.precision,
                       "l2_policy": f"input {R * M * 2 / 1e6:.0f} MB per step > 126 MB L2 (no flush needed)",
                       "timing": "CUDA events on the launching stream around K steps, max over ranks; e2e = wall clock of the synchronous host-buffer C-ABI call",
                       "exchange": "one all-gather of 16-byte hit records per step, double-buffered (overlaps the next step's kernels; drained inside the timed region)" if world > 1 else "none (1 GPU)",
                       "dtw_lanes_per_read": args.lanes or "auto", "dtw_plan": args.plan},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": kname,
                         "kernel_ms_per_launch": dtw_ms, "algorithmic_bytes_per_read": BYTES_PER_READ,
                         "note": "the DTW recurrence is ALU-issue bound (SURVEY F7): see roofline_alu",
                         "stats_kernel_ms_per_launch": kt["stats"]["ms"] / max(1, kt["stats"]["launches"]),
                         "exact_windows_ms_per_step": (kt["dtw_win"]["ms"] / max(1, kt["dtw_win"]["launches"])) if two_pass else None},
            "plan": {"name": "two_pass" if two_pass else "single_pass",
                     "exact_windows_per_step": plan_counters["windows"] if two_pass else None,
                     "full_length_fallback_reads_per_step": plan_counters["fallback_reads"] if two_pass else None},
            "roofline_alu": {"achieved_cells_per_s": cells_s, "peak_cells_per_s": alu_peak,
                             "frac": (cells_s / alu_peak) if alu_peak else None, "peak_source": alu_src},
            "cpu_baseline": cpu,
            "clocks": clocks.summary(),
            "e2e": ({"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": int(R * M * 2 + (R + 1) * 8),
                     "d2h_bytes_per_step": int(R * 16), "steps": e2e_steps} if e2e_value is not None else None),
            # stats + DTW kernels timed by the library; the two-pass plan launches 3 kernels (windows, finalize, fallback) behind its "dtw_win" timer
            "gpu_launches": int(kt["dtw"]["launches"] + kt["stats"]["launches"] + kt["dtw_lb"]["launches"] + 3 * kt["dtw_win"]["launches"]),
            "parity": parity,
        }
        print(json.dumps(line), flush=True)
    if h_sig is not None:
        sqk.pinned_free(h_sig)
        sqk.pinned_free(h_hits)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
