#!/usr/bin/env python
"""Secondary benchmark (BASELINE configs[1]): segmenter stall + homopolymer detection on synthetic
4096-sample int16 reads, 1 B200.  Prints one JSON line per batch size with reads/s (device-resident and
end-to-end through the host-buffer C ABI), per-kernel device times, achieved GB/s against the measured HBM
peak, and the CPU oracle port on one host core.  Not the driver's bench (that is bench.py); used for
DESIGN.md / profiles.

    python tools/bench_segmenter.py [--reads 10000 1000000] [--steps 10]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

M = 4096
MAX_SEGS = 16
BYTES_PER_READ = 2 * M + 4 * (1 + 2 * MAX_SEGS)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, nargs="+", default=[10_000, 1_000_000])
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pa", action="store_true", help="pA mode (per-read calibration) instead of raw")
    ap.add_argument("--no-e2e", action="store_true", help="device-resident steps only (profiling runs)")
    args = ap.parse_args()

    import torch

    import oracle
    import squigglekit_b200 as sqk
    from squigglekit_b200 import synth

    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json)"
    except Exception:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    dev = torch.device("cuda", 0)
    ctx = sqk.Context(0)
    cfg = sqk.SegConfig(stall=True, max_segs=MAX_SEGS)          # -ku: stall detection + test_segs on the host
    for R in args.reads:
        sig = synth.segmenter_reads_torch(R, M, dev).view(-1)
        off = torch.arange(R + 1, dtype=torch.int64, device=dev) * M
        kw = {}
        if args.pa:
            g = torch.Generator(device=dev); g.manual_seed(3)
            kw = dict(pa_offset=torch.randint(-30, 40, (R,), generator=g, device=dev).double(),
                      pa_scale=(torch.rand(R, generator=g, device=dev, dtype=torch.float64) * 500 + 1100).mul(100).round().div(100) / 8192.0)
        for _ in range(args.warmup):
            segs, nsegs = ctx.segmenter(sig, off, cfg, max_read_len=M, **kw)
        torch.cuda.synchronize()
        ctx.enable_timing(True); ctx.timing(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            segs, nsegs = ctx.segmenter(sig, off, cfg, max_read_len=M, **kw)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        kt = ctx.timing(reset=True); ctx.enable_timing(False)
        # parity on a sub-sample
        idx = np.arange(0, R, max(1, R // 512))[:512]
        sub = sig.view(R, M)[torch.from_numpy(idx).to(dev)].cpu().numpy().reshape(-1)
        suboff = np.arange(idx.size + 1, dtype=np.int64) * M
        if args.pa:
            want, wn = oracle.segmenter_batch_pa(sub, suboff, kw["pa_offset"].cpu().numpy()[idx], kw["pa_scale"].cpu().numpy()[idx],
                                                 oracle.SegCfg(), 0, 900, 0, MAX_SEGS)
        else:
            want, wn = oracle.segmenter_batch(sub, suboff, oracle.SegCfg(), 0, 900, 0, MAX_SEGS)
        got, gn = segs.cpu().numpy()[idx], nsegs.cpu().numpy()[idx]
        mask = np.arange(MAX_SEGS)[None, :, None] < wn[:, None, None]
        parity = bool(np.array_equal(gn, wn) and np.array_equal(np.where(mask, got, 0), np.where(mask, want, 0)))
        if args.no_e2e:
            continue
        # e2e: host buffers
        h_sig = sqk.pinned_empty(R * M, np.int16); h_sig[:] = sig.cpu().numpy()
        h_off = off.cpu().numpy()
        hkw = {k: v.cpu().numpy() for k, v in kw.items()}
        for _ in range(2):
            ctx.segmenter(h_sig, h_off, cfg, max_read_len=M, **hkw)
        t0 = time.perf_counter()
        n_e2e = max(2, min(5, args.steps))
        for _ in range(n_e2e):
            hs, hn = ctx.segmenter(h_sig, h_off, cfg, max_read_len=M, **hkw)
            kept = [sqk.test_segs(s, cfg) for s in sqk.segs_to_lists(hs[:64], hn[:64])]   # host-side -u filter (sampled)
        e2e_s = (time.perf_counter() - t0) / n_e2e
        sqk.pinned_free(h_sig)
        # CPU: oracle port, one core
        ncpu = min(R, 2000)
        t0 = time.perf_counter()
        oracle.segmenter_batch(sub[: min(idx.size, 512) * M], suboff[: min(idx.size, 512) + 1], oracle.SegCfg(), 0, 900, 0, MAX_SEGS, n_threads=1)
        cpu_rate = min(idx.size, 512) / (time.perf_counter() - t0)
        stats_ms = kt["stats"]["ms"] / max(1, kt["stats"]["launches"])
        fsm_ms = kt["seg_fsm"]["ms"] / max(1, kt["seg_fsm"]["launches"])      # 0 when get_segs ran fused in the stats kernel
        print(json.dumps({
            "metric": "segmenter reads/sec (4096-sample int16 reads, get_segs -ku)", "mode": "pA" if args.pa else "raw",
            "reads": R, "value": R / (ms * 1e-3), "unit": "reads/s", "ms_per_step": ms,
            "kernels_ms": {"stats (sqk_stats3_kernel + redo list; pA: sqk_stats2_kernel)": stats_ms, "state machine (sqk_fsm_mask_kernel + redo list)": fsm_ms},
            "roofline": {"bound": "hbm", "unit": "GB/s", "peak": peak, "peak_source": peak_src,
                         "achieved_step": R * BYTES_PER_READ / (ms * 1e-3) / 1e9,
                         "frac_step": R * BYTES_PER_READ / (ms * 1e-3) / 1e9 / peak,
                         "achieved_stats_kernel": R * 2 * M / (stats_ms * 1e-3) / 1e9,
                         "achieved_fsm_kernel": (R * BYTES_PER_READ / (fsm_ms * 1e-3) / 1e9) if fsm_ms > 0 else None,
                         "algorithmic_bytes_per_read": BYTES_PER_READ},
            "e2e": {"value": R / e2e_s, "unit": "reads/s", "h2d_bytes_per_step": R * M * 2 + (R + 1) * 8,
                    "d2h_bytes_per_step": R * (MAX_SEGS * 8 + 4)},
            "cpu_baseline": {"value": cpu_rate, "unit": "reads/s", "cores": 1, "kind": "port",
                             "sample": f"{min(idx.size, 512)} reads, C restatement of get_segs (the reference's pure-Python loop is ~640 reads/s, SURVEY §6)"},
            "parity_subsample_bit_exact": parity,
            "segments_found_mean": float(nsegs.float().mean().item()),
        }), flush=True)
        del sig, off, segs, nsegs
        torch.cuda.empty_cache()
    ctx.close()


if __name__ == "__main__":
    main()
