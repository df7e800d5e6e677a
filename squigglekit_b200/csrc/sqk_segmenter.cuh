// sqk_segmenter.cuh -- K3: segmenter.get_segs (segmenter.py:399-470) over many reads.
//
// The thresholds (median +- stdev*std_scale, segmenter.py:407-414) come from the stats kernel
// (sqk_stats.cuh, SQK_STATS_SEGMENTER) already converted to the equivalent integer window
// seg_lo <= x <= seg_hi.  What is left is the error-tolerant run-length state machine
// (segmenter.py:420-464): inherently sequential per read, integer state.  One thread owns one
// read and walks it with 16-byte loads (each lane streams its own cache line: sector-efficient,
// L1 holds the 32 lines of a warp), dropping outliers on the fly so positions are in the
// post-outlier index space exactly as in the reference; 32 reads advance per warp instruction.
//
// Kept quirks: `w` is never reset between segments; a segment still open at
// the end of the read is not flushed; the first segment may be shorter (window*stall_len).
#pragma once
#include "sqk_common.cuh"

#define SQK_FSM_THREADS 128

struct FsmArgs {
    const int16_t *base;
    int64_t alloc_lo, alloc_hi;
    const int64_t *offsets;
    int64_t read0;
    int n_reads;
    const ReadStats *stats;
    int lo, hi, num;
    int error, corrector, window, seg_dist;
    int first_min;            // ceil(window * stall_len): c >= window*stall_len  <=>  c >= first_min
    int max_segs;
    int32_t *segs;            // [n_reads][max_segs][2]
    int32_t *n_segs;          // [n_reads]
    int code_rows;            // float64 front end: `base` holds 0/1 code rows of n_kept samples (no truncation here)
    const int *list;          // optional work list (launch-local read indices) and its length in device memory:
    const unsigned int *n_list;   //   the reads sqk_stats2_kernel handed back (no bit mask for them)
};

__global__ void __launch_bounds__(SQK_FSM_THREADS) sqk_fsm_kernel(const FsmArgs a)
{
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= (a.list ? (int)*a.n_list : a.n_reads)) return;
    const int i = a.list ? a.list[item] : item;
    int64_t alloc_lo = a.alloc_lo, alloc_hi = a.alloc_hi;
    resolve_bounds(a.offsets, a.read0, a.n_reads, alloc_lo, alloc_hi);
    const int64_t r = a.read0 + i;
    const int64_t begin = a.offsets[r];
    const ReadStats st = a.stats[i];
    const int64_t end = begin + (a.code_rows ? (int64_t)st.n_kept : sqk_truncate_len(a.offsets[r + 1] - begin, a.num));
    if (st.flags & SQK_FLAG_TOO_LONG) { a.n_segs[i] = -1; return; }   // longer than the declared max_read_len
    const int seg_lo = st.seg_lo, seg_hi = st.seg_hi, out_lo = st.out_lo, out_hi = st.out_hi;
    int32_t *out = a.segs + (int64_t)i * a.max_segs * 2;

    bool open = false;
    int err = 0, run_err = 0, c = 0, w = a.corrector;
    int start = 0, pos = 0, nseg = 0;
    int last_start = 0, last_end = 0;

    // Each lane streams its own read 16 bytes at a time through L1 (cached loads: the other half of the 32-byte
    // sector and the rest of the 128-byte line are L1 hits, so HBM traffic stays at the algorithmic bytes); the
    // next block is requested before the current one is consumed.
    int64_t blk = aligned_block_start(a.base, begin);
    Samples8 cur;
    if (blk < end) cur = load_block8<false>(a.base, blk, alloc_lo, alloc_hi);
    for (; blk < end; blk += 8) {
        Samples8 nxt;
        if (blk + 8 < end) nxt = load_block8<false>(a.base, blk + 8, alloc_lo, alloc_hi);
        {
            const Samples8 smp = cur;
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const int64_t idx = blk + e;
                const int v = smp.get(e);
                if (idx < begin || idx >= end || v < out_lo || v > out_hi) continue;   // scale_outliers
                if (v >= seg_lo && v <= seg_hi) {
                    if (!open) { start = pos; open = true; }
                    c++; w++;
                    run_err = 0;
                    if (c >= a.window && c >= w && (c % w) == 0) err--;
                } else if (open) {
                    if (err < a.error) {
                        c++; err++; run_err++;
                        if (c >= a.window && c >= w && (c % w) == 0) err--;
                    } else {
                        if (c >= a.window || (nseg == 0 && c >= a.first_min)) {
                            const int stop = pos - run_err;
                            if (nseg > 0 && start - last_end < a.seg_dist) {
                                last_end = stop;
                            } else {
                                nseg++;
                                last_start = start; last_end = stop;
                            }
                            if (nseg <= a.max_segs) { out[2 * (nseg - 1)] = last_start; out[2 * (nseg - 1) + 1] = last_end; }
                        }
                        open = false; c = 0; err = 0; run_err = 0;
                    }
                }
                pos++;
            }
        }
        cur = nxt;
    }
    a.n_segs[i] = nseg;
}

// ------------------------------------------------------------------------------------------------------------------
// K3, second generation: the same state machine on the in-range BIT MASK that sqk_stats2_kernel emits (one bit per
// post-outlier sample, 1 = bot < x < top).  One thread still owns one read, but it advances run by run (find-first-set
// on the mask word) instead of sample by sample, and it reads 1/16 of the bytes.  Per-sample semantics are kept by
// taking a run in one step only when that is provably the same as stepping through it:
//   * in-range run of L samples: c += L, w += L, run_err = 0.  The `err -= 1` branch (segmenter.py:439-440) needs
//     c >= w; c - w does not change inside such a run, so one test at its start decides whether to step instead.
//   * out-of-range run while a segment is open: k = min(L, error - err) samples are tolerated (c, err, run_err += k)
//     unless c could reach w inside (then: stepping); the next sample closes the segment, the rest of the run is idle.
// ------------------------------------------------------------------------------------------------------------------
struct FsmMaskArgs {
    const uint32_t *mask;     // [n_reads][mask_stride]
    int mask_stride;
    int n_reads;
    const ReadStats *stats;   // n_kept; flags
    int error, corrector, window, seg_dist, first_min, max_segs;
    int32_t *segs;
    int32_t *n_segs;
};

struct FsmState {
    bool open;
    int err, run_err, c, w, start, nseg, last_start, last_end;
};

__device__ __forceinline__ void fsm_close(FsmState &s, int pos, const FsmMaskArgs &a, int32_t *out)
{
    if (s.c >= a.window || (s.nseg == 0 && s.c >= a.first_min)) {
        const int stop = pos - s.run_err;
        if (s.nseg > 0 && s.start - s.last_end < a.seg_dist) {
            s.last_end = stop;
        } else {
            s.nseg++;
            s.last_start = s.start; s.last_end = stop;
        }
        if (s.nseg <= a.max_segs) { out[2 * (s.nseg - 1)] = s.last_start; out[2 * (s.nseg - 1) + 1] = s.last_end; }
    }
    s.open = false; s.c = 0; s.err = 0; s.run_err = 0;
}

__global__ void __launch_bounds__(SQK_FSM_THREADS) sqk_fsm_mask_kernel(const FsmMaskArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_reads) return;
    const ReadStats st = a.stats[i];
    if (st.flags & SQK_FLAG_NO_MASK) return;            // on the redo list: sqk_fsm_kernel handles it
    if (st.flags & SQK_FLAG_TOO_LONG) { a.n_segs[i] = -1; return; }
    const int n = st.n_kept;
    const uint32_t *row = a.mask + (int64_t)i * a.mask_stride;
    int32_t *out = a.segs + (int64_t)i * a.max_segs * 2;

    FsmState s;
    s.open = false; s.err = 0; s.run_err = 0; s.c = 0; s.w = a.corrector; s.start = 0; s.nseg = 0; s.last_start = 0; s.last_end = 0;
    int pos = 0;
    uint32_t cur = n > 0 ? __ldg(row) : 0u;
    while (pos < n) {
        const int sh = pos & 31;
        const uint32_t x = cur >> sh;
        const bool bit = x & 1u;
        int rem = 32 - sh;
        if (rem > n - pos) rem = n - pos;
        const uint32_t y = bit ? ~x : x;
        int L = y ? __ffs((int)y) - 1 : 32;      // samples equal to the current one, within this word
        if (L > rem) L = rem;
        if (bit) {
            if (!s.open) { s.start = pos; s.open = true; }
            s.run_err = 0;
            if (s.c + 1 < s.w + 1) {
                // c < w now and c - w is constant while both count up: the corrector branch cannot fire in this run
                s.c += L; s.w += L;
            } else {
                for (int q = 0; q < L; q++) {
                    s.c++; s.w++;
                    if (s.c >= a.window && s.c >= s.w && (s.c % s.w) == 0) s.err--;
                }
            }
            pos += L;
        } else if (!s.open) {
            pos += L;                               // idle
        } else {
            int used = 0;
            const int room = a.error - s.err;       // samples the segment still tolerates (without corrections)
            if (room > 0 && s.c + (room < L ? room : L) < s.w) {
                const int k = room < L ? room : L;
                s.c += k; s.err += k; s.run_err += k;
                used = k;
            } else {
                while (used < L && s.err < a.error) {
                    s.c++; s.err++; s.run_err++;
                    if (s.c >= a.window && s.c >= s.w && (s.c % s.w) == 0) s.err--;
                    used++;
                }
            }
            if (used < L) {                         // the next sample closes the segment; the rest of the run is idle
                fsm_close(s, pos + used, a, out);
                used = L;
            }
            pos += used;
        }
        if ((pos & 31) == 0 && pos < n) cur = __ldg(row + (pos >> 5));
    }
    a.n_segs[i] = s.nseg;
}
