// sqk_adapter.cuh -- the adapter-stall finder of dRNA_segmenter.py (slow5 branch, dRNA_segmenter.py:86-176) over many
// reads: SURVEY.md section 8(f) row f3, "extra mode of K3".  The threshold (median + sd*0.8 of the post-outlier samples
// [1000, 5000), :104-106) comes from the stats kernel (SQK_STATS_ADAPTER) as the integer bound x <= seg_hi  <=>  x < top.
// The state machine (:108-166) differs from get_segs in four ways, all kept:
//   * the error count is reset when a run opens (:115) and tolerated out-of-range samples only count as errors from
//     position no_err_thresh on (:127-129);
//   * the corrector w is a constant, not a counter (:95, :122, :130);
//   * there is no shorter first segment;
//   * the scan stops once a sample lies more than seg_dist behind the last closed segment (:154-161).
// Only the first segment is reported (:171-174), after any merges that happened before the scan stopped.
// Same mapping as K3: one thread per read, 16-byte L1-cached loads with one block of prefetch, positions in the
// post-outlier index space.
#pragma once
#include "sqk_common.cuh"

#define SQK_ADAPTER_THREADS 128

struct AdapterArgs {
    const int16_t *base;
    int64_t alloc_lo, alloc_hi;
    const int64_t *offsets;
    int64_t read0;
    int n_reads;
    const ReadStats *stats;
    int error, no_err_thresh, corrector, window, seg_dist;
    int32_t *segs;            // [n_reads][2]
    int32_t *found;           // [n_reads]: 1 = segs holds (start, end), 0 = none, -1 = read longer than declared
};

__global__ void __launch_bounds__(SQK_ADAPTER_THREADS) sqk_adapter_fsm_kernel(const AdapterArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_reads) return;
    int64_t alloc_lo = a.alloc_lo, alloc_hi = a.alloc_hi;
    resolve_bounds(a.offsets, a.read0, a.n_reads, alloc_lo, alloc_hi);
    const int64_t r = a.read0 + i;
    const int64_t begin = a.offsets[r], end = a.offsets[r + 1];
    const ReadStats st = a.stats[i];
    a.segs[2 * (int64_t)i] = 0; a.segs[2 * (int64_t)i + 1] = 0;
    if (st.flags & SQK_FLAG_TOO_LONG) { a.found[i] = -1; return; }
    const int seg_hi = st.seg_hi, out_lo = st.out_lo, out_hi = st.out_hi;
    const int w = a.corrector;

    bool open = false, stop = false;
    int err = 0, run_err = 0, c = 0, start = 0, pos = 0, nseg = 0;
    int first_start = 0, first_end = 0, last_end = 0;

    int64_t blk = aligned_block_start(a.base, begin);
    Samples8 cur;
    if (blk < end) cur = load_block8<false>(a.base, blk, alloc_lo, alloc_hi);
    for (; blk < end && !stop; blk += 8) {
        Samples8 nxt;
        if (blk + 8 < end) nxt = load_block8<false>(a.base, blk + 8, alloc_lo, alloc_hi);
        const Samples8 smp = cur;
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const int64_t idx = blk + e;
            const int v = smp.get(e);
            if (stop || idx < begin || idx >= end || v < out_lo || v > out_hi) continue;   // scale_outliers
            if (v <= seg_hi) {
                if (!open) { start = pos; open = true; err = 0; }
                c++;
                run_err = 0;
                if (c >= a.window && c >= w && (c % w) == 0) err--;
            } else if (open && err < a.error) {
                c++;
                if (pos >= a.no_err_thresh) { err++; run_err++; }
                if (c >= a.window && c >= w && (c % w) == 0) err--;
            } else if (open) {
                if (c >= a.window) {
                    const int stop_at = pos - run_err;
                    if (nseg > 0 && start - last_end < a.seg_dist) {
                        last_end = stop_at;
                        if (nseg == 1) first_end = stop_at;
                    } else {
                        nseg++;
                        last_end = stop_at;
                        if (nseg == 1) { first_start = start; first_end = stop_at; }
                    }
                }
                open = false; c = 0; err = 0; run_err = 0;
            } else if (nseg > 0 && pos - last_end > a.seg_dist) {
                stop = true;
            }
            pos++;
        }
        cur = nxt;
    }
    if (nseg > 0) { a.segs[2 * (int64_t)i] = first_start; a.segs[2 * (int64_t)i + 1] = first_end; }
    a.found[i] = nseg > 0 ? 1 : 0;
}
