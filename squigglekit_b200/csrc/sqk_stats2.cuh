// sqk_stats2.cuh -- K1, second generation: the per-read statistics of sqk_stats.cuh (same modes, same bits) for reads
// that fit a shared-memory window, restructured around what the first version spent its time on (profiles/r01_stats_*):
//
//   * no compaction pass.  The raw read is brought into shared memory by the copy engine: 1-D bulk copies
//     (cp.async.bulk, SASS UBLKCP) signalled through an mbarrier, double-buffered so the next read lands while this
//     one is being reduced.  Outliers (scale_outliers, MotifSeq.py:317-324 / segmenter.py:311-318) are rare, so they
//     are kept as a short sorted exception list: kept sample c sits at raw position c + #{outliers in front of it}.
//     Reads with more than SQK_S2_MAXOUT outliers go to a redo list that the first-generation kernel works off.
//   * the window test and the integer sum run on packed int16 pairs (VIMNMX.S16x2 clamp + xor, IDP.2A), 16 bytes per
//     shared-memory load; only 16-byte units that hold an outlier or a read edge are looked at sample by sample.
//   * the pairwise tree (numpy's order, sqk_stats_plan.cuh) reads its leaves straight from the raw copy: a leaf without
//     an outlier inside is a contiguous run at a constant shift.  The copy is laid out in 256-byte segments padded by
//     16 bytes, so the four 8-lane teams of a warp, which work on neighbouring leaves (~256 bytes apart), hit different
//     banks (the first version had a 4-way conflict on every leaf load).
//   * medians come from one shared-memory histogram filled in the same pass as the window test; selection is done by
//     the thread whose bin run contains the rank (no serial binary search); the MAD comes from the histogram folded
//     around the median (no second pass over the samples).
//   * segmenter mode also emits the in-range bit mask of the post-outlier signal (1 bit per kept sample), which is all
//     the get_segs state machine needs: K3 (sqk_fsm_mask_kernel) then walks runs of bits instead of samples and reads
//     1/16 of the bytes.
#pragma once
#include "sqk_stats.cuh"

#define SQK_S2_THREADS 128
#define SQK_S2_MAXOUT 64            // outliers per read the exception list holds
#define SQK_S2_MAX_LEN 8176         // longest read (samples) this kernel stages; longer reads use sqk_stats_kernel
#define SQK_S2_LEAF_SLOTS 128       // depth <= 7 for n <= 8192
#define SQK_S2_MAX_BINS 2048        // histogram capacity (outlier window span)

struct Stats2Args {
    StatsArgs s;              // same meaning as for sqk_stats_kernel (cap / gstage unused)
    int buf_bytes;            // bytes of one staging buffer (multiple of 16)
    int hist_words;           // words reserved for the histogram(s): 0, SQK_S2_MAX_BINS or 2 * SQK_S2_MAX_BINS
    int mask_words;           // words of the raw-space mask scratch in shared memory (segmenter mode)
    uint32_t *mask;           // segmenter mode: [n_reads][mask_stride] in-range bits of the kept samples, or null
    int mask_stride;          // words per read
    int *redo;                // reads this kernel hands to sqk_stats_kernel (launch-local indices)
    unsigned int *n_redo;
};

struct S2Shared {
    unsigned long long bar[2];
    double leaf[SQK_S2_LEAF_SLOTS];
    uint32_t leaf_desc[SQK_S2_LEAF_SLOTS];   // per slot: offset << 17 | length << 9 | outliers in front << 2 | clean << 1 | has-a-leaf
    double tree_out;
    int sum_part[4];
    uint32_t scan_part[2][4];
    int sel[4];
    int out_cnt;
    int out_pos[SQK_S2_MAXOUT];     // raw positions (relative to the read's first sample), unsorted
    int out_adj[SQK_S2_MAXOUT];     // sorted, minus rank: outlier i sits in front of kept sample out_adj[i]
};

__device__ __forceinline__ unsigned s2_smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void s2_mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void s2_mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void s2_mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// global -> shared bulk copy (16-byte aligned, multiple of 16 bytes), completion counted on the mbarrier
__device__ __forceinline__ void s2_bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ int s2_clamp2(int w, int lo2, int hi2)
{
    int c;
    asm("{.reg .b32 t; max.s16x2 t, %1, %2; min.s16x2 %0, t, %3;}" : "=r"(c) : "r"(w), "r"(lo2), "r"(hi2));
    return c;
}

// byte offset of buffer sample B / of 16-byte unit u in the staging buffer.  (A layout padded by 16 bytes per 256 to keep
// the four teams of a warp off each other's banks was tried: the 32 bulk copies per read it needs cost more issue slots
// than the 4-way conflicts cost LSU cycles -- the kernel is issue-bound, not LSU-bound.)
__device__ __forceinline__ unsigned s2_sample_off(int B) { return 2u * (unsigned)B; }
__device__ __forceinline__ unsigned s2_unit_off(int u) { return 16u * (unsigned)u; }

__device__ __forceinline__ int s2_lds_s16(unsigned addr)
{
    int v;
    asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ int4 s2_lds128(unsigned addr)
{
    int4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void s2_sts128(unsigned addr, int4 v)
{
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// One read as the kernel sees it: `len` samples starting h0 samples into the staged 16-byte hull.
struct S2Read {
    int64_t begin;
    int len, h0, units;
    bool tma;         // staged by bulk copy (else: by the threads, for hulls that stick out of the allocation)
    bool ok;          // fits the staging buffer
};

__device__ __forceinline__ S2Read s2_describe(const StatsArgs &a, int64_t i, int64_t alloc_lo, int64_t alloc_hi, int buf_bytes)
{
    S2Read rd;
    const int64_t r = a.read0 + i;
    rd.begin = a.offsets[r];
    int64_t len = a.offsets[r + 1] - rd.begin;
    if (a.mode == SQK_STATS_SEGMENTER) len = sqk_truncate_len(len, a.num);
    const int64_t blk0 = aligned_block_start(a.base, rd.begin);
    rd.h0 = (int)(rd.begin - blk0);
    const int64_t units = (rd.h0 + len + 7) >> 3;
    rd.ok = units * 16 <= (int64_t)buf_bytes;
    rd.len = rd.ok ? (int)len : 0;
    rd.units = rd.ok ? (int)units : 0;
    rd.tma = blk0 >= alloc_lo && blk0 + units * 8 <= alloc_hi;
    return rd;
}

// Start staging a read into buffer `bufs` (shared address).  Called by every thread of the CTA.
__device__ __forceinline__ void s2_stage(const StatsArgs &a, const S2Read &rd, unsigned bufs, unsigned bar, int64_t alloc_lo,
                                         int64_t alloc_hi)
{
    if (rd.units == 0) return;
    const int tid = threadIdx.x;
    const int16_t *src = a.base + (rd.begin - rd.h0);
    if (rd.tma) {
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic-proxy accesses to the buffer
            s2_mbar_expect_tx(bar, (unsigned)rd.units * 16u);
            s2_bulk_g2s(bufs, src, (unsigned)rd.units * 16u, bar);
        }
    } else {
        for (int u = tid; u < rd.units; u += SQK_S2_THREADS) {
            const Samples8 sv = load_block8(a.base, rd.begin - rd.h0 + 8ll * u, alloc_lo, alloc_hi);
            s2_sts128(bufs + s2_unit_off(u), sv.v);
        }
    }
}

// number of outliers in front of (or at) kept sample c:  #{i : out_adj[i] <= c}
__device__ __forceinline__ int s2_shift(const S2Shared &sh, int n_out, int c)
{
    int s = 0;
    for (int i = 0; i < n_out; i++) s += (sh.out_adj[i] <= c) ? 1 : 0;
    return s;
}

// fold of the eight accumulators of a team, numpy's order ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)); the mask names the team
// only, so teams of one warp may sit in different branches
__device__ __forceinline__ double s2_team_fold(double r, unsigned mask)
{
#pragma unroll
    for (int m = 1; m < 8; m <<= 1) {
        int lo = __double2loint(r), hi = __double2hiint(r);
        lo = __shfl_xor_sync(mask, lo, m, 8);
        hi = __shfl_xor_sync(mask, hi, m, 8);
        r = __dadd_rn(r, __hiloint2double(hi, lo));
    }
    return r;
}

// numpy pairwise leaf (sqk_stats.cuh: stats_leaf_sum) over term(raw value) of kept samples [off, off+len), read from the
// raw copy.  k = accumulator (lane of the 8-lane team); s0 = outliers in front of the leaf; clean = no outlier inside and
// a whole number of 8-sample rows.
template <class Term>
__device__ __forceinline__ double s2_leaf_sum(Term term, unsigned bufs, int h0, const S2Shared &sh, int n_out, int off, int len,
                                              int s0, bool clean, int k)
{
    const unsigned tmask = 0xffu << (threadIdx.x & 24);
    double r;
    if (clean) {
        // a contiguous run of the raw copy: rows of 8 samples, 16 bytes apart
        unsigned addr = bufs + s2_sample_off(h0 + off + s0 + k);
        const int rows = len >> 3;
        r = term(s2_lds_s16(addr));
        for (int i = 1; i < rows; i++) { addr += 16; r = __dadd_rn(r, term(s2_lds_s16(addr))); }
        return s2_team_fold(r, tmask);
    }
    // general leaf (an outlier inside, or the ragged last leaf): kept sample c sits at raw position c + (outliers in
    // front of it); c only grows, so the shift is advanced along the sorted exception list
    int oi = s0;                                    // out_adj[0 .. oi) are <= off
    int nxt = oi < n_out ? sh.out_adj[oi] : 0x7fffffff;
    auto at = [&](int c) -> double {
        while (c >= nxt) { oi++; nxt = oi < n_out ? sh.out_adj[oi] : 0x7fffffff; }
        return term(s2_lds_s16(bufs + s2_sample_off(h0 + c + oi)));
    };
    if (len < 8) {
        r = 0.0;
        for (int i = 0; i < len; i++) r = __dadd_rn(r, at(off + i));
        return r;
    }
    const int body = len - (len & 7);
    r = at(off + k);
    for (int i = 8; i < body; i += 8) r = __dadd_rn(r, at(off + i + k));
    r = s2_team_fold(r, tmask);
    for (int i = body; i < len; i++) r = __dadd_rn(r, at(off + i));   // (off + body > every c seen so far: still monotone)
    return r;
}

// Slot table of the pairwise tree over n elements (sqk_stats_plan.cuh): thread j describes slot j -- which kept samples
// its leaf covers, how many outliers sit in front of it and whether it is a clean run of the raw copy.  Visible after
// the next barrier.  (out_adj must be visible: call after the barrier that publishes the exception list.)
__device__ __forceinline__ void s2_tree_describe(S2Shared &sh, int n, int n_out)
{
    const int j = threadIdx.x;
    const int depth = sqk_tree_depth(n);
    if (j < (1 << depth)) {
        int off = 0, len = 0;
        const bool mine = sqk_tree_leaf(n, depth, j, &off, &len);
        int s0 = 0, s1 = 0;
        for (int i = 0; i < n_out; i++) {
            const int adj = sh.out_adj[i];
            s0 += (adj <= off) ? 1 : 0;
            s1 += (adj <= off + len - 1) ? 1 : 0;
        }
        const bool clean = s0 == s1 && len >= 8 && (len & 7) == 0;
        sh.leaf_desc[j] = ((uint32_t)off << 17) | ((uint32_t)len << 9) | ((uint32_t)s0 << 2) | (clean ? 2u : 0u) | (mine ? 1u : 0u);
    }
}

// np.sum of term over the kept samples [0, n) in numpy's pairwise order: leaf sums by the 16 teams into sh.leaf (after
// s2_tree_describe + barrier).  The caller puts a barrier behind it and lets warp 0 fold (s2_tree_fold).
template <class Term>
__device__ __forceinline__ void s2_tree_leaves(Term term, unsigned bufs, int h0, S2Shared &sh, int n_out, int n)
{
    const int tid = threadIdx.x, team = tid >> 3, k = tid & 7;
    const int slots = 1 << sqk_tree_depth(n);
    for (int j = team; j < slots; j += SQK_S2_THREADS / 8) {
        const uint32_t d = sh.leaf_desc[j];
        double v = 0.0;
        if (d & 1u)                                                       // (team-uniform)
            v = s2_leaf_sum(term, bufs, h0, sh, n_out, (int)(d >> 17), (int)((d >> 9) & 0xffu), (int)((d >> 2) & 0x7fu), (d & 2u) != 0, k);
        if (k == 0) sh.leaf[j] = v;
    }
}

// warp 0, after a barrier: fold the leaf slots in slot order (numpy's order).  Result valid in every lane of warp 0.
__device__ __forceinline__ double s2_tree_fold(const S2Shared &sh, int n)
{
    const int lane = threadIdx.x & 31;
    const int slots = 1 << sqk_tree_depth(n);
    const int per = slots > 32 ? slots / 32 : 1;   // 1, 2 or 4
    double r;
    if (per == 1) r = lane < slots ? sh.leaf[lane] : 0.0;
    else if (per == 2) r = __dadd_rn(sh.leaf[2 * lane], sh.leaf[2 * lane + 1]);
    else r = __dadd_rn(__dadd_rn(sh.leaf[4 * lane], sh.leaf[4 * lane + 1]), __dadd_rn(sh.leaf[4 * lane + 2], sh.leaf[4 * lane + 3]));
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) r = __dadd_rn(r, shfl_xor_f64(r, m, 32));
    return r;
}

// Block-wide selection on a histogram of `per * 128` bins (per a power of two <= 16): the threads whose run of `per`
// consecutive bins contains rank0 / rank1 (0-based order statistics) write the bin index to sel[0] / sel[1]; the
// histogram is zeroed on the way (ready for the next read).  phase = which scan_part row to use (callers alternate).
// Contains one barrier; the results are visible after the NEXT barrier.
template <bool ZERO>
__device__ __forceinline__ void s2_hist_select(uint32_t *hist, int per, int rank0, int rank1, S2Shared &sh, int phase)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t v[16];
    uint32_t run = 0;
    if (per >= 4) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (q * 4 < per) {
                const uint4 w = *reinterpret_cast<const uint4 *>(hist + tid * per + q * 4);
                v[q * 4] = w.x; v[q * 4 + 1] = w.y; v[q * 4 + 2] = w.z; v[q * 4 + 3] = w.w;
                if (ZERO) *reinterpret_cast<uint4 *>(hist + tid * per + q * 4) = make_uint4(0, 0, 0, 0);
            } else {
                v[q * 4] = v[q * 4 + 1] = v[q * 4 + 2] = v[q * 4 + 3] = 0;
            }
        }
    } else {
#pragma unroll
        for (int q = 0; q < 16; q++) {
            v[q] = 0;
            if (q < per) { v[q] = hist[tid * per + q]; if (ZERO) hist[tid * per + q] = 0; }
        }
    }
#pragma unroll
    for (int q = 0; q < 16; q++) run += v[q];
    uint32_t incl = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(SQK_FULL_MASK, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) sh.scan_part[phase][warp] = incl;
    __syncthreads();
    uint32_t excl = incl - run;
    for (int w = 0; w < warp; w++) excl += sh.scan_part[phase][w];
    const uint32_t ranks[2] = {(uint32_t)rank0, (uint32_t)rank1};
#pragma unroll
    for (int t = 0; t < 2; t++) {
        if (ranks[t] >= excl && ranks[t] < excl + run) {
            uint32_t cum = excl;
            int bin = 0;
#pragma unroll
            for (int q = 0; q < 16; q++) {
                if (ranks[t] >= cum + v[q]) { cum += v[q]; bin = q + 1; }
                else break;
            }
            sh.sel[t] = tid * per + bin;
        }
    }
}

// MODE / PA are compile-time so that every launch carries only its own code (the all-modes kernel was ~9000
// instructions and stalled on instruction fetch).
template <int MODE, bool PA>
__global__ void __launch_bounds__(SQK_S2_THREADS, 6) sqk_stats2_kernel(const Stats2Args A)
{
    const StatsArgs &a = A.s;
    extern __shared__ __align__(16) unsigned char s2_smem[];
    S2Shared &sh = *reinterpret_cast<S2Shared *>(s2_smem);
    constexpr unsigned FIXED = (sizeof(S2Shared) + 15) & ~15u;
    uint32_t *hist = reinterpret_cast<uint32_t *>(s2_smem + FIXED);
    uint32_t *fold = hist + SQK_S2_MAX_BINS;                                    // medmad only
    uint32_t *rawmask = reinterpret_cast<uint32_t *>(s2_smem + FIXED + 4u * (unsigned)A.hist_words);
    const unsigned buf0 = s2_smem_addr(s2_smem + FIXED + 4u * (unsigned)(A.hist_words + A.mask_words));
    const unsigned bar0 = s2_smem_addr(&sh.bar[0]);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int64_t alloc_lo = a.alloc_lo, alloc_hi = a.alloc_hi;
    resolve_bounds(a.offsets, a.read0, a.n_reads, alloc_lo, alloc_hi);
    constexpr int mode = MODE;
    constexpr bool pa_mode = PA;
    constexpr bool want_hist = (mode == SQK_STATS_MEDMAD || mode == SQK_STATS_SEGMENTER);

    if (tid == 0) {
        s2_mbar_init(bar0, 1);
        s2_mbar_init(bar0 + 8, 1);
        sh.out_cnt = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int b = tid; b < A.hist_words; b += SQK_S2_THREADS) hist[b] = 0;
    __syncthreads();

    const int64_t stride = gridDim.x;
    int64_t i = blockIdx.x;
    unsigned uses0 = 0, uses1 = 0;         // bulk-staged reads each buffer has held (mbarrier phase)
    S2Read rd{};
    if (i < a.n_reads) {
        rd = s2_describe(a, i, alloc_lo, alloc_hi, A.buf_bytes);
        s2_stage(a, rd, buf0, bar0, alloc_lo, alloc_hi);
    }
    int it = 0;
    for (; i < a.n_reads; i += stride, it++) {
        const int b = it & 1;
        const unsigned bufs = buf0 + (unsigned)b * (unsigned)A.buf_bytes;
        const unsigned bar = bar0 + 8u * (unsigned)b;
        const int64_t inext = i + stride;
        S2Read nx{};
        if (inext < a.n_reads) nx = s2_describe(a, inext, alloc_lo, alloc_hi, A.buf_bytes);

        // ---- outlier window on the raw sample (inclusive) ------------------------------------
        int out_lo = a.lo + 1, out_hi = a.hi - 1;
        double pa_off = 0.0, pa_unit = 1.0;
        if constexpr (pa_mode) {
            pa_off = a.pa_offset[a.read0 + i]; pa_unit = a.pa_scale[a.read0 + i];
            if (pa_unit > 0.0) {
                out_lo = stats_first_above((double)a.lo, pa_off, pa_unit);
                out_hi = stats_last_below((double)a.hi, pa_off, pa_unit);
            } else {
                out_lo = 1; out_hi = 0;
            }
        }
        // the packed test works on int16 lanes: clip the window to the int16 range (samples cannot be outside it)
        const int wlo = out_lo < -32768 ? -32768 : out_lo, whi = out_hi > 32767 ? 32767 : out_hi;
        const bool window_ok = whi >= wlo;
        const int nbins = window_ok ? whi - wlo + 1 : 0;
        const bool hist_ok = !want_hist || nbins <= SQK_S2_MAX_BINS;
        const int lo2 = (wlo & 0xffff) | (wlo << 16), hi2 = (whi & 0xffff) | (whi << 16);
        const unsigned span = (unsigned)(whi - wlo);

        // ---- wait for the staged read ---------------------------------------------------------
        if (rd.units > 0 && rd.tma) {
            s2_mbar_wait(bar, (b ? uses1 : uses0) & 1u);
            if (b) uses1++; else uses0++;
        }
        // (thread-staged reads became visible at the barrier that closed the previous iteration)

        const bool punt = !rd.ok || !window_ok || !hist_ok;     // not for this kernel: the redo list takes it
        // ---- phase A: window test, integer sum, histogram, outlier list ---------------------------
        int lsum = 0;
        if (!punt) {
            const int h0 = rd.h0, len = rd.len;
            for (int u = tid; u < rd.units; u += SQK_S2_THREADS) {
                const int4 q = s2_lds128(bufs + s2_unit_off(u));
                const int w[4] = {q.x, q.y, q.z, q.w};
                int d[4];
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    d[e] = s2_clamp2(w[e], lo2, hi2) ^ w[e];          // != 0 in the halves that hold an outlier
                    lsum = __dp2a_lo(w[e], 0x0101, lsum);
                }
                const bool edge = (8 * u < h0) || (8 * u + 8 > h0 + len);
                if (!edge) {
                    if (want_hist) {
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            if (!(d[e] & 0xffff)) atomicAdd(&hist[(int)(short)(w[e] & 0xffff) - wlo], 1u);
                            if (!(d[e] & 0xffff0000)) atomicAdd(&hist[(w[e] >> 16) - wlo], 1u);
                        }
                    }
                    if ((d[0] | d[1]) | (d[2] | d[3])) {
                        // rare: record the outliers of this unit (only the flagged halves are looked at)
#pragma unroll
                        for (int e = 0; e < 8; e++) {
                            const int dd = (e & 1) ? (d[e >> 1] & 0xffff0000) : (d[e >> 1] & 0xffff);
                            if (dd) {
                                lsum -= (e & 1) ? (w[e >> 1] >> 16) : (int)(short)(w[e >> 1] & 0xffff);
                                const int at = atomicAdd(&sh.out_cnt, 1);
                                if (at < SQK_S2_MAXOUT) sh.out_pos[at] = 8 * u + e - h0;
                            }
                        }
                    }
                } else {
                    // first / last unit of a read that does not start or end on a 16-byte boundary: sample by sample
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        const int v = (e & 1) ? (w[e >> 1] >> 16) : (int)(short)(w[e >> 1] & 0xffff);
                        const int pos = 8 * u + e - h0;
                        const bool inside = pos >= 0 && pos < len;
                        const bool keep = inside && (unsigned)(v - wlo) <= span;
                        if (!keep) lsum -= v;
                        if (inside && !keep) {
                            const int at = atomicAdd(&sh.out_cnt, 1);
                            if (at < SQK_S2_MAXOUT) sh.out_pos[at] = pos;
                        }
                        if (keep && want_hist) atomicAdd(&hist[v - wlo], 1u);
                    }
                }
            }
        }
        lsum = __reduce_add_sync(SQK_FULL_MASK, lsum);
        if (lane == 0) sh.sum_part[warp] = lsum;
        __syncthreads();                                                    // B1
        // every thread is past the previous read: its buffer may be refilled
        if (inext < a.n_reads) s2_stage(a, nx, buf0 + (unsigned)(b ^ 1) * (unsigned)A.buf_bytes, bar0 + 8u * (unsigned)(b ^ 1), alloc_lo, alloc_hi);

        const int n_out_all = sh.out_cnt;
        const bool redo = punt || n_out_all > SQK_S2_MAXOUT;
        const int n_out = redo ? 0 : n_out_all;
        const int n = redo ? 0 : rd.len - n_out;
        const long long tot_sum = (long long)sh.sum_part[0] + sh.sum_part[1] + sh.sum_part[2] + sh.sum_part[3];
        const int h0 = rd.h0;

        if (n_out > 0) {
            if (warp == 0) {
                for (int q = lane; q < n_out; q += 32) {
                    const int p = sh.out_pos[q];
                    int rank = 0;
                    for (int j = 0; j < n_out; j++) rank += (sh.out_pos[j] < p) ? 1 : 0;
                    sh.out_adj[rank] = p - rank;
                }
            }
            __syncthreads();                                                // (uniform condition) exception list visible
        }
        const bool want_sd = n > 0 && (mode == SQK_STATS_ZSCALE || mode == SQK_STATS_SEGMENTER);
        if (want_sd) s2_tree_describe(sh, n, n_out);
        __syncthreads();                                                    // B2: slot table visible

        ReadStats out;
        out.center = 0.0; out.scale = 1.0; out.n_kept = n; out.flags = 0; out.seg_lo = 0; out.seg_hi = -1;
        out.out_lo = out_lo; out.out_hi = out_hi;

        // ---- leaf sums of the pairwise tree + the histogram's scan, then one barrier for both -------
        double mean = 0.0;
        if (want_sd) {
            if constexpr (!pa_mode) {
                mean = __ddiv_rn((double)tot_sum, (double)n);     // integer samples: the sum is exact in any order
            } else {
                // pA samples are not integers: np.std's mean is itself a pairwise sum
                auto val = [pa_off, pa_unit](int v) -> double { return sqk_pa_value(v, pa_off, pa_unit); };
                s2_tree_leaves(val, bufs, h0, sh, n_out, n);
                __syncthreads();
                if (warp == 0) { const double s = s2_tree_fold(sh, n); if (lane == 0) sh.tree_out = s; }
                __syncthreads();
                mean = __ddiv_rn(sh.tree_out, (double)n);
                __syncthreads();
            }
            if constexpr (!pa_mode) {
                auto sq = [mean](int v) -> double { const double d = __dsub_rn((double)v, mean); return __dmul_rn(d, d); };
                s2_tree_leaves(sq, bufs, h0, sh, n_out, n);
            } else {
                auto sq = [pa_off, pa_unit, mean](int v) -> double {
                    const double d = __dsub_rn(sqk_pa_value(v, pa_off, pa_unit), mean);
                    return __dmul_rn(d, d);
                };
                s2_tree_leaves(sq, bufs, h0, sh, n_out, n);
            }
        }
        int per = 1;
        while (per * SQK_S2_THREADS < nbins) per <<= 1;
        const bool want_med = n > 0 && want_hist;
        if (want_hist && !punt) {
            // also when the read goes to the redo list: the selection pass is what zeroes the histogram
            s2_hist_select<mode != SQK_STATS_MEDMAD>(hist, per, want_med ? (n - 1) / 2 : -1, want_med ? n / 2 : -1, sh, 0);
        }
        __syncthreads();                                                    // B3
        double sd = 0.0;
        if (want_sd && warp == 0) {
            sd = __dsqrt_rn(__ddiv_rn(s2_tree_fold(sh, n), (double)n));
            if (lane == 0) sh.tree_out = sd;
        }
        if (tid == 0) sh.out_cnt = 0;

        if (redo) {
            if (tid == 0) A.redo[atomicAdd(A.n_redo, 1u)] = (int)i;
        } else if (mode == SQK_STATS_ZSCALE || mode == SQK_STATS_NONE) {
            if (tid == 0) {
                if (mode == SQK_STATS_ZSCALE && n > 0) {
                    if (sd == 0.0) sd = 1.0;                  // sklearn _handle_zeros_in_scale
                    out.center = mean; out.scale = sd;
                }
                a.stats[i] = out;
                if (a.n_kept_out) a.n_kept_out[i] = n;
            }
        }
        if constexpr (mode == SQK_STATS_MEDMAD) if (!punt) {
            // The MAD comes from the same histogram, folded around the median: class t of the doubled distance
            // |2v - med2| (= 2t + parity of med2) holds the values ceil(med2/2) + t and floor(med2/2) - t (one value
            // when they coincide).  No second pass over the samples.
            double scaled = 0.0, median = 0.0;
            const int lo_v = wlo + sh.sel[0], hi_v = wlo + sh.sel[1];
            const int med2 = (n > 0 && !redo) ? lo_v + hi_v : 2 * wlo;
            const int up0 = (med2 + (med2 & 1)) / 2 - wlo, dn0 = (med2 - (med2 & 1)) / 2 - wlo;   // as bins; med2 >= 2*wlo
            uint32_t fv[16];
#pragma unroll
            for (int q = 0; q < 16; q++) {
                fv[q] = 0;
                if (q < per) {
                    const int t = tid * per + q;
                    const int ub = up0 + t, db = dn0 - t;
                    uint32_t cnt = (ub < nbins) ? hist[ub] : 0u;
                    if (db >= 0 && db != ub) cnt += hist[db];
                    fv[q] = cnt;
                }
            }
            __syncthreads();                                                // every thread has read `hist`
#pragma unroll
            for (int q = 0; q < 16; q++)
                if (q < per) { fold[tid * per + q] = fv[q]; hist[tid * per + q] = 0; }
            __syncthreads();
            const bool go = n > 0 && !redo;
            s2_hist_select<true>(fold, per, go ? (n - 1) / 2 : -1, go ? n / 2 : -1, sh, 1);
            __syncthreads();
            if (go) {
                median = (double)med2 * 0.5;
                const int par = med2 & 1;
                const int d0 = 2 * sh.sel[0] + par, d1 = 2 * sh.sel[1] + par;   // the two middle doubled distances
                scaled = __dmul_rn((double)(d0 + d1) * 0.25, 1.4826);
            }
            if (tid == 0 && !redo) {
                if (n > 0) {
                    out.center = median; out.scale = scaled;
                    if (scaled == 0.0) out.flags |= SQK_FLAG_DEGENERATE;
                }
                a.stats[i] = out;
                if (a.n_kept_out) a.n_kept_out[i] = n;
            }
        }
        if constexpr (mode == SQK_STATS_SEGMENTER) if (!redo) {
            __syncthreads();                                                // B4: sd and sel visible
            int seg_lo = 0, seg_hi = -1;
            if (n > 0) {
                const int lo_v = wlo + sh.sel[0], hi_v = wlo + sh.sel[1];
                const double sdv = sh.tree_out;
                double median;
                if (!pa_mode) median = (double)(lo_v + hi_v) * 0.5;
                else if (n & 1) median = sqk_pa_value(lo_v, pa_off, pa_unit);
                else median = __ddiv_rn(__dadd_rn(sqk_pa_value(lo_v, pa_off, pa_unit), sqk_pa_value(hi_v, pa_off, pa_unit)), 2.0);
                const double spread = __dmul_rn(sdv, a.std_scale);
                const double top = __dadd_rn(median, spread);
                const double bot = __dsub_rn(median, spread);
                if (!pa_mode) {
                    // integer x:  x < top  <=>  x <= ceil(top)-1 ;  x > bot  <=>  x >= floor(bot)+1
                    const double hi_d = fmin(fmax(ceil(top) - 1.0, -40000.0), 40000.0);
                    const double lo_d = fmin(fmax(floor(bot) + 1.0, -40000.0), 40000.0);
                    seg_hi = (top == top) ? (int)hi_d : -40000;   // NaN threshold: nothing is in range
                    seg_lo = (bot == bot) ? (int)lo_d : 40000;
                } else {
                    seg_hi = (top == top) ? stats_last_below(top, pa_off, pa_unit) : -40000;
                    seg_lo = (bot == bot) ? stats_first_above(bot, pa_off, pa_unit) : 40000;
                }
                out.seg_lo = seg_lo; out.seg_hi = seg_hi;
                out.center = top; out.scale = bot;
            }
            if (tid == 0) {
                a.stats[i] = out;
                if (a.n_kept_out) a.n_kept_out[i] = n;
            }
            if (A.mask) {
                // raw-space in-range bits, one 32-bit word per 4 units (bit B = buffer sample B)
                const int slo = seg_lo < -32768 ? -32768 : seg_lo, shi = seg_hi > 32767 ? 32767 : seg_hi;
                const unsigned sspan = (unsigned)(shi - slo);
                const bool any = shi >= slo && n > 0;
                const int raw_words = (rd.units + 3) >> 2;
                for (int wd = tid; wd < raw_words + 1; wd += SQK_S2_THREADS) {
                    uint32_t bits = 0;
                    if (any && wd < raw_words) {
#pragma unroll
                        for (int uu = 0; uu < 4; uu++) {
                            const int u = 4 * wd + uu;
                            if (u < rd.units) {
                                const int4 q = s2_lds128(bufs + s2_unit_off(u));
                                const int w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                                for (int e = 0; e < 8; e++) {
                                    const int v = (e & 1) ? (w[e >> 1] >> 16) : (int)(short)(w[e >> 1] & 0xffff);
                                    if ((unsigned)(v - slo) <= sspan) bits |= 1u << (8 * uu + e);
                                }
                            }
                        }
                    }
                    rawmask[wd] = bits;
                }
                __syncthreads();
                // compacted word t holds kept samples [32t, 32t+32): a funnel shift of the raw-space words unless an
                // outlier sits inside
                uint32_t *row = A.mask + (int64_t)i * A.mask_stride;
                const int words = (n + 31) >> 5;
                for (int t = tid; t < A.mask_stride; t += SQK_S2_THREADS) {
                    uint32_t wv = 0;
                    if (t < words) {
                        const int c0 = 32 * t;
                        const int clast = c0 + 31 < n ? c0 + 31 : n - 1;
                        const int s0 = n_out ? s2_shift(sh, n_out, c0) : 0;
                        const int s1 = n_out ? s2_shift(sh, n_out, clast) : 0;
                        if (s0 == s1) {
                            const int R = h0 + c0 + s0;
                            wv = __funnelshift_r(rawmask[R >> 5], rawmask[(R >> 5) + 1], R & 31);
                        } else {
                            // outliers inside: one funnel shift per piece between them
                            int c = c0, sft = s0;
                            while (c <= clast) {
                                while (sft < n_out && sh.out_adj[sft] <= c) sft++;     // outliers in front of kept sample c
                                int stop = clast + 1;                                   // first kept sample of the next piece
                                if (sft < n_out && sh.out_adj[sft] <= clast) stop = sh.out_adj[sft];
                                const int R = h0 + c + sft;
                                uint32_t piece = __funnelshift_r(rawmask[R >> 5], rawmask[(R >> 5) + 1], R & 31);
                                const int plen = stop - c;
                                if (plen < 32) piece &= (1u << plen) - 1u;
                                wv |= piece << (c - c0);
                                c = stop;
                            }
                        }
                        const int valid = n - c0;
                        if (valid < 32) wv &= (1u << valid) - 1u;
                    }
                    row[t] = wv;
                }
            }
        }
        __syncthreads();                                                    // end of read: shared scratch reusable
        rd = nx;
    }
}
