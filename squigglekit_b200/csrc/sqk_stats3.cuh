// sqk_stats3.cuh -- K1, third generation: ONE WARP PER READ, for reads of up to SQK_S3_MAX_LEN samples in the zscale,
// medmad, segmenter (raw-integer) and "none" modes.  Same outputs, bit for bit, as sqk_stats_kernel / sqk_stats2_kernel.
//
// What changed against the second generation (profiles/r02_stats2_*: 5.5 k warp-instructions per 4096-sample read in
// zscale mode, 9.5 k in segmenter mode, 21-24 of 32 lanes active, six CTA barriers per read):
//   * a warp owns a read from the bulk copy to the result: no CTA barrier anywhere, no cross-warp reduction through
//     shared memory, every phase hand-over is a __syncwarp.  The CTA is one warp with ONE staging buffer (shared memory per
//     warp limits the warps per SM, and the phases are dependent chains that need warps); the next read's bulk copy
//     (cp.async.bulk + mbarrier) is issued as soon as the last pass over the staged samples is done.
//   * the pairwise tree (numpy's order, sqk_stats_plan.cuh) is walked from a DENSE list of leaves: one 32-bit word per
//     leaf (shared address of its first sample | length | slot), built once per read with ballots.  Leaves that contain an
//     outlier are compacted in place first (inside their own span of the staged read; the exception list is rewritten to
//     match), so the summation loop knows only one kind of leaf: a contiguous run of int16 at some 2-byte aligned
//     address.  No per-sample exception handling, no divergence between "clean" and "unclean" teams, no side buffer.
//   * the four 8-lane teams of the warp sum four leaves at a time, skewed by one row (16 bytes) per team: leaves lie
//     256 bytes apart, so unskewed teams would always hit the same four banks.  Rows outside a team's leaf are predicated
//     off; the loads themselves are unconditional (the layout keeps 64 bytes of slack in front of and 384 bytes behind
//     everything a skewed row can touch).
//   * the histogram for the median takes every sample with one unconditional shared-memory atomic (clean 16-byte units)
//     and is searched by chunk totals (one REDUX per 128 bins) instead of a block-wide scan.
//   * the in-range bit mask for K3 is built with two packed instructions per sample pair (VIADDMNMX.U16x2,
//     VIADDMNMX.S16x2.RELU leave exactly 0/1 in each half), one shift-add per pair and one 16-instruction perfect shuffle
//     per 32 samples.
#pragma once
#include "sqk_stats2.cuh"

#define SQK_S3_MAXOUT 32            // outliers per read the exception list holds
#define SQK_S3_SLOTS 128            // leaf slots: depth <= 7 for n <= 8192
#define SQK_S3_MAX_LEN SQK_S2_MAX_LEN
#define SQK_S3_MAX_BINS 2048
#define SQK_S3_FRONT_SLACK 64
#define SQK_S3_BACK_SLACK 384

struct Stats3Args {
    StatsArgs s;              // cap / gstage / list unused
    int buf_bytes;            // bytes of one staging buffer (multiple of 16)
    int hist_words;           // 0, or a multiple of 128 that covers the outlier window (segmenter)
    int mask_words;           // words of the raw-space mask scratch (multiple of 4), or 0
    uint32_t *mask;           // segmenter: [n_reads][mask_stride] in-range bits of the kept samples, or null
    int mask_stride;
    int *redo;                // reads handed to sqk_stats_kernel (launch-local indices)
    unsigned int *n_redo;
};

struct S3Shared {
    unsigned long long bar[2];
    double leafsum[SQK_S3_SLOTS];
    uint32_t list[SQK_S3_SLOTS + 4];    // dense leaf list: (shared address >> 1) | len << 17 | slot << 25, padded to a multiple of 4
    int out_pos[SQK_S3_MAXOUT];         // raw positions of the outliers (relative to the read's first sample), unsorted
    int out_adj[SQK_S3_MAXOUT];         // sorted, minus rank: outlier i sits in front of kept sample out_adj[i]
    int patch_src[SQK_S3_MAXOUT][4];    // leaves with an outlier inside: off, len, outliers in front of / up to the end of the leaf
    int out_cnt, patch_cnt, pad0, pad1;
};

static inline size_t sqk_s3_smem_bytes(const Stats3Args &A)
{
    return ((sizeof(S3Shared) + 15) & ~(size_t)15) + 4 * (size_t)(A.hist_words + A.mask_words) + SQK_S3_FRONT_SLACK +
           (size_t)A.buf_bytes + SQK_S3_BACK_SLACK;
}

__device__ __forceinline__ int s3_lds_s16(unsigned addr)
{
    int v;
    asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// ++hist[v - wlo] by shared address (hist_b = shared address of the bin of raw value 0): two integer instructions + ATOMS
__device__ __forceinline__ void s3_hist_inc(unsigned hist_b, int v)
{
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(hist_b + 4u * (unsigned)v) : "memory");
}
__device__ __forceinline__ void s3_sts_u16(unsigned addr, int v)
{
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((short)v) : "memory");
}

// Start staging a read into the buffer at shared address `bufs`; called by the whole warp.  Only bulk copies: a read whose
// 16-byte hull sticks out of the allocation (the first / last read of a buffer at most) goes to the redo list.
__device__ __forceinline__ void s3_stage(const StatsArgs &a, const S2Read &rd, unsigned bufs, unsigned bar)
{
    if (rd.units == 0 || !rd.tma) return;
    if (threadIdx.x == 0) {
        const int16_t *src = a.base + (rd.begin - rd.h0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic-proxy accesses to the buffer
        s2_mbar_expect_tx(bar, (unsigned)rd.units * 16u);
        s2_bulk_g2s(bufs, src, (unsigned)rd.units * 16u, bar);
    }
}

// Rare paths, out of line: the kernel's hot code has to stay small (one-warp CTAs at different places of a long
// straight-line kernel share the instruction cache; the first version stalled on instruction fetch more than on data).

// a 16-byte unit with outliers inside: only the flagged words are looked at, half by half
template <bool HIST>
__device__ __noinline__ int s3_unit_outliers(int4 wq, int4 dq, int pos0, int *out_cnt, int *out_pos, uint32_t *hist_m)
{
    const int w[4] = {wq.x, wq.y, wq.z, wq.w}, dw[4] = {dq.x, dq.y, dq.z, dq.w};
    int sub = 0;
#pragma unroll
    for (int e = 0; e < 4; e++) {
        const int vl = (int)(short)(w[e] & 0xffff), vh = w[e] >> 16;
        if (dw[e] & 0xffff) {
            sub += vl;
            const int at = atomicAdd(out_cnt, 1);
            if (at < SQK_S3_MAXOUT) out_pos[at] = pos0 + 2 * e;
        } else if (HIST) atomicAdd(hist_m + vl, 1u);
        if (dw[e] & 0xffff0000) {
            sub += vh;
            const int at = atomicAdd(out_cnt, 1);
            if (at < SQK_S3_MAXOUT) out_pos[at] = pos0 + 2 * e + 1;
        } else if (HIST) atomicAdd(hist_m + vh, 1u);
    }
    return sub;                                         // what the packed sum has to lose again
}

// Leaves with an outlier inside are compacted IN PLACE: the kept samples of such a leaf are moved to the front of the
// leaf's own span of the staged read (they start where the leaf's first kept sample already is), which leaves as many
// stale slots at the end of the span as it had outliers.  Afterwards every leaf is a contiguous run again, and the
// exception list is rewritten to say so: the leaf's outliers now sit in front of the kept sample that follows the leaf.
// (Leaves are at most 128 samples: four per lane, all loads before any store.)
__device__ __noinline__ void s3_patch_in_place(S3Shared &sh, int n_patch, unsigned bufs, int h0)
{
    const int lane = threadIdx.x;
    for (int pi = 0; pi < n_patch; pi++) {
        const int off = sh.patch_src[pi][0], ln = sh.patch_src[pi][1], c0 = sh.patch_src[pi][2], c1 = sh.patch_src[pi][3];
        const int adj0 = sh.out_adj[c0], adj1 = c1 - c0 > 1 ? sh.out_adj[c0 + 1] : 0x7fffffff;   // nearly always one or two
        int v[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int c = lane + 32 * q;
            v[q] = 0;
            if (c < ln) {
                const int g = off + c;
                int shf = c0 + ((adj0 <= g) ? 1 : 0) + ((adj1 <= g) ? 1 : 0);
                for (int j = c0 + 2; j < c1; j++) shf += (sh.out_adj[j] <= g) ? 1 : 0;
                v[q] = s2_lds_s16(bufs + 2u * (unsigned)(h0 + g + shf));
            }
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int c = lane + 32 * q;
            if (c < ln) s3_sts_u16(bufs + 2u * (unsigned)(h0 + off + c0 + c), v[q]);
        }
        __syncwarp();
    }
    // (after all leaves: a leaf's rewrite must not change what a later leaf's search sees -- entries only move to
    // positions <= the next leaf's first index, so doing it at the end is merely simpler to reason about)
    for (int pi = 0; pi < n_patch; pi++) {
        const int off = sh.patch_src[pi][0], ln = sh.patch_src[pi][1], c0 = sh.patch_src[pi][2], c1 = sh.patch_src[pi][3];
        for (int j = c0 + lane; j < c1; j += 32) sh.out_adj[j] = off + ln;
    }
    __syncwarp();
}

// A read with more outliers than the exception list holds (rare): compact the whole staged read in place, 32 samples per
// step (kept samples only move towards the front, and a step's loads are done before its stores).  Afterwards it is a read
// without outliers.
__device__ __noinline__ void s3_compact_all(unsigned bufs, int h0, int len, int wlo, unsigned span)
{
    const int lane = threadIdx.x;
    int wpos = 0;
    for (int p0 = 0; p0 < len; p0 += 32) {
        const int p = p0 + lane;
        const int v = p < len ? s2_lds_s16(bufs + 2u * (unsigned)(h0 + p)) : 0;
        const bool keep = p < len && (unsigned)(v - wlo) <= span;
        const unsigned m = __ballot_sync(SQK_FULL_MASK, keep);
        __syncwarp();                                           // every lane's load before any lane's store
        if (keep) s3_sts_u16(bufs + 2u * (unsigned)(h0 + wpos + __popc(m & ((1u << lane) - 1u))), v);
        wpos += __popc(m);
        __syncwarp();
    }
}

// compacted mask word of kept samples [c0, clast] when outliers sit inside: one funnel shift per piece between them
__device__ __noinline__ uint32_t s3_mask_pieces(const S3Shared &sh, const uint32_t *rawmask, int n_out, int h0, int c0, int clast, int s0)
{
    uint32_t wv = 0;
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
    return wv;
}

// Leaf rows of one round.  FULL: all four teams of the warp have a 16-row leaf, so skewed rows 3..15 are rows of every
// team's leaf and need no predicate; otherwise every row is predicated.
template <bool FULL, class Term>
__device__ __forceinline__ double s3_rows(Term term, unsigned base, unsigned rows_mask)
{
    double acc = 0.0;                                   // 0.0 + t == t exactly for the (non-negative) terms
#pragma unroll
    for (int i = 0; i < 19; i++) {
        const double t = term(s3_lds_s16(base + 16u * (unsigned)i));
        if (FULL && i >= 3 && i < 16) acc = __dadd_rn(acc, t);
        else if (rows_mask & (1u << i)) acc = __dadd_rn(acc, t);
    }
    return acc;
}

// 16 + 16 -> 32 bit perfect shuffle: bit j of the result's even positions from x[j], odd positions from x[16 + j]
__device__ __forceinline__ uint32_t s3_interleave16(uint32_t x)
{
    uint32_t t;
    t = (x ^ (x >> 8)) & 0x0000ff00u; x = x ^ t ^ (t << 8);
    t = (x ^ (x >> 4)) & 0x00f000f0u; x = x ^ t ^ (t << 4);
    t = (x ^ (x >> 2)) & 0x0c0c0c0cu; x = x ^ t ^ (t << 2);
    t = (x ^ (x >> 1)) & 0x22222222u; x = x ^ t ^ (t << 1);
    return x;
}

template <int MODE>
__global__ void __launch_bounds__(32) sqk_stats3_kernel(const Stats3Args A)
{
    const StatsArgs &a = A.s;
    extern __shared__ __align__(16) unsigned char s3_smem[];
    S3Shared &sh = *reinterpret_cast<S3Shared *>(s3_smem);
    constexpr unsigned FIXED = (sizeof(S3Shared) + 15) & ~15u;
    constexpr bool HIST = (MODE == SQK_STATS_SEGMENTER || MODE == SQK_STATS_MEDMAD);
    constexpr bool WANT_SD = (MODE == SQK_STATS_ZSCALE || MODE == SQK_STATS_SEGMENTER);
    uint32_t *hist = reinterpret_cast<uint32_t *>(s3_smem + FIXED);
    uint32_t *rawmask = hist + A.hist_words;
    const unsigned bufs = s2_smem_addr(rawmask + A.mask_words) + SQK_S3_FRONT_SLACK;
    const unsigned bar = s2_smem_addr(&sh.bar[0]);

    const int lane = threadIdx.x, team = lane >> 3, k = lane & 7;
    int64_t alloc_lo = a.alloc_lo, alloc_hi = a.alloc_hi;
    resolve_bounds(a.offsets, a.read0, a.n_reads, alloc_lo, alloc_hi);

    if (lane == 0) {
        s2_mbar_init(bar, 1);
        sh.out_cnt = 0; sh.patch_cnt = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int b = lane; b < A.hist_words; b += 32) hist[b] = 0;
    __syncwarp();

    // ---- outlier window on the raw sample (inclusive), clipped to the int16 range: the same for every read --------
    const int out_lo = a.lo + 1, out_hi = a.hi - 1;
    const int wlo = out_lo < -32768 ? -32768 : out_lo, whi = out_hi > 32767 ? 32767 : out_hi;
    const bool window_ok = whi >= wlo;
    const int nbins = window_ok ? whi - wlo + 1 : 0;
    const bool hist_ok = !HIST || nbins <= A.hist_words;
    const int lo2 = (wlo & 0xffff) | (wlo << 16), hi2 = (whi & 0xffff) | (whi << 16);
    const unsigned span = (unsigned)(whi - wlo);
    uint32_t *const hist_m = hist - wlo;                      // the bin of raw value 0
    const unsigned hist_b = s2_smem_addr(hist) - 4u * (unsigned)wlo;

    // One staging buffer per warp (shared memory per warp, not the copy engine, limits the warps per SM, and this kernel
    // needs warps: its phases are dependent chains).  The next read's bulk copy is issued as soon as the last pass over
    // the staged samples is done; the other warps of the SM cover the rest of its latency.
    const int64_t stride = gridDim.x;
    int64_t i = blockIdx.x;
    unsigned uses = 0;                     // bulk-staged reads so far (mbarrier phase)
    S2Read rd{};
    if (i < a.n_reads) {
        rd = s2_describe(a, i, alloc_lo, alloc_hi, A.buf_bytes);
        s3_stage(a, rd, bufs, bar);
    }
    for (; i < a.n_reads; i += stride) {
        const int64_t inext = i + stride;
        S2Read nx{};
        if (inext < a.n_reads) nx = s2_describe(a, inext, alloc_lo, alloc_hi, A.buf_bytes);   // (its loads overlap the wait)
        bool staged_next = false;
        auto stage_next = [&]() {
            __syncwarp();                                       // every lane is done with the staged samples (and, for
                                                                // the callers that rely on it, with its shared stores)
            if (!staged_next && inext < a.n_reads) s3_stage(a, nx, bufs, bar);
            staged_next = true;
        };
        if (rd.units > 0 && rd.tma) {
            s2_mbar_wait(bar, uses & 1u);
            uses++;
        }

        const bool punt = !rd.ok || !rd.tma || !window_ok || !hist_ok;     // not for this kernel: the redo list takes it
        const int h0 = rd.h0, len = rd.len;
        // ---- phase A: window test, integer sum, histogram, outlier list --------------------------------------------
        int lsum = 0;
        if (!punt) {
            const int ua = h0 ? 1 : 0, ub = (h0 + len) >> 3;    // whole 16-byte units of the read: [ua, ub)
            for (int u = ua + lane; u < ub; u += 32) {
                const int4 q = s2_lds128(bufs + 16u * (unsigned)u);
                const int w[4] = {q.x, q.y, q.z, q.w};
                int dw[4];
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    dw[e] = s2_clamp2(w[e], lo2, hi2) ^ w[e];   // != 0 in the halves that hold an outlier
                    lsum = __dp2a_lo(w[e], 0x0101, lsum);
                }
                if (((dw[0] | dw[1]) | (dw[2] | dw[3])) == 0) {
                    if (HIST) {
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            s3_hist_inc(hist_b, (int)(short)(w[e] & 0xffff));
                            s3_hist_inc(hist_b, w[e] >> 16);
                        }
                    }
                } else {
                    lsum -= s3_unit_outliers<HIST>(q, make_int4(dw[0], dw[1], dw[2], dw[3]), 8 * u - h0, &sh.out_cnt, sh.out_pos, hist_m);   // rare
                }
            }
            // the (at most 7 + 7) samples in front of / behind the whole units: one lane each
            const int head_n = ua ? (len < 8 - h0 ? len : 8 - h0) : 0;
            const int tail0 = (8 * ub - h0) > head_n ? (8 * ub - h0) : head_n;
            if (head_n > 0 || tail0 < len) {
                int p = -1;
                if (lane < 7) { if (lane < head_n) p = lane; }
                else if (lane < 14) { const int q = tail0 + lane - 7; if (q < len) p = q; }
                if (p >= 0) {
                    const int v = s2_lds_s16(bufs + 2u * (unsigned)(h0 + p));
                    if ((unsigned)(v - wlo) > span) {
                        const int at = atomicAdd(&sh.out_cnt, 1);
                        if (at < SQK_S3_MAXOUT) sh.out_pos[at] = p;
                    } else {
                        lsum += v;
                        if (HIST) atomicAdd(&hist[v - wlo], 1u);
                    }
                }
            }
        }
        __syncwarp();
        const int tot_sum = __reduce_add_sync(SQK_FULL_MASK, lsum);    // |sum| <= 8176 * 32768 < 2^31
        const int n_out_all = punt ? 0 : sh.out_cnt;
        bool redo = punt;
        int n_out = redo ? 0 : n_out_all;
        const int n = redo ? 0 : len - n_out_all;
        if (n_out > SQK_S3_MAXOUT) {                            // rare: too many for the exception list
            s3_compact_all(bufs, h0, len, wlo, span);
            n_out = 0;
        }
        if constexpr (!WANT_SD) stage_next();                  // medmad / none: the staged samples are not read again

        if (n_out > 0) {
            if (lane < n_out) {
                const int p = sh.out_pos[lane];
                int rank = 0;
                for (int j = 0; j < n_out; j++) rank += (sh.out_pos[j] < p) ? 1 : 0;
                sh.out_adj[rank] = p - rank;
            }
            __syncwarp();
        }

        // ---- the pairwise tree: dense leaf list, patches for leaves with an outlier inside --------------------------
        // list entry: (shared address >> 1) | len << 17 | slot << 25; padded with empty entries to a multiple of four
        int n_leaves = 0, slots = 1;
        const bool want_sd = WANT_SD && n > 0;
        if (want_sd) {
            const int depth = sqk_tree_depth7(n);
            slots = 1 << depth;
            for (int sb = 0; sb < slots; sb += 32) {
                const int s = sb + lane;
                int off = 0, ln = 0;
                const bool mine = sqk_tree_leaf7(n, depth, s, &off, &ln) && s < slots;
                unsigned addr = 0;
                if (mine) {
                    int c0 = 0, c1 = 0;
                    for (int j = 0; j < n_out; j++) {
                        const int adj = sh.out_adj[j];
                        c0 += (adj <= off) ? 1 : 0;
                        c1 += (adj <= off + ln - 1) ? 1 : 0;
                    }
                    addr = bufs + 2u * (unsigned)(h0 + off + c0);   // (a leaf with outliers inside is compacted in place below)
                    if (c0 != c1) {
                        const int pi = atomicAdd(&sh.patch_cnt, 1);    // <= n_out <= SQK_S3_MAXOUT
                        sh.patch_src[pi][0] = off; sh.patch_src[pi][1] = ln; sh.patch_src[pi][2] = c0; sh.patch_src[pi][3] = c1;
                    }
                }
                const unsigned m = __ballot_sync(SQK_FULL_MASK, mine);
                if (mine) sh.list[n_leaves + __popc(m & ((1u << lane) - 1u))] = (addr >> 1) | ((unsigned)ln << 17) | ((unsigned)s << 25);
                if (s < slots) sh.leafsum[s] = 0.0;
                n_leaves += __popc(m);
            }
            if (lane < 3) sh.list[n_leaves + lane] = bufs >> 1;           // empty entries: len 0
            __syncwarp();
            const int n_patch = sh.patch_cnt;
            if (n_patch) s3_patch_in_place(sh, n_patch, bufs, h0);
        }

        ReadStats out;
        out.center = 0.0; out.scale = 1.0; out.n_kept = n; out.flags = 0; out.seg_lo = 0; out.seg_hi = -1;
        out.out_lo = out_lo; out.out_hi = out_hi;

        double mean = 0.0, sd = 0.0;
        if (want_sd && !redo) {
            mean = __ddiv_rn((double)tot_sum, (double)n);       // integer samples: the sum is exact in any order
            auto sq = [mean](int v) -> double { const double d = __dsub_rn((double)v, mean); return __dmul_rn(d, d); };
            const unsigned lane_off = 2u * (unsigned)k - 16u * (unsigned)team;
            const unsigned tmask = 0xffu << (lane & 24);
            for (int r0 = 0; r0 < n_leaves; r0 += 4) {
                const uint32_t d = sh.list[r0 + team];
                const unsigned rows = (d >> 20) & 31u;          // len >> 3
                const unsigned rows_mask = ((1u << rows) - 1u) << team;
                const unsigned rbase = ((d & 0x1ffffu) << 1) + lane_off;
                double acc = __all_sync(SQK_FULL_MASK, rows == 16u) ? s3_rows<true>(sq, rbase, rows_mask) : s3_rows<false>(sq, rbase, rows_mask);
                acc = s2_team_fold(acc, tmask);
                if (k == 0 && (d >> 17) != 0) sh.leafsum[d >> 25] = acc;
            }
            __syncwarp();
            // the ragged end of the last leaf (or a whole read of fewer than 8 samples): serial adds after the team fold
            const uint32_t dl = sh.list[n_leaves - 1];
            const int ln_l = (int)((dl >> 17) & 255u), rem = ln_l & 7;
            if (rem) {
                if (lane == 0) {
                    double acc = sh.leafsum[dl >> 25];
                    const unsigned tb = ((dl & 0x1ffffu) << 1) + 2u * (unsigned)(ln_l - rem);
                    for (int e = 0; e < rem; e++) acc = __dadd_rn(acc, sq(s2_lds_s16(tb + 2u * (unsigned)e)));
                    sh.leafsum[dl >> 25] = acc;
                }
                __syncwarp();
            }
            if (!HIST || A.mask == nullptr) stage_next();       // zscale: nothing reads the staged samples any more
            // fold the slots in slot order (numpy's order): 1, 2 or 4 per lane, then an xor butterfly
            double r;
            if (slots > 64) r = __dadd_rn(__dadd_rn(sh.leafsum[4 * lane], sh.leafsum[4 * lane + 1]), __dadd_rn(sh.leafsum[4 * lane + 2], sh.leafsum[4 * lane + 3]));
            else if (slots > 32) r = __dadd_rn(sh.leafsum[2 * lane], sh.leafsum[2 * lane + 1]);
            else r = lane < slots ? sh.leafsum[lane] : 0.0;
#pragma unroll
            for (int m = 1; m < 32; m <<= 1) r = __dadd_rn(r, shfl_xor_f64(r, m, 32));
            sd = __dsqrt_rn(__ddiv_rn(r, (double)n));
        }

        // ---- medmad: median and MAD from the histogram, turned into prefix sums in place ---------------------------------
        if constexpr (MODE == SQK_STATS_MEDMAD) {
            if (!punt) {
                const int groups = A.hist_words >> 7;           // chunks of 128 bins: lane owns bins 128 c + 4 lane .. + 4
                int carry = 0;
                for (int c = 0; c < groups; c++) {
                    uint4 w = *reinterpret_cast<const uint4 *>(hist + 128 * c + 4 * lane);
                    w.y += w.x; w.z += w.y; w.w += w.z;
                    int incl = (int)w.w;
#pragma unroll
                    for (int dd = 1; dd < 32; dd <<= 1) {
                        const int tt = __shfl_up_sync(SQK_FULL_MASK, incl, dd);
                        if (lane >= dd) incl += tt;
                    }
                    const unsigned add = (unsigned)(carry + incl) - w.w;
                    w.x += add; w.y += add; w.z += add; w.w += add;
                    *reinterpret_cast<uint4 *>(hist + 128 * c + 4 * lane) = w;
                    carry += __shfl_sync(SQK_FULL_MASK, incl, 31);
                }
                __syncwarp();
                if (n > 0 && !redo) {
                    // every lane runs the same searches on the prefix sums (broadcast loads)
                    auto P = [hist, nbins](int b) -> int { return b < 0 ? 0 : (int)hist[b < nbins ? b : nbins - 1]; };
                    auto kth = [&P, nbins](int r) -> int {          // smallest bin b with P(b) > r
                        int lo = 0, hi = nbins - 1;
                        while (lo < hi) { const int mid = (lo + hi) >> 1; if (P(mid) > r) hi = mid; else lo = mid + 1; }
                        return lo;
                    };
                    const int r0 = (n - 1) >> 1, r1 = n >> 1;
                    const int b0 = kth(r0), b1 = P(b0) > r1 ? b0 : kth(r1);
                    const int med2 = 2 * wlo + b0 + b1, par = med2 & 1;              // doubled median, its parity
                    const int lo_b = (med2 - par) / 2 - wlo, hi_b = (med2 + par) / 2 - wlo;
                    // samples with |2v - med2| <= 2t + par  <=>  bins [lo_b - t, hi_b + t]
                    auto within = [&P, lo_b, hi_b](int t) -> int { return P(hi_b + t) - P(lo_b - t - 1); };
                    auto tth = [&within, nbins](int r) -> int {
                        int lo = 0, hi = nbins;
                        while (lo < hi) { const int mid = (lo + hi) >> 1; if (within(mid) > r) hi = mid; else lo = mid + 1; }
                        return lo;
                    };
                    const int t0 = tth(r0), t1 = within(t0) > r1 ? t0 : tth(r1);
                    const int d0 = 2 * t0 + par, d1 = 2 * t1 + par;                // the two middle doubled distances
                    const double scaled = __dmul_rn((double)(d0 + d1) * 0.25, 1.4826);
                    out.center = (double)med2 * 0.5; out.scale = scaled;
                    if (scaled == 0.0) out.flags |= SQK_FLAG_DEGENERATE;
                }
                __syncwarp();
                for (int g = lane; g < (A.hist_words >> 2); g += 32) *reinterpret_cast<uint4 *>(hist + 4 * g) = make_uint4(0, 0, 0, 0);
            }
        }

        // ---- segmenter: the median from the histogram (zeroed on the way), thresholds, in-range bit mask -----------
        if constexpr (MODE == SQK_STATS_SEGMENTER) {
            if (!punt) {
                const int groups = A.hist_words >> 7;           // chunks of 128 bins: lane owns bins 128 c + 4 lane .. + 4
                const int ranks[2] = {(n - 1) >> 1, n >> 1};
                int sel[2] = {0, 0};
                if (n > 0 && !redo) {
                    // chunk by chunk (128 bins: lane owns bins 128 c + 4 lane .. + 4) until the chunk that holds the rank,
                    // found from the chunk totals (one REDUX each); the two ranks are equal or adjacent, so the second
                    // search almost always ends in the chunk the first one stopped in
                    int cbase = 0, c = -1, tot = 0;             // samples in front of chunk c / inside it
                    uint4 wv = make_uint4(0, 0, 0, 0);
#pragma unroll
                    for (int t = 0; t < 2; t++) {
                        const int rk = ranks[t];
                        while ((c < 0 || rk >= cbase + tot) && c + 1 < groups) {
                            cbase += tot; c++;
                            wv = *reinterpret_cast<const uint4 *>(hist + 128 * c + 4 * lane);
                            tot = __reduce_add_sync(SQK_FULL_MASK, (int)(wv.x + wv.y + wv.z + wv.w));
                        }
                        // inside the chunk: inclusive scan of the lanes' 4-bin sums
                        const int mine = (int)(wv.x + wv.y + wv.z + wv.w);
                        int incl = mine;
#pragma unroll
                        for (int dd = 1; dd < 32; dd <<= 1) {
                            const int tt = __shfl_up_sync(SQK_FULL_MASK, incl, dd);
                            if (lane >= dd) incl += tt;
                        }
                        const int excl = cbase + incl - mine;
                        int bin = -1;
                        if (rk >= excl && rk < excl + mine) {
                            int cum = excl;
                            bin = 0;
                            if (rk >= cum + (int)wv.x) { cum += (int)wv.x; bin = 1;
                                if (rk >= cum + (int)wv.y) { cum += (int)wv.y; bin = 2;
                                    if (rk >= cum + (int)wv.z) { bin = 3; } } }
                            bin += 128 * c + 4 * lane;
                        }
                        const unsigned who = __ballot_sync(SQK_FULL_MASK, bin >= 0);
                        sel[t] = __shfl_sync(SQK_FULL_MASK, bin, who ? __ffs((int)who) - 1 : 0);
                    }
                }
                __syncwarp();
                for (int g = lane; g < (A.hist_words >> 2); g += 32) *reinterpret_cast<uint4 *>(hist + 4 * g) = make_uint4(0, 0, 0, 0);

                int seg_lo = 0, seg_hi = -1;
                if (n > 0 && !redo) {
                    const int lo_v = wlo + sel[0], hi_v = wlo + sel[1];
                    const double median = (double)(lo_v + hi_v) * 0.5;
                    const double spread = __dmul_rn(sd, a.std_scale);
                    const double top = __dadd_rn(median, spread);
                    const double bot = __dsub_rn(median, spread);
                    // integer x:  x < top  <=>  x <= ceil(top)-1 ;  x > bot  <=>  x >= floor(bot)+1
                    const double hi_d = fmin(fmax(ceil(top) - 1.0, -40000.0), 40000.0);
                    const double lo_d = fmin(fmax(floor(bot) + 1.0, -40000.0), 40000.0);
                    seg_hi = (top == top) ? (int)hi_d : -40000;   // NaN threshold: nothing is in range
                    seg_lo = (bot == bot) ? (int)lo_d : 40000;
                    out.seg_lo = seg_lo; out.seg_hi = seg_hi;
                    out.center = top; out.scale = bot;
                }
                // the packed in-range test needs the window span to fit a signed 16-bit lane; wider windows (thresholds
                // further apart than half the int16 range) are left to the first-generation kernel
                const int slo = seg_lo < -32768 ? -32768 : seg_lo, shi = seg_hi > 32767 ? 32767 : seg_hi;
                const bool any = shi >= slo && n > 0;
                const int sspan = shi - slo;
                if (A.mask && !redo && any && sspan > 32766) redo = true;
                if (A.mask && !redo) {
                    // raw-space in-range bits, one 32-bit word per 4 units (bit B = buffer sample B)
                    const int raw_words = (rd.units + 3) >> 2;
                    const unsigned negslo2 = ((unsigned)(-slo) & 0xffffu) * 0x10001u;
                    const unsigned lim2 = (unsigned)((sspan + 1) & 0xffff) * 0x10001u;
                    const unsigned negspan2 = ((unsigned)(-sspan) & 0xffffu) * 0x10001u;
                    for (int wd = lane; wd <= raw_words; wd += 32) {
                        uint32_t acc = 0;
                        if (any && wd < raw_words) {
#pragma unroll
                            for (int uu = 0; uu < 4; uu++) {
                                // (units behind the read's last one are inside the buffer or the slack behind it)
                                const int4 q = s2_lds128(bufs + 16u * (unsigned)(4 * wd + uu));
                                const unsigned w[4] = {(unsigned)q.x, (unsigned)q.y, (unsigned)q.z, (unsigned)q.w};
#pragma unroll
                                for (int e = 0; e < 4; e++) {
                                    // per half: min(v - slo mod 2^16, span + 1) - span, floored at 0  ->  1 = out of range
                                    const unsigned t = __viaddmin_u16x2(w[e], negslo2, lim2);
                                    const unsigned o = __viaddmax_s16x2_relu(t, negspan2, 0u);
                                    acc += o << (4 * uu + e);        // even samples -> bits 0..15, odd -> 16..31
                                }
                            }
                            acc = ~s3_interleave16(acc);
                        }
                        rawmask[wd] = acc;
                    }
                    stage_next();                               // (syncs the warp) the staged samples are not read again
                    // compacted word t holds kept samples [32t, 32t+32): a funnel shift of the raw-space words unless an
                    // outlier sits inside
                    uint32_t *row = A.mask + (int64_t)i * A.mask_stride;
                    const int words = (n + 31) >> 5;
                    for (int t = lane; t < A.mask_stride; t += 32) {
                        uint32_t wv = 0;
                        if (t < words) {
                            const int c0 = 32 * t;
                            const int clast = c0 + 31 < n ? c0 + 31 : n - 1;
                            int s0 = 0, s1 = 0;
                            for (int j = 0; j < n_out; j++) {
                                const int adj = sh.out_adj[j];
                                s0 += (adj <= c0) ? 1 : 0;
                                s1 += (adj <= clast) ? 1 : 0;
                            }
                            if (s0 == s1) {
                                const int R = h0 + c0 + s0;
                                wv = __funnelshift_r(rawmask[R >> 5], rawmask[(R >> 5) + 1], R & 31);
                            } else {
                                wv = s3_mask_pieces(sh, rawmask, n_out, h0, c0, clast, s0);
                            }
                            const int valid = n - c0;
                            if (valid < 32) wv &= (1u << valid) - 1u;
                        }
                        row[t] = wv;
                    }
                }
            }
        }

        stage_next();
        if (redo) {
            if (lane == 0) A.redo[atomicAdd(A.n_redo, 1u)] = (int)i;
        } else if (lane == 0) {
            if (MODE == SQK_STATS_ZSCALE && n > 0) {
                if (sd == 0.0) sd = 1.0;                  // sklearn _handle_zeros_in_scale
                out.center = mean; out.scale = sd;
            }
            a.stats[i] = out;
            if (a.n_kept_out) a.n_kept_out[i] = n;
        }
        if (lane == 0) { sh.out_cnt = 0; sh.patch_cnt = 0; }
        __syncwarp();                                      // end of read: shared scratch reusable
        rd = nx;
    }
}
