// sqk_stats.cuh -- K1: per-read outlier compaction + the statistics that normalisation
// (MotifSeq.py:186-200) and get_segs (segmenter.py:407-414) need, bit-identical to numpy.
//
// A read is owned by a group of NT threads: NT = 128 (one CTA per read) by default; NT = 32 (one warp per
// read, four reads per CTA, only __syncwarp between phases) exists for experiments (SQK_STATS_NT=32) and
// measured slower on B200: fewer resident warps per SM than the CTA-per-read form.  The
// read is fetched from HBM ONCE with 16-byte streaming loads, filtered  lo < s < hi  (scale_outliers,
// MotifSeq.py:317-324 / segmenter.py:311-318) with a popc prefix, and the survivors are parked as int16
// in shared memory (reads longer than the shared-memory window: a global scratch row); every later
// pass runs from that copy.
//
//   zscale    mean = sum/n (integer sum: exact, order-free);  sd = sqrt(S/n) with S = the sum of
//             fl(fl(x-mean)^2) taken in numpy's pairwise order: 8-lane teams own the <=128-element
//             leaves (one lane per strided accumulator, xor-butterfly = numpy's fold); each team finds its leaf by
//             walking the split rule down from the root, one warp folds the leaf sums of a subtree of <= 8192
//             elements in slot order, the levels above such subtrees are walked depth-first
//             -> same bits as sklearn.preprocessing.scale / np.std.
//   medmad    median and MAD from ONE shared-memory histogram of the raw values + prefix scan when the
//             outlier window spans <= 2048 values (MAD by binary search on count(|2v - 2med| <= D)),
//             else by two-pass radix select: exact, including the x.5 medians of even-length reads.
//   segmenter median + sd -> thresholds -> integer window seg_lo <= x <= seg_hi equivalent to
//             bot < x < top; pA mode evaluates convert_to_pA_numpy + np.round per sample with numpy's
//             roundings (monotone in the raw value, so every window stays an integer window).
//             The state machine itself is K3 (sqk_segmenter.cuh).  (A variant that ran get_segs here, on a bit
//             mask of the staged samples with popc run-skipping, was measured and dropped: one thread per read
//             walking ~300 candidate segments is slower than K3's 32 reads per warp.)
//   adapter   dRNA_segmenter.py:104-106: median and sd of the post-outlier samples [t_start, t_end) only,
//             top = median + sd*std_scale, in range <=> x < top -> seg_hi = ceil(top)-1 (sqk_adapter.cuh has the
//             state machine).
#pragma once
#include "sqk_common.cuh"
#include "sqk_stats_plan.cuh"

#define SQK_STATS_THREADS 128       // CTA size for NT <= 128; a read is owned by NT = 32, 128 or 256 threads
#define SQK_STATS_MAX_WARPS 8       // warps per read at most (NT = 256: reads so long that shared memory, not registers,
                                    // limits the CTAs per SM -- twice the warps per staged read)
#define SQK_TREE_DEPTH 24
#define SQK_HIST_BINS 2048          // direct histogram when the outlier window spans <= 2048 raw values
#define SQK_RADIX_BINS 512          // fallback two-pass radix select (9 + 8 bits)
#define SQK_TREE_SLOTS 128          // leaf slots of the parallel pairwise tree: depth <= 7 for n <= SQK_HEAP_MAX_N
#define SQK_HEAP_MAX_N 8192

enum { SQK_STATS_ZSCALE = 0, SQK_STATS_MEDMAD = 1, SQK_STATS_NONE = 2, SQK_STATS_SEGMENTER = 3, SQK_STATS_ADAPTER = 4 };

struct StatsArgs {
    const int16_t *base;      // base[i] = absolute sample i
    int64_t alloc_lo, alloc_hi;
    const int64_t *offsets;   // absolute sample offsets, [.. read0 + n_reads]
    int64_t read0, n_reads;
    ReadStats *stats;         // [n_reads], launch-local index
    int32_t *n_kept_out;      // [n_reads] or null
    int mode, lo, hi, num;
    int t_start, t_end;       // SQK_STATS_ADAPTER: the statistics come from kept samples [t_start, t_end) only
    double std_scale;
    const double *pa_offset;  // segmenter pA mode: per-read calibration, indexable by absolute read id (or null)
    const double *pa_scale;   //   pA = round((d + pa_offset) * pa_scale, 2)
    int cap;                  // shared-memory staging capacity per group (samples, multiple of 64)
    int16_t *gstage;          // global staging rows for reads longer than cap (or null)
    int64_t gstage_stride;
    const int *list;          // optional work list (launch-local read indices) instead of all n_reads reads ...
    const unsigned int *n_list;   // ... and its length (device memory): the redo list of sqk_stats2_kernel
    int extra_flags;          // OR-ed into ReadStats.flags (SQK_FLAG_NO_MASK for redo reads)
};

struct StatsShared {
    union {                                           // the two users never overlap in time
        uint32_t hist[SQK_HIST_BINS + SQK_HIST_BINS / 32];   // median / MAD histogram, then its prefix sums (padded, see HB)
        double leaf[SQK_TREE_SLOTS];                  // leaf sums of the pairwise tree (n <= SQK_HEAP_MAX_N), in slot order
    };
    double tree_out;
    unsigned long long sum_part[SQK_STATS_MAX_WARPS];
    int warp_tot[SQK_STATS_MAX_WARPS];
    int scan_tot[2][4][SQK_STATS_MAX_WARPS];          // compaction pass: [iteration parity][block of the thread][warp]
    uint32_t scan_part[SQK_STATS_MAX_WARPS];
    uint32_t sel[2];
};

// barrier over the NT threads that own a read
template <int NT> __device__ __forceinline__ void stats_sync()
{
    if (NT == 32) __syncwarp(); else __syncthreads();
}
template <int NT> __device__ __forceinline__ int stats_sync_or(int pred)
{
    if (NT == 32) return __any_sync(SQK_FULL_MASK, pred);
    return __syncthreads_or(pred);
}

// numpy pairwise leaf (n <= 128) over term(i), evaluated by an 8-lane team (lane k = accumulator k).
template <class Term>
__device__ __forceinline__ double stats_leaf_sum(Term term, int off, int len, int k)
{
    if (len < 8) {
        double r = 0.0;
        for (int i = 0; i < len; i++) r = __dadd_rn(r, term(off + i));
        return r;
    }
    const int body = len - (len & 7);
    double r = term(off + k);
    for (int i = 8; i < body; i += 8) r = __dadd_rn(r, term(off + i + k));
    r = __dadd_rn(r, shfl_xor_f64(r, 1, 8));   // (r0+r1) (r2+r3) (r4+r5) (r6+r7)
    r = __dadd_rn(r, shfl_xor_f64(r, 2, 8));   // ((r0+r1)+(r2+r3)) ...
    r = __dadd_rn(r, shfl_xor_f64(r, 4, 8));
    for (int i = body; i < len; i++) r = __dadd_rn(r, term(off + i));
    return r;
}

// Same sum for n <= SQK_HEAP_MAX_N with two barriers.  numpy's split rule (left half = floor(len/2) rounded down to a
// multiple of 8) is a pure function of n, so every 8-lane team finds the leaf of slot j on its own by walking j's bits
// down from the root -- no tree is built in memory.  The right child is never the smaller one, so the depth of the
// right-most path is the depth D of the tree; a node that drops to <= 128 elements above level D is a leaf whose sum is
// stored in the slot of its left-most descendant, the other descendant slots hold 0.0 (x + 0.0 == x exactly).  One warp
// then folds the 2^D slots pairwise in slot order, which is exactly the order in which numpy adds the halves.
template <int NT, class Term>
__device__ double stats_pairwise_heap(Term term, int n, StatsShared &sh)
{
    const int tid = threadIdx.x % NT;
    constexpr int TEAMS = NT / 8;
    const int depth = sqk_tree_depth(n);
    const int slots = 1 << depth;                      // <= SQK_TREE_SLOTS
    const int team = tid >> 3, k = tid & 7;
    for (int j0 = 0; j0 < slots; j0 += TEAMS) {
        const int j = j0 + team;
        int off = 0, len = 0;
        const bool mine = sqk_tree_leaf(n, depth, j < slots ? j : 0, &off, &len) && j < slots;
        // every team runs the shuffles; teams without a leaf sum a dummy one at offset 0
        const double v = stats_leaf_sum(term, mine ? off : 0, mine ? len : (n >= 8 ? 8 : n), k);
        if (j < slots && k == 0) sh.leaf[j] = mine ? v : 0.0;
    }
    stats_sync<NT>();
    if (tid < 32) {
        // 2^depth slots -> one warp: lane i folds its run of consecutive slots pairwise, then the lanes fold pairwise
        const int per = slots > 32 ? slots / 32 : 1;   // 1, 2 or 4
        double r;
        if (per == 1) r = tid < slots ? sh.leaf[tid] : 0.0;
        else if (per == 2) r = __dadd_rn(sh.leaf[2 * tid], sh.leaf[2 * tid + 1]);
        else r = __dadd_rn(__dadd_rn(sh.leaf[4 * tid], sh.leaf[4 * tid + 1]), __dadd_rn(sh.leaf[4 * tid + 2], sh.leaf[4 * tid + 3]));
#pragma unroll
        for (int m = 1; m < 32; m <<= 1) r = __dadd_rn(r, shfl_xor_f64(r, m, 32));
        if (tid == 0) sh.tree_out = r;
    }
    stats_sync<NT>();
    const double r = sh.tree_out;
    stats_sync<NT>();
    return r;
}

// np.sum over term(0..n-1) in numpy's pairwise order, any n.  numpy's recursion applies the same rule to every
// subtree, so each subtree of <= SQK_HEAP_MAX_N elements is summed by the parallel routine above and the few levels
// above are walked depth-first and folded with a depth stack -- identically in every thread of the group (the
// control flow depends on n only), so the collective calls inside stay converged.
template <int NT, class Term>
__device__ __noinline__ double stats_sum_long(Term term, int n, StatsShared &sh)
{
    int st_off[SQK_TREE_DEPTH], st_len[SQK_TREE_DEPTH], st_dep[SQK_TREE_DEPTH], cs_dep[SQK_TREE_DEPTH];
    double cs_val[SQK_TREE_DEPTH];
    int sp = 1, csp = 0;
    st_off[0] = 0; st_len[0] = n; st_dep[0] = 0;
    while (sp > 0) {
        sp--;
        const int off = st_off[sp], len = st_len[sp];
        int dep = st_dep[sp];
        if (len <= SQK_HEAP_MAX_N) {
            auto sub = [term, off](int q) -> double { return term(off + q); };
            double v = stats_pairwise_heap<NT>(sub, len, sh);
            while (csp > 0 && cs_dep[csp - 1] == dep) {          // sibling on the stack: left + right
                v = __dadd_rn(cs_val[csp - 1], v);
                dep--; csp--;
            }
            cs_val[csp] = v; cs_dep[csp] = dep; csp++;
        } else {
            int h = len / 2;
            h -= h % 8;
            st_off[sp] = off + h; st_len[sp] = len - h; st_dep[sp] = dep + 1; sp++;   // right, popped later
            st_off[sp] = off; st_len[sp] = h; st_dep[sp] = dep + 1; sp++;             // left, popped next
        }
    }
    return cs_val[0];
}

template <int NT, class Term>
__device__ __forceinline__ double stats_sum(Term term, int n, StatsShared &sh)
{
    // the long-read walk keeps its little stacks in local memory: out of line, so the common path has no frame
    return n <= SQK_HEAP_MAX_N ? stats_pairwise_heap<NT>(term, n, sh) : stats_sum_long<NT>(term, n, sh);
}

// direct-histogram bins are padded by one word per 32 so that the per-thread runs of consecutive bins in the
// prefix scan fall into different shared-memory banks
#define HB(b) ((b) + ((b) >> 5))

// ---- order statistics from a direct histogram of the raw values (window of <= SQK_HIST_BINS values) ----
// After stats_histogram, sh.hist[b] = number of staged samples with value <= base + b (inclusive prefix).
template <int NT>
__device__ void stats_histogram(const int16_t *stage, int n, int base, StatsShared &sh)
{
    const int tid = threadIdx.x % NT, lane = tid & 31, warp = tid >> 5;
    for (int b = tid; b < SQK_HIST_BINS + SQK_HIST_BINS / 32; b += NT) sh.hist[b] = 0;
    stats_sync<NT>();
    for (int i = tid; i < n; i += NT) atomicAdd(&sh.hist[HB((int)stage[i] - base)], 1u);
    stats_sync<NT>();
    constexpr int PER = SQK_HIST_BINS / NT;      // consecutive bins per thread
    uint32_t run = 0;
    for (int q = 0; q < PER; q++) run += sh.hist[HB(tid * PER + q)];
    uint32_t incl = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(SQK_FULL_MASK, incl, d);
        if (lane >= d) incl += t;
    }
    if (NT > 32) {
        if (lane == 31) sh.scan_part[warp] = incl;
        stats_sync<NT>();
    }
    uint32_t acc = incl - run;
    if (NT > 32)
        for (int w = 0; w < warp; w++) acc += sh.scan_part[w];
    for (int q = 0; q < PER; q++) {
        acc += sh.hist[HB(tid * PER + q)];
        sh.hist[HB(tid * PER + q)] = acc;
    }
    stats_sync<NT>();
}

// smallest bin b with prefix[b] > rank  (rank-th smallest value = base + b)
__device__ __forceinline__ int stats_hist_select(const StatsShared &sh, int nbins, int rank)
{
    int lo = 0, hi = nbins - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (sh.hist[HB(mid)] > (uint32_t)rank) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// number of staged samples v with |2v - med2| <= D
__device__ __forceinline__ int stats_hist_within(const StatsShared &sh, int base, int nbins, int med2, int D)
{
    // v >= ceil((med2 - D) / 2),  v <= floor((med2 + D) / 2)   (floor division on possibly negative numbers)
    const int a = med2 - D, b = med2 + D;
    const int vlo = (a >= 0) ? (a + 1) / 2 : -((-a) / 2);
    const int vhi = (b >= 0) ? b / 2 : -((-b + 1) / 2);
    int ilo = vlo - base, ihi = vhi - base;
    if (ihi >= nbins) ihi = nbins - 1;
    if (ilo < 0) ilo = 0;
    if (ihi < ilo) return 0;
    return (int)(sh.hist[HB(ihi)] - (ilo > 0 ? sh.hist[HB(ilo - 1)] : 0u));
}

// rank-th smallest doubled distance |2v - med2|
__device__ __forceinline__ int stats_hist_mad(const StatsShared &sh, int base, int nbins, int med2, int rank)
{
    int lo = 0, hi = 4 * SQK_HIST_BINS;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (stats_hist_within(sh, base, nbins, med2, mid) > rank) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// rank-th smallest (0-based) of key(i), i < n, keys < 512*256.  Two-pass radix select (wide windows).
template <int NT, class KeyFn>
__device__ uint32_t stats_select(KeyFn key, int n, int rank, StatsShared &sh)
{
    const int tid = threadIdx.x % NT;
    uint32_t prefix = 0;
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
        for (int b = tid; b < SQK_RADIX_BINS; b += NT) sh.hist[b] = 0;
        stats_sync<NT>();
        for (int i = tid; i < n; i += NT) {
            const uint32_t kv = key(i);
            if (pass == 0) atomicAdd(&sh.hist[kv >> 8], 1u);
            else if ((kv >> 8) == prefix) atomicAdd(&sh.hist[kv & 255u], 1u);
        }
        stats_sync<NT>();
        if (tid < 32) {
            const int per = SQK_RADIX_BINS / 32;
            uint32_t mine = 0;
            for (int b = 0; b < per; b++) mine += sh.hist[tid * per + b];
            uint32_t incl = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(SQK_FULL_MASK, incl, d);
                if (tid >= d) incl += t;
            }
            uint32_t cum = incl - mine;
            if ((uint32_t)rank >= cum && (uint32_t)rank < incl) {
                for (int b = 0; b < per; b++) {
                    const uint32_t h = sh.hist[tid * per + b];
                    if ((uint32_t)rank < cum + h) { sh.sel[0] = tid * per + b; sh.sel[1] = rank - cum; break; }
                    cum += h;
                }
            }
        }
        stats_sync<NT>();
        if (pass == 0) { prefix = sh.sel[0]; rank = (int)sh.sel[1]; }
        else prefix = (prefix << 8) | sh.sel[0];
        stats_sync<NT>();
    }
    return prefix;
}

// smallest d in [-32768, 32767] with pa(d) > limit  (32768 if none);  pa is monotone in d
__device__ __forceinline__ int stats_first_above(double limit, double off, double unit)
{
    int lo = -32768, hi = 32768;
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if (sqk_pa_value(mid, off, unit) > limit) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// largest d in [-32768, 32767] with pa(d) < limit  (-32769 if none)
__device__ __forceinline__ int stats_last_below(double limit, double off, double unit)
{
    int lo = -32769, hi = 32767;
    while (lo < hi) {
        const int mid = lo + ((hi - lo + 1) >> 1);
        if (sqk_pa_value(mid, off, unit) < limit) lo = mid; else hi = mid - 1;
    }
    return lo;
}

template <int NT> struct StatsCta { static constexpr int threads = NT > SQK_STATS_THREADS ? NT : SQK_STATS_THREADS; };

template <int NT>
__global__ void __launch_bounds__(StatsCta<NT>::threads, NT > SQK_STATS_THREADS ? 2 : 6) sqk_stats_kernel(const StatsArgs a)
{
    constexpr int GROUPS = StatsCta<NT>::threads / NT;
    constexpr int WARPS = NT / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x % NT, gi = threadIdx.x / NT, lane = tid & 31, warp = tid >> 5;
    const size_t group_bytes = ((sizeof(StatsShared) + 15) & ~(size_t)15) + (size_t)a.cap * sizeof(int16_t);
    unsigned char *mine = smem_raw + gi * group_bytes;
    StatsShared &sh = *reinterpret_cast<StatsShared *>(mine);
    int16_t *smem_stage = reinterpret_cast<int16_t *>(mine + ((sizeof(StatsShared) + 15) & ~(size_t)15));
    int64_t alloc_lo = a.alloc_lo, alloc_hi = a.alloc_hi;
    resolve_bounds(a.offsets, a.read0, a.n_reads, alloc_lo, alloc_hi);
    const bool pa_mode = (a.mode == SQK_STATS_SEGMENTER && a.pa_offset != nullptr);
    const int64_t slot = (int64_t)blockIdx.x * GROUPS + gi;

    const int64_t n_items = a.list ? (int64_t)*a.n_list : a.n_reads;
    for (int64_t item = slot; item < n_items; item += (int64_t)gridDim.x * GROUPS) {
        const int64_t i = a.list ? (int64_t)a.list[item] : item;
        const int64_t r = a.read0 + i;
        const int64_t begin = a.offsets[r];
        int64_t len = a.offsets[r + 1] - begin;
        if (a.mode == SQK_STATS_SEGMENTER) len = sqk_truncate_len(len, a.num);
        const int64_t end = begin + len;
        const bool staged = (a.mode != SQK_STATS_NONE);
        const bool in_smem = (len <= a.cap);
        if (staged && !in_smem && (a.gstage == nullptr || len > a.gstage_stride)) {
            // longer than the max_read_len the caller declared: no staging row was provisioned for it
            if (tid == 0) {
                ReadStats bad;
                bad.center = 0.0; bad.scale = 1.0; bad.n_kept = 0; bad.flags = SQK_FLAG_TOO_LONG | a.extra_flags;
                bad.seg_lo = 0; bad.seg_hi = -1; bad.out_lo = 1; bad.out_hi = 0;
                a.stats[i] = bad;
                if (a.n_kept_out) a.n_kept_out[i] = -1;
            }
            continue;
        }
        int16_t *stage = in_smem ? smem_stage : a.gstage + slot * a.gstage_stride;

        // ---- outlier window on the raw sample (inclusive) ------------------------------------
        int out_lo = a.lo + 1, out_hi = a.hi - 1;
        double pa_off = 0.0, pa_unit = 1.0;
        if (pa_mode) {
            // lim_low < pA(d) < lim_hi  <=>  out_lo <= d <= out_hi   (pA is monotone in d)
            pa_off = a.pa_offset[r]; pa_unit = a.pa_scale[r];
            if (pa_unit > 0.0) {
                out_lo = stats_first_above((double)a.lo, pa_off, pa_unit);
                out_hi = stats_last_below((double)a.hi, pa_off, pa_unit);
            } else {
                out_lo = 1; out_hi = 0;   // unusable calibration: keep nothing, report no segments
            }
        }

        // ---- pass over HBM: filter, compact, integer sum ------------------------------------
        long long sum = 0;
        int total = 0;
        const bool window_ok = out_hi >= out_lo;
        const unsigned span = (unsigned)(out_hi - out_lo);
        const int64_t blk0 = aligned_block_start(a.base, begin);
        constexpr int U = 4;                      // 16-byte loads in flight per thread
        int parity = 0;
        for (int64_t cb = blk0; cb < end; cb += (int64_t)NT * 8 * U, parity ^= 1) {
            // Block b = u*NT + tid of this iteration holds samples cb + 8b .. +8; compacted positions follow block order.
            Samples8 sv[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int64_t blk = cb + ((int64_t)u * NT + tid) * 8;
                if (blk < end && blk + 8 > begin) sv[u] = load_block8(a.base, blk, alloc_lo, alloc_hi);
            }
            unsigned keep[U];
            int incl[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int64_t blk = cb + ((int64_t)u * NT + tid) * 8;
                const Samples8 s = sv[u];
                keep[u] = 0;
                if (blk < end && blk + 8 > begin) {
                    // one unsigned compare per sample for the window; edge blocks mask the samples outside the read
                    int lsum = 0;
                    unsigned kp = 0;
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        const int v = s.get(e);
                        if ((unsigned)(v - out_lo) <= span) { kp |= 1u << e; lsum += v; }
                    }
                    if (blk < begin || blk + 8 > end) {
                        const int first = blk < begin ? (int)(begin - blk) : 0;
                        const int last = blk + 8 > end ? (int)(end - blk) : 8;          // valid samples: [first, last)
                        const unsigned valid = ((1u << last) - 1u) & ~((1u << first) - 1u);
                        const unsigned drop = kp & ~valid;
                        kp &= valid;
#pragma unroll
                        for (int e = 0; e < 8; e++)
                            if (drop & (1u << e)) lsum -= s.get(e);
                    }
                    if (!window_ok) { kp = 0; lsum = 0; }
                    sum += lsum;
                    keep[u] = kp;
                }
                // inclusive scan of the kept counts over the warp, one per block row u
                int inc = __popc(keep[u]);
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int t = __shfl_up_sync(SQK_FULL_MASK, inc, d);
                    if (lane >= d) inc += t;
                }
                incl[u] = inc;
                if (NT > 32 && lane == 31) sh.scan_tot[parity][u][warp] = inc;
            }
            // ONE barrier per iteration: the warp totals of all U block rows become visible together (the buffer
            // alternates with the iteration, so the next iteration's writes cannot overtake this iteration's reads)
            if (NT > 32) __syncthreads();
#pragma unroll
            for (int u = 0; u < U; u++) {
                int pos = total + incl[u] - __popc(keep[u]), all;
                if (NT == 32) {
                    all = __shfl_sync(SQK_FULL_MASK, incl[u], 31);
                } else {
                    all = 0;
#pragma unroll
                    for (int w = 0; w < WARPS; w++) {
                        const int t = sh.scan_tot[parity][u][w];
                        if (w < warp) pos += t;
                        all += t;
                    }
                }
                const unsigned kp = keep[u];
                const Samples8 s = sv[u];
                if (staged && kp) {
                    if (in_smem) {
                        // shared-memory window: a fully kept block landing on an even position goes out as four words
                        if (kp == 0xffu && !(pos & 1)) {
                            uint32_t *w32 = reinterpret_cast<uint32_t *>(smem_stage + pos);
                            w32[0] = (uint32_t)s.v.x; w32[1] = (uint32_t)s.v.y; w32[2] = (uint32_t)s.v.z; w32[3] = (uint32_t)s.v.w;
                        } else {
#pragma unroll
                            for (int e = 0; e < 8; e++)
                                if (kp & (1u << e)) smem_stage[pos++] = (int16_t)s.get(e);
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 8; e++)
                            if (kp & (1u << e)) stage[pos++] = (int16_t)s.get(e);
                    }
                }
                total += all;
            }
        }
        const int n = total;

#pragma unroll
        for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(SQK_FULL_MASK, sum, d);
        long long tot_sum = sum;
        if (NT > 32) {
            if (lane == 0) sh.sum_part[warp] = (unsigned long long)sum;
            __syncthreads();
            tot_sum = 0;
#pragma unroll
            for (int w = 0; w < WARPS; w++) tot_sum += (long long)sh.sum_part[w];
        }
        stats_sync<NT>();   // staged samples visible to the whole group

        ReadStats out;
        out.center = 0.0; out.scale = 1.0; out.n_kept = n; out.flags = a.extra_flags; out.seg_lo = 0; out.seg_hi = -1;
        out.out_lo = out_lo; out.out_hi = out_hi;

        double sd = 0.0;
        if (n > 0 && (a.mode == SQK_STATS_ZSCALE || a.mode == SQK_STATS_SEGMENTER)) {
            if (!pa_mode) {
                // integer samples: the sum is exact in any order
                const double mean = __ddiv_rn((double)tot_sum, (double)n);
                auto sq = [stage, mean](int q) -> double { const double d = __dsub_rn((double)stage[q], mean); return __dmul_rn(d, d); };
                sd = __dsqrt_rn(__ddiv_rn(stats_sum<NT>(sq, n, sh), (double)n));
            } else {
                // pA samples are not integers: np.std's mean is itself a pairwise sum
                auto val = [stage, pa_off, pa_unit](int q) -> double { return sqk_pa_value((int)stage[q], pa_off, pa_unit); };
                const double mean = __ddiv_rn(stats_sum<NT>(val, n, sh), (double)n);
                auto sq = [stage, pa_off, pa_unit, mean](int q) -> double {
                    const double d = __dsub_rn(sqk_pa_value((int)stage[q], pa_off, pa_unit), mean);
                    return __dmul_rn(d, d);
                };
                sd = __dsqrt_rn(__ddiv_rn(stats_sum<NT>(sq, n, sh), (double)n));
            }
            if (a.mode == SQK_STATS_ZSCALE) {
                if (sd == 0.0) sd = 1.0;          // sklearn _handle_zeros_in_scale
                out.center = __ddiv_rn((double)tot_sum, (double)n); out.scale = sd;
            }
        }
        if (n > 0 && (a.mode == SQK_STATS_MEDMAD || a.mode == SQK_STATS_SEGMENTER)) {
            const int nbins = out_hi - out_lo + 1;
            const bool direct = nbins <= SQK_HIST_BINS;      // narrow window: one histogram serves median and MAD
            auto key_x = [stage](int q) -> uint32_t { return (uint32_t)((int)stage[q] + 32768); };
            int lo_v, hi_v;   // the two middle order statistics (equal for odd n)
            if (direct) {
                stats_histogram<NT>(stage, n, out_lo, sh);
                lo_v = out_lo + stats_hist_select(sh, nbins, (n - 1) / 2);
                hi_v = (n & 1) ? lo_v : out_lo + stats_hist_select(sh, nbins, n / 2);
            } else if (n & 1) {
                lo_v = hi_v = (int)stats_select<NT>(key_x, n, (n - 1) / 2, sh) - 32768;
            } else {
                lo_v = (int)stats_select<NT>(key_x, n, n / 2 - 1, sh) - 32768;
                hi_v = (int)stats_select<NT>(key_x, n, n / 2, sh) - 32768;
            }
            const int med2 = lo_v + hi_v;   // 2 * median of the raw integers
            if (a.mode == SQK_STATS_MEDMAD) {
                const double median = (double)med2 * 0.5;
                auto key_d = [stage, med2](int q) -> uint32_t {
                    const int d = 2 * (int)stage[q] - med2;
                    return (uint32_t)(d < 0 ? -d : d);
                };
                double mad;
                if (direct) {
                    const int d0 = stats_hist_mad(sh, out_lo, nbins, med2, (n - 1) / 2);
                    const int d1 = (n & 1) ? d0 : stats_hist_mad(sh, out_lo, nbins, med2, n / 2);
                    mad = (double)(d0 + d1) * 0.25;
                } else if (n & 1) {
                    mad = (double)stats_select<NT>(key_d, n, (n - 1) / 2, sh) * 0.5;
                } else {
                    const uint32_t d0 = stats_select<NT>(key_d, n, n / 2 - 1, sh);
                    const uint32_t d1 = stats_select<NT>(key_d, n, n / 2, sh);
                    mad = (double)(d0 + d1) * 0.25;
                }
                const double scaled = __dmul_rn(mad, 1.4826);
                out.center = median; out.scale = scaled;
                if (scaled == 0.0) out.flags |= SQK_FLAG_DEGENERATE;
            } else {
                double median;
                if (!pa_mode) median = (double)med2 * 0.5;
                else if (n & 1) median = sqk_pa_value(lo_v, pa_off, pa_unit);
                else median = __ddiv_rn(__dadd_rn(sqk_pa_value(lo_v, pa_off, pa_unit), sqk_pa_value(hi_v, pa_off, pa_unit)), 2.0);
                const double spread = __dmul_rn(sd, a.std_scale);
                const double top = __dadd_rn(median, spread);
                const double bot = __dsub_rn(median, spread);
                if (!pa_mode) {
                    // integer x:  x < top  <=>  x <= ceil(top)-1 ;  x > bot  <=>  x >= floor(bot)+1
                    const double hi_d = fmin(fmax(ceil(top) - 1.0, -40000.0), 40000.0);
                    const double lo_d = fmin(fmax(floor(bot) + 1.0, -40000.0), 40000.0);
                    out.seg_hi = (top == top) ? (int)hi_d : -40000;   // NaN threshold: nothing is in range
                    out.seg_lo = (bot == bot) ? (int)lo_d : 40000;
                } else {
                    // pA(d) < top  <=>  d <= last_below(top);  pA(d) > bot  <=>  d >= first_above(bot)
                    out.seg_hi = (top == top) ? stats_last_below(top, pa_off, pa_unit) : -40000;
                    out.seg_lo = (bot == bot) ? stats_first_above(bot, pa_off, pa_unit) : 40000;
                }
                out.center = top; out.scale = bot;
            }
        }

        if (a.mode == SQK_STATS_ADAPTER) {
            // python slice sig[t_start:t_end] of the post-outlier signal; an empty slice gives NaN statistics
            int s0 = a.t_start < n ? a.t_start : n, s1 = a.t_end < n ? a.t_end : n;
            if (s0 < 0) s0 = 0;
            if (s1 < s0) s1 = s0;
            const int ns = s1 - s0;
            out.seg_lo = -40000; out.seg_hi = -40001;            // nothing is in range
            out.center = __longlong_as_double(0x7ff8000000000000LL); out.scale = out.center;
            if (ns > 0) {
                const int16_t *sl = stage + s0;
                long long part = 0;
                for (int q = tid; q < ns; q += NT) part += sl[q];
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(SQK_FULL_MASK, part, d);
                long long slice_sum = part;
                if (NT > 32) {
                    stats_sync<NT>();
                    if (lane == 0) sh.sum_part[warp] = (unsigned long long)part;
                    stats_sync<NT>();
                    slice_sum = 0;
#pragma unroll
                    for (int w = 0; w < WARPS; w++) slice_sum += (long long)sh.sum_part[w];
                    stats_sync<NT>();
                }
                const double mean = __ddiv_rn((double)slice_sum, (double)ns);
                auto sq = [sl, mean](int q) -> double { const double d = __dsub_rn((double)sl[q], mean); return __dmul_rn(d, d); };
                const double sdev = __dsqrt_rn(__ddiv_rn(stats_sum<NT>(sq, ns, sh), (double)ns));
                const int nbins = out_hi - out_lo + 1;
                auto key_x = [sl](int q) -> uint32_t { return (uint32_t)((int)sl[q] + 32768); };
                int lo_v, hi_v;
                if (nbins <= SQK_HIST_BINS) {
                    stats_histogram<NT>(sl, ns, out_lo, sh);
                    lo_v = out_lo + stats_hist_select(sh, nbins, (ns - 1) / 2);
                    hi_v = (ns & 1) ? lo_v : out_lo + stats_hist_select(sh, nbins, ns / 2);
                } else if (ns & 1) {
                    lo_v = hi_v = (int)stats_select<NT>(key_x, ns, (ns - 1) / 2, sh) - 32768;
                } else {
                    lo_v = (int)stats_select<NT>(key_x, ns, ns / 2 - 1, sh) - 32768;
                    hi_v = (int)stats_select<NT>(key_x, ns, ns / 2, sh) - 32768;
                }
                const double median = (double)(lo_v + hi_v) * 0.5;
                const double top = __dadd_rn(median, __dmul_rn(sdev, a.std_scale));
                // integer x:  x < top  <=>  x <= ceil(top)-1
                if (top == top) out.seg_hi = (int)fmin(fmax(ceil(top) - 1.0, -40000.0), 40000.0);
                out.center = top; out.scale = median;
            }
        }

        if (tid == 0) {
            a.stats[i] = out;
            if (a.n_kept_out) a.n_kept_out[i] = n;
        }
        stats_sync<NT>();
    }
}
