// sqk_rollmean.cuh -- the rolling-mean adapter finder of dRNA_segmenter.py's TSV branch (dRNA_segmenter.py:272-326):
//
//     sig = scale_outliers(...)                         :278, :331-334   (0 < s < 1200)
//     t   = pd.Series(sig).rolling(window=w).mean()     :281-282         (w = 2000, :81)
//     bot = t.mean() - t.std() * 0.5                    :283-287
//     run detector on t < bot / t > bot, merge rule, first segment with lo_thresh <= b - a <= hi_thresh   :291-323
//
// One CTA (256 threads) per read, persistent over the reads of the launch.
//   1. one pass over the raw int16 read (16-byte loads): outlier filter + block-wide prefix sums; P[c] = sum of the first c
//      kept samples (int32, wrapping: only window sums are ever formed) goes to a global scratch row.
//   2. pandas' rolling mean of integer samples is the exact window sum divided by w, rounded once: pandas keeps the
//      window sum with Kahan-compensated adds and removes, and for integers below 2^53 every partial result is exact, so
//      the compensation terms stay 0 (oracle/sqk_oracle.c restates the Kahan form and is pinned on real pandas).  So
//      t[i] = fl((P[i+1] - P[i+1-w]) / w) for i >= w-1, NaN in front.
//   3. t.mean() / t.std() are pandas nanops: NaN slots become 0.0 and the sums run over ALL n slots in numpy's pairwise
//      order -- stats_sum (sqk_stats.cuh) with term(i) evaluated on the fly from P; avg = sum / count,
//      var = sum((avg - t)^2) / (count - 1).
//   4. t[i] < bot and t[i] > bot are turned into integer tests on the window sum (fl(S / w) is monotone in S; the two
//      integer thresholds are found by evaluating the division on 32 neighbouring candidates), so the detector pass needs no
//      division: every thread tests one position, warp ballots give one "below" and one "above" bit mask per 32 positions.
//   5. warp 0 walks the masks 32 words at a time: stretches without an "above" bit (or without a "below" bit) are taken in
//      one step, the rest event by event with find-first-set; the list of segments is never stored -- an entry is final as
//      soon as the next one is appended, and only the first qualifying one is reported.
#pragma once
#include "sqk_stats.cuh"

#define SQK_RM_THREADS 256

struct RollmeanArgs {
    const int16_t *base;      // base[i] = absolute sample i
    int64_t alloc_lo, alloc_hi;
    const int64_t *offsets;   // absolute
    int64_t read0;
    int n_reads;
    int lo, hi;               // outlier window, exclusive
    int w, seg_dist, lo_thresh, hi_thresh, shift;
    double std_factor;
    int32_t *P;               // scratch [grid][p_stride]
    int64_t p_stride;
    uint32_t *masks;          // scratch [grid][2][m_stride]
    int64_t m_stride;
    int32_t *segs;            // [n_reads][2]
    int32_t *found;           // [n_reads]
};

struct RmShared {
    StatsShared st;
    long long wsum[8];
    int wcnt[8];
    int thr[2];
};

struct RmDetector {
    bool begin, have_last, found;
    int start, end, last_a, last_b, x, y;
    int seg_dist, lo_thresh, hi_thresh, shift;

    __device__ __forceinline__ void test_last()
    {
        const int d = last_b - last_a;
        if (!found && !(d > hi_thresh) && !(d < lo_thresh)) { found = true; x = last_a - shift; y = last_b - shift; }
    }
    __device__ __forceinline__ void close()
    {
        if (have_last && start - last_b < seg_dist) {
            last_b = end;
        } else {
            if (have_last) test_last();        // the previous entry can no longer change
            last_a = start; last_b = end; have_last = true;
        }
        start = 0; end = 0; begin = false;
    }
    // events of one 32-position word (B = below bits, A = above bits, disjoint), first position = pos0
    __device__ __forceinline__ void word(uint32_t B, uint32_t A, int pos0)
    {
        uint32_t ev = B | A;
        while (ev) {
            const int p = __ffs((int)ev) - 1;
            uint32_t rest;
            if ((B >> p) & 1u) {
                const uint32_t a_up = A & ev;                     // the next "above" bit ends this run of "below" bits
                const int q = a_up ? __ffs((int)a_up) - 1 : 32;
                const uint32_t below_q = q < 32 ? ((1u << q) - 1u) : 0xffffffffu;
                const int last = 31 - __clz((int)(B & ev & below_q));
                if (!begin) { start = pos0 + p; begin = true; if (last > p) end = pos0 + last; }
                else end = pos0 + last;
                rest = ~below_q;
            } else {
                if (begin) close();
                const uint32_t b_up = B & ev;                     // further "above" bits in front of the next "below" bit do nothing
                const int q = b_up ? __ffs((int)b_up) - 1 : 32;
                rest = q < 32 ? ~((1u << q) - 1u) : 0u;
            }
            ev &= rest;
        }
    }
};

__global__ void __launch_bounds__(SQK_RM_THREADS) sqk_rollmean_kernel(const RollmeanArgs a)
{
    extern __shared__ __align__(16) unsigned char rm_smem[];
    RmShared &sh = *reinterpret_cast<RmShared *>(rm_smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int64_t alloc_lo = a.alloc_lo, alloc_hi = a.alloc_hi;
    resolve_bounds(a.offsets, a.read0, a.n_reads, alloc_lo, alloc_hi);
    int32_t *P = a.P + (int64_t)blockIdx.x * a.p_stride;
    uint32_t *Bm = a.masks + (int64_t)blockIdx.x * 2 * a.m_stride, *Am = Bm + a.m_stride;
    const unsigned span = (unsigned)(a.hi - 1 - (a.lo + 1));
    const bool window_ok = a.hi - 1 >= a.lo + 1;
    const int w = a.w;

    for (int r = blockIdx.x; r < a.n_reads; r += gridDim.x) {
        const int64_t begin = a.offsets[a.read0 + r], end = a.offsets[a.read0 + r + 1];
        // ---- 1. outlier filter + prefix sums of the kept samples -----------------------------------------------------
        int n = 0;
        long long run_sum = 0;
        if (tid == 0) P[0] = 0;
        const int64_t blk0 = aligned_block_start(a.base, begin);
        for (int64_t cb = blk0; cb < end; cb += 8ll * SQK_RM_THREADS) {
            const int64_t ub = cb + 8ll * tid;
            int v[8];
            int cnt = 0, sum = 0;
            unsigned keep = 0;
            if (ub < end && window_ok) {
                const Samples8 sv = load_block8(a.base, ub, alloc_lo, alloc_hi);
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    v[e] = sv.get(e);
                    const bool k = ub + e >= begin && ub + e < end && (unsigned)(v[e] - (a.lo + 1)) <= span;
                    keep |= k ? 1u << e : 0u;
                    cnt += k ? 1 : 0;
                    sum += k ? v[e] : 0;
                }
            }
            // block-wide exclusive scan of (cnt, sum)
            int icnt = cnt;
            long long isum = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int tc = __shfl_up_sync(SQK_FULL_MASK, icnt, d);
                const long long ts = __shfl_up_sync(SQK_FULL_MASK, isum, d);
                if (lane >= d) { icnt += tc; isum += ts; }
            }
            if (lane == 31) { sh.wcnt[warp] = icnt; sh.wsum[warp] = isum; }
            __syncthreads();
            int ecnt = icnt - cnt, tot_cnt = 0;
            long long esum = isum - sum, tot_sum = 0;
#pragma unroll
            for (int q = 0; q < SQK_RM_THREADS / 32; q++) {
                if (q < warp) { ecnt += sh.wcnt[q]; esum += sh.wsum[q]; }
                tot_cnt += sh.wcnt[q]; tot_sum += sh.wsum[q];
            }
            if (keep) {
                int pos = n + ecnt;
                long long s = run_sum + esum;
#pragma unroll
                for (int e = 0; e < 8; e++)
                    if (keep & (1u << e)) { s += v[e]; pos++; P[pos] = (int32_t)s; }
            }
            n += tot_cnt; run_sum += tot_sum;
            __syncthreads();                                   // wcnt / wsum reusable
        }
        __threadfence_block();
        __syncthreads();                                       // P visible to the whole CTA

        // ---- 2./3. mean and standard deviation of the rolling mean (pandas nanops over all n slots) -------------------
        const int count = n - (w - 1);
        bool have_bot = false;
        double bot = 0.0;
        if (count >= 2) {                                      // count <= 1: std is NaN, so is bot: nothing compares
            const double wd = (double)w;
            auto tval = [P, w, wd](int i) -> double {
                return i >= w - 1 ? __ddiv_rn((double)(P[i + 1] - P[i + 1 - w]), wd) : 0.0;
            };
            const double the_sum = stats_sum<SQK_RM_THREADS>(tval, n, sh.st);
            const double avg = __ddiv_rn(the_sum, (double)count);
            auto sq = [P, w, wd, avg](int i) -> double {
                if (i < w - 1) return 0.0;
                const double d = __dsub_rn(avg, __ddiv_rn((double)(P[i + 1] - P[i + 1 - w]), wd));
                return __dmul_rn(d, d);
            };
            const double ssq = stats_sum<SQK_RM_THREADS>(sq, n, sh.st);
            const double sd = __dsqrt_rn(__ddiv_rn(ssq, (double)(count - 1)));
            bot = __dsub_rn(avg, __dmul_rn(sd, a.std_factor));
            have_bot = bot == bot;
        }
        // ---- 4. integer forms of the two comparisons:  t < bot <=> S < thr[0] ;  t > bot <=> S >= thr[1] -----------------
        if (have_bot) {
            if (warp == 0) {
                const double wd = (double)w;
                // fl(S / w) is monotone in S; bot * w is within a few units of both thresholds
                const double guess = fmin(fmax(floor(__dmul_rn(bot, wd)), -2147483000.0), 2147483000.0);
                const long long c0 = (long long)guess - 15;
                const long long cand = c0 + lane;
                const double f = __ddiv_rn((double)cand, wd);
                const unsigned ge = __ballot_sync(SQK_FULL_MASK, f >= bot);    // monotone: a suffix of the lanes
                const unsigned gt = __ballot_sync(SQK_FULL_MASK, f > bot);
                if (lane == 0) {
                    // (the guess is off by far less than 15: both masks are non-trivial suffixes; be safe anyway)
                    const long long t0 = ge ? c0 + (__ffs((int)ge) - 1) : c0 + 32;
                    const long long t1 = gt ? c0 + (__ffs((int)gt) - 1) : c0 + 32;
                    sh.thr[0] = (int)(t0 < -2147483647ll ? -2147483647ll : (t0 > 2147483647ll ? 2147483647ll : t0));
                    sh.thr[1] = (int)(t1 < -2147483647ll ? -2147483647ll : (t1 > 2147483647ll ? 2147483647ll : t1));
                }
            }
            __syncthreads();
            const int thr_b = sh.thr[0], thr_a = sh.thr[1];
            const int nw = (n + 31) >> 5;
            for (int i0 = warp * 32; i0 < nw * 32; i0 += SQK_RM_THREADS) {
                const int i = i0 + lane;
                bool below = false, above = false;
                if (i < n && i >= w - 1) {
                    const int S = P[i + 1] - P[i + 1 - w];
                    below = S < thr_b; above = S >= thr_a;
                }
                const unsigned bm = __ballot_sync(SQK_FULL_MASK, below), am = __ballot_sync(SQK_FULL_MASK, above);
                if (lane == 0) { Bm[i0 >> 5] = bm; Am[i0 >> 5] = am; }
            }
            __threadfence_block();
        }
        __syncthreads();

        // ---- 5. the run detector on the masks (warp 0; every lane carries the same state) -----------------------------
        if (warp == 0) {
            RmDetector dt;
            dt.begin = false; dt.have_last = false; dt.found = false;
            dt.start = 0; dt.end = 0; dt.last_a = 0; dt.last_b = 0; dt.x = 0; dt.y = 0;
            dt.seg_dist = a.seg_dist; dt.lo_thresh = a.lo_thresh; dt.hi_thresh = a.hi_thresh; dt.shift = a.shift;
            if (have_bot) {
                const int nw = (n + 31) >> 5;
                for (int g0 = 0; g0 < nw && !dt.found; g0 += 32) {
                    const int wi = g0 + lane;
                    const uint32_t B = wi < nw ? Bm[wi] : 0u, A = wi < nw ? Am[wi] : 0u;
                    const unsigned hasB = __ballot_sync(SQK_FULL_MASK, B != 0u), hasA = __ballot_sync(SQK_FULL_MASK, A != 0u);
                    if (!hasA) {
                        if (hasB) {
                            // only "below" positions in these 1024: first / last of them
                            const int lf = __ffs((int)hasB) - 1, ll = 31 - __clz((int)hasB);
                            const uint32_t Bf = __shfl_sync(SQK_FULL_MASK, B, lf), Bl = __shfl_sync(SQK_FULL_MASK, B, ll);
                            const int first = 32 * (g0 + lf) + __ffs((int)Bf) - 1, last = 32 * (g0 + ll) + 31 - __clz((int)Bl);
                            if (!dt.begin) { dt.start = first; dt.begin = true; if (last > first) dt.end = last; }
                            else dt.end = last;
                        }
                    } else if (!hasB) {
                        if (dt.begin) dt.close();              // the first "above" position closes; the others do nothing
                    } else {
                        unsigned todo = hasB | hasA;
                        while (todo && !dt.found) {
                            const int l = __ffs((int)todo) - 1;
                            todo &= todo - 1;
                            dt.word(__shfl_sync(SQK_FULL_MASK, B, l), __shfl_sync(SQK_FULL_MASK, A, l), 32 * (g0 + l));
                        }
                    }
                }
                if (dt.have_last) dt.test_last();              // end of the read: the last entry is final, too
            }
            if (lane == 0) {
                a.segs[2 * (int64_t)r] = dt.found ? dt.x : 0;
                a.segs[2 * (int64_t)r + 1] = dt.found ? dt.y : 0;
                a.found[r] = dt.found ? 1 : 0;
            }
        }
        __syncthreads();                                       // scratch rows reusable
    }
}
