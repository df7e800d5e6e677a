// sqk_dtw_lb.cuh -- pass 1 of the exact two-pass plan (sqk_dtw_plan.cuh): a float32, cost-only, rounded-down
// lower bound of the last row of mlpy's subsequence-DTW cost matrix (MotifSeq.py:437), fused with the same
// int16 -> outlier filter -> normalise front end as the exact kernel.  3 instructions per cell
// (FADD.RZ, FMNMX3, FADD.RM) instead of ~12; no start pointers, no float64 in the recurrence (the derivation of
// the bound -- the U recurrence with the free-start row fed j*w -- is in sqk_dtw_plan.cuh).
//
// Same mapping as sqk_dtw_kernel: L lanes per read, K motif rows per lane in registers, skewed wavefront (two
// signal columns per step: two dependency chains per lane), one __shfl_up per column, normalised samples staged
// in a per-group shared-memory ring.  The lane that owns the last motif row watches U against a per-block
// threshold (one compare and a branch per two columns; candidates are rare) and keeps the candidate clusters;
// refill checkpoints (kept-sample count in front of every refill) let it translate a cluster into the raw
// position its exact window starts at.  At the end of a read it appends one DtwJob per cluster to the job list
// consumed by sqk_dtw_kernel<JOBS>.
#pragma once
#include "sqk_common.cuh"
#include "sqk_dtw_plan.cuh"

#define SQK_LB_WARPS 4
#define SQK_LB_THREADS (SQK_LB_WARPS * 32)
#ifdef SQK_LB_MINB_OVERRIDE     // experiments (sqk_ubench variants)
#define SQK_LB_MINB(K) SQK_LB_MINB_OVERRIDE
#else
#define SQK_LB_MINB(K) ((K) <= 10 ? 6 : ((K) <= 16 ? 5 : ((K) <= 20 ? 4 : 2)))
#endif

struct LbArgs {
    const int16_t *base;      // base[i] = absolute sample i
    int64_t alloc_lo, alloc_hi;
    const int64_t *offsets;   // absolute
    int64_t read0;
    int n_reads;
    const ReadStats *stats;   // [n_reads]
    const double *model;      // N motif points
    int N;
    int lo, hi;
    sqk_hit *hits;            // hits[i * hit_stride]: only written for empty / degenerate reads
    int hit_stride;
    unsigned int *counter;    // work queue head, zeroed before launch
    DtwJob *jobs;             // job list (capacity n_reads * SQK_LB_MAX_CLUSTERS)
    unsigned int *n_jobs;     // its length, zeroed before launch
    LbRead *reads;            // [n_reads]
    double xmax_abs;          // max |motif point|
    int W;                    // window columns in front of a cluster
    int W2;                   // ... of the second attempt (0: none)
    DtwJob *jobs2;            // [n_reads * SQK_LB_MAX_CLUSTERS]: the second-attempt job of cluster q of read r at r * MAX + q
    int short_len;            // reads with n_kept <= short_len skip pass 1 (one full-length job)
    int cols;                 // signal columns per wavefront step: 2, 4, or 0 = the launcher's rule
};

// Shared-memory ring access by 32-bit shared address.  Every group's ring is aligned to its size (RC*4 bytes), so
// advancing by one entry with wrap-around is one add and one bit-select.
__device__ __forceinline__ float lb_lds(unsigned addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
template <int RING_BYTES>
__device__ __forceinline__ unsigned lb_ring_next(unsigned addr)
{
    return (addr & ~(unsigned)(RING_BYTES - 1)) | ((addr + 4u) & (unsigned)(RING_BYTES - 1));
}

// State of the candidate watch of one read (meaningful in the lane that owns the last motif row).
struct LbWatch {
    float runmin, thr;     // running minimum of L and its candidate threshold
    float thr_u;           // thr + an upper bound of (j + N) * w over the current block of steps: the cheap test on U
                           // (-inf in every other lane)
    float aeps, bslack, w;
    float wstep;           // COLS * w, rounded down: what the free-start row advances by per wavefront step
    int n;                 // columns of the read
    int N;
};

__device__ __forceinline__ void lb_candidate(float u, int j, LbWatch &wt, LbClusters *cl, const int32_t *ck, int n_ref,
                                             int64_t cursor0, int ch, int W, int W2 = 0)
{
    if (u <= wt.thr_u && (unsigned)j < (unsigned)wt.n) {   // (stale ring entries past the end of the read never count)
        const float lj = sqk_lb_adjust(u, j, wt.N, wt.w);
        if (lj <= wt.thr) {
            LbScan sc; sc.ck = ck; sc.n_ref = n_ref; sc.cursor0 = cursor0; sc.ch = ch; sc.W = W; sc.W2 = W2;
            lbc_event(*cl, j, lj, wt.runmin, wt.thr, wt.aeps, wt.bslack, sc);
            // the threshold may have moved: keep the cheap test valid for the rest of this block and the next
            wt.thr_u = sqk_lb_thr_u(wt.thr, sqk_mul_ru((float)(j + wt.N + 2 * ch), wt.w));
        }
    }
}

// The rare path of the scan, OUT OF LINE: up to eight last-row values of consecutive columns from j0 on, in column order.
// Everything travels by value (the watch state comes back in a small struct), so that the hot loop's register allocation
// knows nothing of the cluster bookkeeping behind this call.
struct LbWatchUpd { float runmin, thr, thr_u; };
struct LbVals8 { float v[8]; };
static __device__ __noinline__ LbWatchUpd lb_candidates(LbVals8 vals, int count, int j0, float runmin, float thr, float thr_u, float aeps,
                                                 float bslack, float w, int n, int N, LbClusters *cl, const int32_t *ck, int n_ref,
                                                 int64_t cursor0, int ch, int W, int W2)
{
    LbWatch wt;
    wt.runmin = runmin; wt.thr = thr; wt.thr_u = thr_u; wt.aeps = aeps; wt.bslack = bslack; wt.w = w; wt.wstep = 0.0f; wt.n = n; wt.N = N;
#pragma unroll 1
    for (int q = 0; q < count; q++) lb_candidate(vals.v[q], j0 + q, wt, cl, ck, n_ref, cursor0, ch, W, W2);
    LbWatchUpd r; r.runmin = wt.runmin; r.thr = wt.thr; r.thr_u = wt.thr_u;
    return r;
}

// One wavefront step of one lane: column (t - l) of the U recurrence for this lane's K rows (sqk_dtw_plan.cuh).
// tf: what the free-start row holds in column t (sqk_lb_virtual).  raddr: shared address of this lane's ring entry.
template <int K, int L, bool RAGGED>
__device__ __forceinline__ void lb_step(const float (&ci)[K], float (&co)[K], const float (&x)[K], unsigned &raddr, int l,
                                        bool pass0, int t, float &tf, float &bot, float &prev_up, LbWatch &wt,
                                        LbClusters *cl, const int32_t *ck, int n_ref, int64_t cursor0, int W)
{
    float up = __shfl_up_sync(SQK_FULL_MASK, bot, 1, L);
    if (l == 0) up = tf;                               // free-start row: (a lower bound of) j*w in column j = t
    tf = sqk_lb_virtual_next(tf, wt.w);
    const float y = lb_lds(raddr);
    raddr = lb_ring_next<16 * L * 4>(raddr);
    float dg = prev_up;
    prev_up = up;
    float u = up;
#pragma unroll
    for (int k = 0; k < K; k++) {
        const float lf = ci[k];
        const float m = fminf(fminf(u, dg), lf);
        float nc = sqk_lb_cell(x[k], y, m);
        if (RAGGED && k == 0 && pass0) nc = up;        // pass-through slot (only when L*K != N)
        dg = lf;
        u = nc;
        co[k] = nc;
    }
    bot = u;
    if (bot <= wt.thr_u)                               // rare: a column of the last row that may be a candidate
        lb_candidate(bot, t - (L - 1), wt, cl, ck, n_ref, cursor0, 8 * L, W);
}

__device__ __forceinline__ float2 lb_lds2(unsigned addr)
{
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
template <int RING_BYTES>
__device__ __forceinline__ unsigned lb_ring_next2(unsigned addr)
{
    return (addr & ~(unsigned)(RING_BYTES - 1)) | ((addr + 8u) & (unsigned)(RING_BYTES - 1));
}

// Two columns per step: this lane's K rows of columns j = t - 2l (cells A) and j + 1 (cells B).  B_k needs A_k, A_{k-1}
// and B_{k-1}; A_k needs A_{k-1}: two dependency chains one cell apart, so the lane always has two independent
// FMNMX3 -> FADD pairs in flight (the one-column step is latency-bound on a single chain).  tf: the free-start
// row's value in column t.
template <int K, int L, bool RAGGED>
__device__ __forceinline__ void lb_step2(const float (&ci)[K], float (&co)[K], const float (&x)[K], unsigned &raddr, int l,
                                         bool pass0, int t, float &tf, float &bot_a, float &bot_b, float &prev_up_b,
                                         LbWatch &wt, LbClusters *cl, const int32_t *ck, int n_ref, int64_t cursor0, int W)
{
    float up_a = __shfl_up_sync(SQK_FULL_MASK, bot_a, 1, L);
    float up_b = __shfl_up_sync(SQK_FULL_MASK, bot_b, 1, L);
    // free-start row: a lower bound of j*w in column j (lane 0 is at columns t, t+1); both columns of the step take the
    // value of the first one (smaller is still a lower bound; it costs one w of tightness), one rounded-down add per step
    if (l == 0) { up_a = tf; up_b = tf; }
    tf = sqk_add_rd(tf, wt.wstep);
    const float2 y = lb_lds2(raddr);
    raddr = lb_ring_next2<16 * L * 4>(raddr);
    float dg_a = prev_up_b;                              // row above, column j-1
    prev_up_b = up_b;
    float ua = up_a, ub = up_b;
#pragma unroll
    for (int k = 0; k < K; k++) {
        const float lf = ci[k];                          // column j-1
        float a = sqk_lb_cell(x[k], y.x, fminf(fminf(ua, dg_a), lf));
        if (RAGGED && k == 0 && pass0) a = up_a;         // pass-through slot (only when L*K != N)
        float b = sqk_lb_cell(x[k], y.y, fminf(fminf(ub, ua), a));
        if (RAGGED && k == 0 && pass0) b = up_b;
        dg_a = lf;
        ua = a; ub = b;
        co[k] = b;
    }
    bot_a = ua; bot_b = ub;                              // (the caller tests the last row for candidates: once per two steps)
}

__device__ __forceinline__ float4 lb_lds4(unsigned addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
template <int RING_BYTES>
__device__ __forceinline__ unsigned lb_ring_next4(unsigned addr)
{
    return (addr & ~(unsigned)(RING_BYTES - 1)) | ((addr + 16u) & (unsigned)(RING_BYTES - 1));
}

// Four columns per step: this lane's K rows of columns j = t - 4l .. j + 3.  Cell (k, q) needs (k-1, q), (k-1, q-1) and
// (k, q-1): four dependency chains one cell apart; the per-step work that does not scale with the cells (shuffles per
// column aside: ring load and address, free-start row, loop control, the candidate test) is paid once per 4K cells.
template <int K, int L, bool RAGGED>
__device__ __forceinline__ void lb_step4(const float (&ci)[K], float (&co)[K], const float (&x)[K], unsigned &raddr, int l,
                                         bool pass0, float &tf, float (&bot)[4], float &prev_up_d, const LbWatch &wt)
{
    float u0 = __shfl_up_sync(SQK_FULL_MASK, bot[0], 1, L);
    float u1 = __shfl_up_sync(SQK_FULL_MASK, bot[1], 1, L);
    float u2 = __shfl_up_sync(SQK_FULL_MASK, bot[2], 1, L);
    float u3 = __shfl_up_sync(SQK_FULL_MASK, bot[3], 1, L);
    if (l == 0) { u0 = tf; u1 = tf; u2 = tf; u3 = tf; }  // free-start row: a lower bound of j*w for the step's first column
    tf = sqk_add_rd(tf, wt.wstep);
    const float4 y = lb_lds4(raddr);
    raddr = lb_ring_next4<32 * L * 4>(raddr);
    const float up0 = u0, up1 = u1, up2 = u2, up3 = u3;
    float dg = prev_up_d;                                // row above, column j-1
    prev_up_d = u3;
#pragma unroll
    for (int k = 0; k < K; k++) {
        const float lf = ci[k];                          // this row, column j-1
        float a = sqk_lb_cell(x[k], y.x, fminf(fminf(u0, dg), lf));
        if (RAGGED && k == 0 && pass0) a = up0;          // pass-through slot (only when L*K != N)
        float b = sqk_lb_cell(x[k], y.y, fminf(fminf(u1, u0), a));
        if (RAGGED && k == 0 && pass0) b = up1;
        float c = sqk_lb_cell(x[k], y.z, fminf(fminf(u2, u1), b));
        if (RAGGED && k == 0 && pass0) c = up2;
        float d = sqk_lb_cell(x[k], y.w, fminf(fminf(u3, u2), c));
        if (RAGGED && k == 0 && pass0) d = up3;
        dg = lf;
        u0 = a; u1 = b; u2 = c; u3 = d;
        co[k] = d;
    }
    bot[0] = u0; bot[1] = u1; bot[2] = u2; bot[3] = u3;
}

template <int K, int L, bool RAGGED, int COLS>
__global__ void __launch_bounds__(SQK_LB_THREADS, SQK_LB_MINB(K)) sqk_dtw_lb_kernel(const LbArgs a)
{
    constexpr int G = 32 / L;          // reads per warp
    constexpr int RC = (COLS == 4 ? 32 : 16) * L;         // ring capacity (entries), power of two
    constexpr int LAG = COLS * (L - 1);      // columns the last lane runs behind lane 0
    // columns between ring refills: S + CH + LAG <= RC so that no entry is overwritten before its last reader
    constexpr int S = (L == 1) ? 8 : (COLS == 4 ? 12 * L : (COLS == 2 ? 6 * L : 7 * L));
    constexpr int CH = 8 * L;          // raw samples fetched per refill
    static_assert(S % (2 * COLS) == 0 && S + CH + LAG <= RC, "ring schedule");
    static_assert(2 * CH >= S + LAG, "lb_candidate refreshes the cheap threshold for 2 * CH columns ahead: must cover the block");

    // Each group's ring must be aligned to its size (RC*4 bytes) in the shared address space for lb_ring_next; static
    // shared memory starts behind a reserved kilobyte, so the alignment is established here, not by __align__.
    __shared__ float ring_raw[SQK_LB_WARPS * G * RC + RC];
    float *ring_all;
    {
        const unsigned s0 = (unsigned)__cvta_generic_to_shared(ring_raw);
        ring_all = ring_raw + (((s0 + RC * 4u - 1u) & ~(RC * 4u - 1u)) - s0) / 4u;
    }
    __shared__ int32_t ck_all[SQK_LB_WARPS * G * SQK_LB_CKPT];
    __shared__ LbClusters cl_all[SQK_LB_WARPS * G];

    int64_t alloc_lo = a.alloc_lo, alloc_hi = a.alloc_hi;
    resolve_bounds(a.offsets, a.read0, a.n_reads, alloc_lo, alloc_hi);

    const int lane = threadIdx.x & 31;
    const int l = lane % L;
    const int g = lane / L;
    const int gid = (threadIdx.x >> 5) * G + g;
    float *ring = ring_all + gid * RC;
    int32_t *ck = ck_all + gid * SQK_LB_CKPT;
    LbClusters *cl = cl_all + gid;
    for (int q = l; q < RC; q += L) ring[q] = 0.0f;
    __syncwarp();

    const int P = L * K - a.N;
    const bool pass0 = (l >= L - P);
    const int row0 = pass0 ? (L - P) * K + (l - (L - P)) * (K - 1) - 1 : l * K;
    float x[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        const int row = row0 + k;
        x[k] = (row >= 0 && row < a.N) ? (float)a.model[row] : 0.0f;
    }

    // outlier window lo < v < hi as one unsigned compare (samples are int16; lo / hi are clamped to +-40000 by the host)
    const int win_lo = a.lo + 1;
    const bool win_ok = a.hi - 1 >= win_lo;
    const unsigned win_span = (unsigned)(a.hi - 1 - win_lo);
    const float inf = __int_as_float(0x7f800000);
    float c[K], c2[K];
    float bot = inf, bot_b = inf, prev_up = inf, tf = 0.0f;   // COLS == 2: bot is column A's bottom, prev_up is column B's
    float bot4[4] = {inf, inf, inf, inf};                      // COLS == 4
    LbWatch wt; wt.runmin = inf; wt.thr = SQK_LB_THR_INIT; wt.thr_u = -inf; wt.aeps = 0.0f; wt.bslack = 0.0f; wt.w = 0.0f; wt.wstep = 0.0f; wt.n = 0; wt.N = a.N;
    // groups of one warp start in phase and, with equal-length reads, stay in phase: rotate each group's ring by the
    // span its lanes read in one step so that simultaneous reads of different groups fall into different banks
    const int rot = (g * L * COLS) & (RC - 1);
    const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring);
    unsigned raddr = ring_s;
    int n = 0, t = 0, wcount = 0, my_read = -1, n_ref = 0;
    int64_t begin = 0, end = 0, cursor = 0, cursor0 = 0;
    double center = 0.0, scale = 1.0;
    float inv_scale = 1.0f;
    bool done = true, exhausted = false;
#pragma unroll
    for (int k = 0; k < K; k++) c[k] = inf;

    for (;;) {
        // ---- groups that finished pull the next read -------------------------------------------
        const unsigned need = __ballot_sync(SQK_FULL_MASK, done && !exhausted && l == 0);
        if (need) {
            unsigned head = 0;
            if (lane == 0) head = atomicAdd(a.counter, (unsigned)__popc(need));
            head = __shfl_sync(SQK_FULL_MASK, head, 0);
            if (done && !exhausted) {
                const unsigned below = need & ((1u << (g * L)) - 1u);
                const unsigned idx = head + __popc(below);
                if (idx >= (unsigned)a.n_reads) {
                    exhausted = true;
                } else {
                    my_read = (int)idx;
                    const int64_t r = a.read0 + idx;
                    begin = a.offsets[r];
                    end = a.offsets[r + 1];
                    const ReadStats st = a.stats[idx];
                    n = st.n_kept;
                    center = st.center; scale = st.scale;
                    if (st.flags & SQK_FLAG_DEGENERATE) n = -1;
                    if (st.flags & SQK_FLAG_TOO_LONG) n = -2;
                    cursor0 = aligned_block_start(a.base, begin);
                    if (n <= 0) {
                        // same status records as sqk_dtw_kernel: -1 empty, -2 scale undefined, -3 longer than declared
                        if (l == L - 1) {
                            sqk_hit h; h.start = n == 0 ? -1 : (n == -1 ? -2 : -3); h.end = h.start; h.dist = __longlong_as_double(0x7ff8000000000000LL);
                            a.hits[(int64_t)idx * a.hit_stride] = h;
                            LbRead rec; rec.min_l = 0.0f; rec.thr = 0.0f; rec.n_jobs = -1; rec.flags = 0;
                            a.reads[idx] = rec;
                        }
                    } else if (n <= a.short_len || n > SQK_LB_MAX_LEN) {
                        // too short for a window to save anything (or beyond what the proof covers): the exact kernel
                        // takes the whole read as one job
                        if (l == L - 1) {
                            DtwJob jb; jb.cursor = cursor0; jb.read = (int)idx; jb.col0 = 0; jb.n_cols = n; jb.arg_lo = 0;
                            jb.tainted = 0; jb.out = (int)idx * SQK_LB_MAX_CLUSTERS;
                            a.jobs[atomicAdd(a.n_jobs, 1u)] = jb;
                            LbRead rec; rec.min_l = 0.0f; rec.thr = inf; rec.n_jobs = 1; rec.flags = 0;
                            a.reads[idx] = rec;
                        }
                    } else {
                        done = false;
                        t = 0; tf = 0.0f; wcount = 0; n_ref = 0;
                        cursor = cursor0;
                        raddr = ring_s + 4u * (unsigned)((rot - COLS * l) & (RC - 1));   // entry of column t - COLS*l
#pragma unroll
                        for (int k = 0; k < K; k++) c[k] = inf;
                        bot = inf; bot_b = inf;
                        bot4[0] = inf; bot4[1] = inf; bot4[2] = inf; bot4[3] = inf;
                        prev_up = (l == 0) ? 0.0f : inf;
                        wt.runmin = inf; wt.thr = SQK_LB_THR_INIT; wt.thr_u = -inf; wt.n = n;
                        inv_scale = sqk_lb_inv_scale(scale);
                        wt.w = sqk_lb_width(a.xmax_abs, sqk_lb_ymax(a.lo, a.hi, center, scale));
                        wt.wstep = sqk_mul_rd((float)COLS, wt.w);
                        sqk_lb_slack(a.N, wt.w, &wt.aeps, &wt.bslack);
                        if (l == L - 1) lbc_reset(*cl);
                    }
                }
            }
        }
        if (__all_sync(SQK_FULL_MASK, exhausted)) break;

        // ---- refill the rings: raw int16 -> filter -> normalise -> float32 -> shared memory ------
        for (;;) {
            const bool want = !done && (wcount < t + S) && (cursor < end);
            if (!__any_sync(SQK_FULL_MASK, want)) break;
            unsigned keep = 0;
            Samples8 smp;
            const int64_t blk = cursor + l * 8;
            if (want && blk < end) {
                smp = load_block8(a.base, blk, alloc_lo, alloc_hi);
                // samples of the block that belong to the read (blk >= begin - 7), then the outlier window on those
                const int e_lo = begin > blk ? (int)(begin - blk) : 0;
                const int e_hi = end - blk < 8 ? (int)(end - blk) : 8;
                const unsigned inside = ((1u << e_hi) - 1u) & ~((1u << e_lo) - 1u);
#pragma unroll
                for (int e = 0; e < 8; e++) keep |= ((unsigned)(smp.get(e) - win_lo) <= win_span) ? 1u << e : 0u;
                keep &= win_ok ? inside : 0u;
            }
            const int cnt = __popc(keep);
            int incl = cnt;
#pragma unroll
            for (int d = 1; d < L; d <<= 1) {
                const int u = __shfl_up_sync(SQK_FULL_MASK, incl, d, L);
                if (l >= d) incl += u;
            }
            const int group_total = __shfl_sync(SQK_FULL_MASK, incl, L - 1, L);
            if (keep) {
                int pos = wcount + incl - cnt;
                const int p0 = (pos + rot) & (RC - 1);
                if (keep == 0xffu && p0 <= RC - 8) {
                    // the common block: all eight samples kept, no wrap inside it -- eight stores at fixed offsets
                    float *dst = ring + p0;
#pragma unroll
                    for (int e = 0; e < 8; e++) dst[e] = sqk_lb_y32((double)smp.get(e), center, inv_scale);
                } else {
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        if (keep & (1u << e)) {
                            ring[(pos + rot) & (RC - 1)] = sqk_lb_y32((double)smp.get(e), center, inv_scale);
                            pos++;
                        }
                    }
                }
            }
            if (want) {
                if (l == 0) ck[n_ref % SQK_LB_CKPT] = wcount;   // kept samples in front of refill n_ref
                n_ref++;
                wcount += group_total; cursor += CH;
            }
        }
        __syncwarp();

        // the cheap candidate test of this block of steps: U <= thr + (largest (j + N) * w of the block)
        if (l == L - 1 && !done) wt.thr_u = sqk_lb_thr_u(wt.thr, sqk_mul_ru((float)(t + S + a.N), wt.w));
        tf = sqk_lb_virtual((float)t, wt.w);          // free-start row at the block's first column; steps add w
        if constexpr (COLS == 4) {
#pragma unroll 1
            for (int it = 0; it < S; it += 8) {
                lb_step4<K, L, RAGGED>(c, c2, x, raddr, l, pass0, tf, bot4, prev_up, wt);
                const float a1 = bot4[0], b1 = bot4[1], c1 = bot4[2], d1 = bot4[3];
                lb_step4<K, L, RAGGED>(c2, c, x, raddr, l, pass0, tf, bot4, prev_up, wt);
                // one test per eight columns of the last row (rare: a column that may be a candidate); in column order
                if (fminf(fminf(fminf(a1, b1), fminf(c1, d1)), fminf(fminf(bot4[0], bot4[1]), fminf(bot4[2], bot4[3]))) <= wt.thr_u) {
                    LbVals8 vals;
                    vals.v[0] = a1; vals.v[1] = b1; vals.v[2] = c1; vals.v[3] = d1;
                    vals.v[4] = bot4[0]; vals.v[5] = bot4[1]; vals.v[6] = bot4[2]; vals.v[7] = bot4[3];
                    const LbWatchUpd up = lb_candidates(vals, 8, t - 4 * (L - 1), wt.runmin, wt.thr, wt.thr_u, wt.aeps, wt.bslack, wt.w,
                                                        wt.n, wt.N, cl, ck, n_ref, cursor0, 8 * L, a.W, a.W2);
                    wt.runmin = up.runmin; wt.thr = up.thr; wt.thr_u = up.thr_u;
                }
                t += 8;
            }
        } else if constexpr (COLS == 2) {
#pragma unroll 1
            for (int it = 0; it < S; it += 4) {
                lb_step2<K, L, RAGGED>(c, c2, x, raddr, l, pass0, t, tf, bot, bot_b, prev_up, wt, cl, ck, n_ref, cursor0, a.W);
                const float a1 = bot, b1 = bot_b;
                lb_step2<K, L, RAGGED>(c2, c, x, raddr, l, pass0, t + 2, tf, bot, bot_b, prev_up, wt, cl, ck, n_ref, cursor0, a.W);
                // one test per four columns of the last row (rare: a column that may be a candidate); in column order
                if (fminf(fminf(a1, b1), fminf(bot, bot_b)) <= wt.thr_u) {
                    LbVals8 vals;
                    vals.v[0] = a1; vals.v[1] = b1; vals.v[2] = bot; vals.v[3] = bot_b;
                    vals.v[4] = 0.0f; vals.v[5] = 0.0f; vals.v[6] = 0.0f; vals.v[7] = 0.0f;
                    const LbWatchUpd up = lb_candidates(vals, 4, t - 2 * (L - 1), wt.runmin, wt.thr, wt.thr_u, wt.aeps, wt.bslack, wt.w,
                                                        wt.n, wt.N, cl, ck, n_ref, cursor0, 8 * L, a.W, a.W2);
                    wt.runmin = up.runmin; wt.thr = up.thr; wt.thr_u = up.thr_u;
                }
                t += 4;
            }
        } else {
#pragma unroll 1
            for (int it = 0; it < S; it += 2) {
                lb_step<K, L, RAGGED>(c, c2, x, raddr, l, pass0, t, tf, bot, prev_up, wt, cl, ck, n_ref, cursor0, a.W);
                t++;
                lb_step<K, L, RAGGED>(c2, c, x, raddr, l, pass0, t, tf, bot, prev_up, wt, cl, ck, n_ref, cursor0, a.W);
                t++;
            }
        }
        __syncwarp();

        if (!done && t >= n + LAG) {
            if (l == L - 1) {
                lbc_finish(*cl, wt.thr);
                LbRead rec; rec.min_l = wt.runmin; rec.thr = wt.thr; rec.n_jobs = 0; rec.flags = cl->overflow;
                if (cl->n == 0) rec.flags |= 2;           // cannot happen (the minimum itself is a candidate); be safe
                for (int q = 0; q < cl->n; q++)
                    if (cl->tainted[q] < 0) rec.flags |= 4;   // its boundary column had already left the checkpoint ring
                if (rec.flags == 0) {
                    const int nj = cl->n;
                    const unsigned at = atomicAdd(a.n_jobs, (unsigned)nj);
                    for (int q = 0; q < nj; q++) {
                        DtwJob jb;
                        jb.cursor = cl->cursor[q]; jb.read = my_read; jb.col0 = cl->col0[q]; jb.n_cols = cl->hi[q] - cl->col0[q] + 1;
                        jb.arg_lo = cl->lo[q] - cl->col0[q]; jb.tainted = cl->tainted[q]; jb.out = my_read * SQK_LB_MAX_CLUSTERS + q;
                        a.jobs[at + q] = jb;
                        if (a.jobs2) {                         // the same cluster behind a wider window, should the first one taint
                            DtwJob j2 = jb;
                            j2.cursor = cl->cursor2[q]; j2.col0 = cl->col02[q]; j2.n_cols = cl->hi[q] - cl->col02[q] + 1;
                            j2.arg_lo = cl->lo[q] - cl->col02[q]; j2.tainted = cl->tainted2[q];      // -1: no second attempt
                            a.jobs2[my_read * SQK_LB_MAX_CLUSTERS + q] = j2;
                        }
                    }
                    rec.n_jobs = nj;
                }
                a.reads[my_read] = rec;
            }
            done = true;
            wt.thr_u = -inf;
        }
    }
}

// Combine the window results of every read; reads that are not proven get a full-length job in the fallback list.
// Multi-GPU publication of hit records (one process per GPU): besides the local `hits` array, a record is stored into the
// gathered buffer of every peer GPU through P2P-mapped pointers (NVLink stores issued by the kernel that produced the
// record -- a fused compute + all-gather; no collective kernel).  peer[p] points at this rank's block in peer p's buffer,
// already offset to the model column, so record r goes to peer[p][r * hit_stride].
#define SQK_MAX_PEERS 16
struct PeerOut {
    sqk_hit *peer[SQK_MAX_PEERS];
    int n;                      // 0 = single GPU: nothing to publish
};

__device__ __forceinline__ void sqk_publish(const PeerOut &po, int64_t at, const sqk_hit &h)
{
    // 16-byte records: one vector store per peer; consecutive threads write consecutive records (coalesced over NVLink)
    const int4 v = *reinterpret_cast<const int4 *>(&h);
    for (int p = 0; p < po.n; p++) *reinterpret_cast<int4 *>(po.peer[p] + at) = v;
}

struct FinalizeArgs {
    const LbRead *reads;
    const sqk_hit *jobres;     // [n_reads][SQK_LB_MAX_CLUSTERS]
    int n_reads;
    sqk_hit *hits; int hit_stride;
    const int16_t *base; const int64_t *offsets; int64_t read0;
    const ReadStats *stats;
    DtwJob *fb_jobs; unsigned int *n_fb;       // full-length jobs
    // second attempt (stage 1 fills, stage 2 reads `pending`)
    int stage;                                 // 1: after the first windows; 2: after the second-attempt windows
    const DtwJob *jobs2;                       // [n_reads][SQK_LB_MAX_CLUSTERS] or null: no second attempt
    DtwJob *rt_jobs; unsigned int *n_rt;       // second-attempt job list
    unsigned char *pending;                    // [n_reads]: 1 = waiting for its second attempt
    PeerOut po;
};

static __global__ void sqk_dtw_finalize_kernel(const FinalizeArgs a)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n_reads) return;
    if (a.stage == 2) {
        if (!a.pending[r]) return;                     // settled in stage 1
    } else if (a.pending) a.pending[r] = 0;
    const LbRead rec = a.reads[r];
    if (rec.n_jobs < 0) {                              // status hit already written by pass 1: only forward it
        if (a.po.n) sqk_publish(a.po, (int64_t)r * a.hit_stride, a.hits[(int64_t)r * a.hit_stride]);
        return;
    }
    SqkHitLite res[SQK_LB_MAX_CLUSTERS], best;
    const int nj = rec.n_jobs < SQK_LB_MAX_CLUSTERS ? rec.n_jobs : SQK_LB_MAX_CLUSTERS;
    for (int q = 0; q < nj; q++) {
        const sqk_hit h = a.jobres[(int64_t)r * SQK_LB_MAX_CLUSTERS + q];
        res[q].start = h.start; res[q].end = h.end; res[q].dist = h.dist;
    }
    bool only_taint = false;
    if (sqk_lb_decide(rec, res, &best, &only_taint)) {
        sqk_hit h; h.start = best.start; h.end = best.end; h.dist = best.dist;
        a.hits[(int64_t)r * a.hit_stride] = h;
        sqk_publish(a.po, (int64_t)r * a.hit_stride, h);
        return;
    }
    if (a.stage == 1 && only_taint && a.jobs2) {
        // a window was too short: the same clusters behind wider windows, if every one of them can be located
        bool ok = true;
        for (int q = 0; q < nj; q++) ok = ok && a.jobs2[(int64_t)r * SQK_LB_MAX_CLUSTERS + q].tainted >= 0;
        if (ok) {
            const unsigned at = atomicAdd(a.n_rt, (unsigned)nj);
            for (int q = 0; q < nj; q++) a.rt_jobs[at + q] = a.jobs2[(int64_t)r * SQK_LB_MAX_CLUSTERS + q];
            a.pending[r] = 1;
            return;
        }
    }
    DtwJob jb;
    jb.cursor = aligned_block_start(a.base, a.offsets[a.read0 + r]);
    jb.read = r; jb.col0 = 0; jb.n_cols = a.stats[r].n_kept; jb.arg_lo = 0; jb.tainted = 0; jb.out = r;
    a.fb_jobs[atomicAdd(a.n_fb, 1u)] = jb;
}

// Publication of records that other kernels wrote into the local array: the reads of a job list (the full-length
// fallback of the two-pass plan; usually empty) or, with list == nullptr, all reads (single-pass plan).
static __global__ void sqk_publish_kernel(const sqk_hit *hits, int hit_stride, const DtwJob *list, const unsigned int *n_list,
                                          int n_reads, PeerOut po)
{
    const int n = list ? (int)*n_list : n_reads;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int r = list ? list[i].read : i;
        sqk_publish(po, (int64_t)r * hit_stride, hits[(int64_t)r * hit_stride]);
    }
}

// One flag per (step, peer): after the kernels of a step, every rank writes the step number into its slot of every
// peer's flag array; a consumer waits until all slots of its own array have reached the step it wants to read.
struct PeerFlags {
    unsigned long long *arr[SQK_MAX_PEERS];   // arr[p] = flag array of peer p (SQK_MAX_PEERS slots)
    int n, self;
};

static __global__ void sqk_peer_signal_kernel(PeerFlags f, unsigned long long value)
{
    const int p = threadIdx.x;
    if (p < f.n) {
        __threadfence_system();               // (the records were stored by earlier kernels of this stream)
        *reinterpret_cast<volatile unsigned long long *>(f.arr[p] + f.self) = value;
    }
}

static __global__ void sqk_peer_wait_kernel(PeerFlags f, unsigned long long value)
{
    const int q = threadIdx.x;
    if (q < f.n) {
        const volatile unsigned long long *slot = f.arr[f.self] + q;
        while (*slot < value) __nanosleep(200);
        __threadfence_system();
    }
}
