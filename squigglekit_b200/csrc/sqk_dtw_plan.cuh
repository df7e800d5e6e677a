// sqk_dtw_plan.cuh -- the exact two-pass plan for mlpy.dtw_subsequence (MotifSeq.py:437): types and the
// scalar decision logic shared by the kernels (sqk_dtw_lb.cuh, sqk_dtw.cuh) and by the host-side test
// harness (tests/plan_harness.cpp compiles this header with g++ and replays the same decisions on the CPU).
//
// Why two passes.  The float64 recurrence costs ~12 instructions per cell (two DSETP + six selects to carry
// cost and start pointer); a cost-only float32 recurrence costs 4.  The result must still be mlpy's float64
// result bit for bit, so float32 is only ever used as a *proof*:
//
//   pass 1 (sqk_dtw_lb_kernel)  computes for every column j a LOWER BOUND  L[j] <= C[N-1][j]  of the last
//       row of mlpy's cost matrix in float32.  Every local cost satisfies  |x_i - y_j| >= |x32_i - y32_j| - w
//       with  w >= |x_i - x32_i| + |y_j - y32_j|, and a path that starts in column s and ends in (N-1, j) has at
//       most  N + (j - s)  cells (N-1 down moves at most, one cell per column move).  So
//           C[N-1][j]  >=  min over paths ( sum |x32 - y32| + s*w )  -  (j + N) * w .
//       The kernel runs the recurrence for  U = sum |x32 - y32| + s*w : the free-start row feeds  j*w  instead
//       of 0, every addition is rounded towards -inf and |x32 - y32| towards zero, so by induction (min and
//       rounding are monotone) U never exceeds its exact value; then  L[j] = rd(U[N-1][j] - ru((j + N) * w)).
//       That is 3 instructions per cell (FADD.RZ, FMNMX3, FADD.RM).  Columns with  L[j] <= thr = minL + slack
//       are candidates; neighbouring candidates form clusters.
//   pass 2 (sqk_dtw_kernel<JOBS>)  runs the exact float64 recurrence with start pointers on a window
//       [lo - W - 1, hi] around each cluster.  The window's first column is a boundary: rows >= 1 are set to
//       cost -1 with the pointer SQK_TAINT.  -1 is strictly below every true cost (costs are >= 0), so every
//       value computed from the boundary is a strict lower bound of the true value of that path family and
//       carries the taint; every untainted cell holds exactly mlpy's value and pointer (a tainted
//       predecessor only loses a min3 when its true value is larger, too).
//   finalize (sqk_dtw_finalize_kernel)  takes the best exact window result E (first minimum over clusters in
//       column order).  If it is untainted and  E <= thr, every column whose true cost is <= E has
//       L <= C <= E <= thr, i.e. was a candidate, so E is the global first argmin: (start, end, dist) are
//       mlpy's.  Otherwise (tainted minimum, cluster overflow, window start not locatable, E > thr) the read
//       is re-run over its full length by the same float64 kernel.  Correctness never depends on the slack
//       or on W; they only decide how often the fallback runs.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SQK_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#define SQK_HD inline
#endif

#ifndef SQK_LB_COLS
#define SQK_LB_COLS 2             // signal columns per wavefront step of the lower-bound kernel (1: lb_step, 2: lb_step2)
#endif
#define SQK_TAINT (-7)            // start pointer of a cell that derives from a window boundary
#define SQK_LB_MAX_CLUSTERS 4     // candidate clusters kept per read; more -> fallback
#define SQK_LB_GAP 16             // candidate columns at most this far apart share one window
#define SQK_LB_CKPT 64            // refill checkpoints remembered per read (ring)
#define SQK_LB_MAX_LEN (1 << 18)  // longest read the lower-bound proof covers (rounding slop of the f64 sums)

// One unit of work for the exact kernel: columns [col0, col0 + n_cols) of a read, argmin over the columns
// >= col0 + arg_lo.  A full read is col0 = 0, n_cols = n_kept, arg_lo = 0, tainted = 0.
struct DtwJob {
    int64_t cursor;   // absolute sample index of the (16-byte aligned) block the window starts in
    int32_t read;     // read index within the launch
    int32_t col0;     // post-outlier index of the first kept sample at/after `cursor`
    int32_t n_cols;
    int32_t arg_lo;
    int32_t tainted;  // 1: column col0 is a boundary (rows >= 1 forced to -1 / SQK_TAINT)
    int32_t out;      // result slot
};
static_assert(sizeof(DtwJob) == 32, "DtwJob layout");

// Per-read outcome of pass 1.
struct LbRead {
    float min_l;      // min_j L[j]
    float thr;        // candidate threshold at the end of the read (>= min_l)
    int32_t n_jobs;   // windows emitted (slots read*MAX .. +n_jobs); -1: hit already final (empty / degenerate read)
    int32_t flags;    // != 0: go straight to the full-length fallback
};

struct LbClusters {
    int32_t n;
    int32_t overflow;
    int32_t lo[SQK_LB_MAX_CLUSTERS], hi[SQK_LB_MAX_CLUSTERS];
    float mn[SQK_LB_MAX_CLUSTERS];
    // where each cluster's window starts (found when the cluster opens; tainted == -1: not locatable)
    int32_t col0[SQK_LB_MAX_CLUSTERS], tainted[SQK_LB_MAX_CLUSTERS];
    int64_t cursor[SQK_LB_MAX_CLUSTERS];
    // ... and where the wider window of the second attempt starts (W2 columns in front; tainted2 == -1: not locatable)
    int32_t col02[SQK_LB_MAX_CLUSTERS], tainted2[SQK_LB_MAX_CLUSTERS];
    int64_t cursor2[SQK_LB_MAX_CLUSTERS];
};

// ---- float32 helpers with a stated rounding direction (device: one instruction; host: emulated) ----------
#if defined(__CUDA_ARCH__)
SQK_HD float sqk_add_rd(float a, float b) { return __fadd_rd(a, b); }
SQK_HD float sqk_add_ru(float a, float b) { return __fadd_ru(a, b); }
SQK_HD float sqk_add_rz(float a, float b) { return __fadd_rz(a, b); }
SQK_HD float sqk_mul_ru(float a, float b) { return __fmul_ru(a, b); }
SQK_HD float sqk_mul_rd(float a, float b) { return __fmul_rd(a, b); }
SQK_HD float sqk_d2f_ru(double a) { return __double2float_ru(a); }
#else
static inline float sqk_round_dir(double v, int dir)   // dir: -1 down, +1 up, 0 towards zero; v is exact
{
    float f = (float)v;
    if (std::isinf(f) && !std::isinf(v)) {             // overflowed to inf: step back when the direction says so
        const float big = 3.402823466e+38f;
        if (v > 0 && (dir == -1 || dir == 0)) return big;
        if (v < 0 && (dir == 1 || dir == 0)) return -big;
        return f;
    }
    if (std::isnan(f) || std::isinf(f)) return f;
    const double d = (double)f;
    if (dir == 0) dir = v >= 0 ? -1 : 1;
    if (dir < 0 && d > v) f = std::nextafterf(f, -INFINITY);
    if (dir > 0 && d < v) f = std::nextafterf(f, INFINITY);
    return f;
}
SQK_HD float sqk_add_rd(float a, float b) { return sqk_round_dir((double)a + (double)b, -1); }
SQK_HD float sqk_add_ru(float a, float b) { return sqk_round_dir((double)a + (double)b, 1); }
SQK_HD float sqk_add_rz(float a, float b) { return sqk_round_dir((double)a + (double)b, 0); }
SQK_HD float sqk_mul_ru(float a, float b) { return sqk_round_dir((double)a * (double)b, 1); }
SQK_HD float sqk_mul_rd(float a, float b) { return sqk_round_dir((double)a * (double)b, -1); }
SQK_HD float sqk_d2f_ru(double a) { return sqk_round_dir(a, 1); }
#endif

// The scan's float32 image of a normalised sample: (s - center) exactly as numpy computes it (one float64
// rounding), narrowed to float32 and multiplied by the float32 image of 1/scale -- three roundings of 2^-24
// relative each instead of a float64 division per sample.  |y - y32| <= 3.01 * 2^-24 * |y|.
SQK_HD float sqk_lb_y32(double s, double center, float inv_scale32)
{
#if defined(__CUDA_ARCH__)
    return __fmul_rn(__double2float_rn(__dsub_rn(s, center)), inv_scale32);
#else
    return (float)((double)(float)(s - center) * (double)inv_scale32);   // product of two floats is exact in double
#endif
}
SQK_HD float sqk_lb_inv_scale(double scale) { return (float)(1.0 / scale); }

// Width of the local-cost deficit: |x - x32| <= 2^-24 |x| (round-to-nearest conversion) and
// |y - y32| <= 3.01 * 2^-24 |y| (sqk_lb_y32); the 2^-8 relative inflation also absorbs the float64 rounding of
// mlpy's own sums for paths of up to 2^19 cells (SQK_LB_MAX_LEN + motif length).
SQK_HD float sqk_lb_width(double xmax_abs, double ymax_abs)
{
    const double w = (xmax_abs + 3.01 * ymax_abs) * (1.0 / 16777216.0) * (1.0 + 1.0 / 256.0);
    return sqk_d2f_ru(w);
}

// Largest |normalised sample| a read can hold, from its outlier window and normalisation constants.
SQK_HD double sqk_lb_ymax(int lo, int hi, double center, double scale)
{
    const double a = fabs((double)lo - center), b = fabs((double)hi - center);
    const double s = fabs(scale);
    return (a > b ? a : b) / s * (1.0 + 1.0 / 1048576.0);
}

// One cell of the U recurrence: x32, y32 are the round-to-nearest float32 images of x_i, y_j; m = min3 of the
// three predecessors.
SQK_HD float sqk_lb_cell(float x32, float y32, float m)
{
    const float t = sqk_add_rz(x32, -y32);        // |t| <= |x32 - y32|
    return sqk_add_rd(fabsf(t), m);
}

// What the free-start row feeds into row 0 at column j: a lower bound of j*w.  Exact-ish at the first column of
// every block of steps (tj = (float)j), advanced by rounded-down additions of w inside the block.
SQK_HD float sqk_lb_virtual(float tj, float w) { return sqk_mul_rd(tj, w); }
SQK_HD float sqk_lb_virtual_next(float v, float w) { return sqk_add_rd(v, w); }

// L[j] from U[N-1][j]: subtract an upper bound of (j + N) * w.
SQK_HD float sqk_lb_adjust(float u, int j, int N, float w)
{
    return sqk_add_rd(u, -sqk_mul_ru((float)(j + N), w));
}

// Candidate threshold for a running minimum v:  v + (|v| * aeps + b), every step rounded up.
SQK_HD float sqk_lb_thr(float v, float aeps, float b)
{
    return sqk_add_ru(v, sqk_add_ru(sqk_mul_ru(fabsf(v), aeps), b));
}

// slack constants: float32 round-down loses at most one ulp (2^-23 relative) per addition; the bound charges w
// for each of up to N + (j - s) cells where the true deficit is usually far smaller; an alignment of an N-point
// motif has about N..2N cells.
SQK_HD void sqk_lb_slack(int N, float w, float *aeps, float *b)
{
    const float cells = (float)(N + 64);
    *aeps = sqk_mul_ru(cells, 1.1920929e-07f);    // 2^-23
    *b = sqk_mul_ru(sqk_mul_ru(cells, 3.0f), w);
}

// The cheap per-step test runs on U:  L[j] = rd(U - off_j) <= thr  implies  U < thr + ulp(thr) + off_j, so with
// offmax >= off_j for every column of the block,  U <= thr_u  catches every candidate (and a few non-candidates).
SQK_HD float sqk_lb_thr_u(float thr, float offmax)
{
    return sqk_add_ru(sqk_add_ru(thr, sqk_mul_ru(fabsf(thr), 2.4e-7f)), sqk_add_ru(offmax, 1e-30f));
}

#define SQK_LB_THR_INIT 1e30f                     // threshold before the first column (finite: inf <= thr must be false)

// Columns in front of a cluster's first candidate.  Alignments of an N-point motif span <= 1.43 N columns on the
// synthetic benchmark reads and <= 1.40 N on the 60 reads of the reference's example_fast5s.tar (163-point example
// model).  A window that turns out too short taints the minimum; the read then gets a SECOND ATTEMPT with windows of
// W2 = 4 (2N + 32) columns (run on 32 lanes per read: a few hundred columns at the latency of a few hundred short steps),
// and only if that fails, too, the full-length re-run -- which costs the latency of one whole read however few reads need
// it.  With the second attempt a taint is cheap, so W is 1.33 N + 16 (CPU replay on 800 benchmark reads, columns per read
// including second attempts: 197 at 1.5 N + 32, 167 at 1.3 N + 16 with one read in 800 needing the second attempt, 149 at
// 1.1 N + 8 with nine; round 1, without a second attempt: 2 N + 32 = 235, because two full-length re-runs per 100 k reads
// cost more than the longer windows).
SQK_HD int sqk_lb_window(int N) { return N + N / 3 + 16; }
SQK_HD int sqk_lb_window_retry(int N) { return 4 * (2 * N + 32); }

// Where does the window of a cluster starting at column `lo` begin?  ck[(k) % SQK_LB_CKPT] = number of kept
// samples in front of refill k (refills fetch `ch` raw samples each, the first at cursor0); n_ref refills have
// happened.  Asked when the cluster opens, i.e. while its neighbourhood is still in the ring.  Returns false
// when the boundary column is no longer in the ring (-> fallback).
SQK_HD bool sqk_lb_window_start(const int32_t *ck, int n_ref, int64_t cursor0, int ch, int lo, int W,
                                int64_t *cursor, int32_t *col0, int32_t *tainted, int k_from = -1)
{
    const int target = lo - W - 1;                // boundary column
    if (target <= 0) { *cursor = cursor0; *col0 = 0; *tainted = 0; return true; }
    const int oldest = n_ref > SQK_LB_CKPT ? n_ref - SQK_LB_CKPT : 0;
    // any refill whose kept-sample count is <= target will do (an earlier one only makes the window longer); the search
    // walks back from the newest one, or from k_from when the caller knows a later start is pointless
    int k = n_ref - 1;
    if (k_from >= 0 && k_from < k) k = k_from;
    for (; k >= oldest; k--) {
        const int c0 = ck[k % SQK_LB_CKPT];
        if (c0 <= target) {
            if (k == 0) { *cursor = cursor0; *col0 = 0; *tainted = 0; return true; }
            *cursor = cursor0 + (int64_t)k * ch; *col0 = c0; *tainted = 1;
            return true;
        }
    }
    return false;
}

// What the scan knows about the refills of the current read (see sqk_lb_window_start).
struct LbScan {
    const int32_t *ck;
    int n_ref;
    int64_t cursor0;
    int ch;
    int W;
    int W2;     // window of the second attempt (0: none)
};

SQK_HD void lbc_reset(LbClusters &c) { c.n = 0; c.overflow = 0; }

// Column j of the last row has L[j] = v <= thr (thr is the threshold of the running minimum BEFORE this column).
SQK_HD void lbc_event(LbClusters &c, int j, float v, float &runmin, float &thr, float aeps, float b, const LbScan &sc)
{
    if (v < runmin) { runmin = v; thr = sqk_lb_thr(v, aeps, b); }
    if (c.n > 0 && j - c.hi[c.n - 1] <= SQK_LB_GAP) {
        c.hi[c.n - 1] = j;
        if (v < c.mn[c.n - 1]) c.mn[c.n - 1] = v;
        return;
    }
    int k = 0;                                    // drop clusters the running minimum has left behind
    for (int i = 0; i < c.n; i++) {
        if (c.mn[i] <= thr) {
            if (k != i) {
                c.lo[k] = c.lo[i]; c.hi[k] = c.hi[i]; c.mn[k] = c.mn[i];
                c.cursor[k] = c.cursor[i]; c.col0[k] = c.col0[i]; c.tainted[k] = c.tainted[i];
                c.cursor2[k] = c.cursor2[i]; c.col02[k] = c.col02[i]; c.tainted2[k] = c.tainted2[i];
            }
            k++;
        }
    }
    c.n = k;
    if (c.n == SQK_LB_MAX_CLUSTERS) { c.overflow = 1; return; }
    c.lo[c.n] = j; c.hi[c.n] = j; c.mn[c.n] = v;
    if (!sqk_lb_window_start(sc.ck, sc.n_ref, sc.cursor0, sc.ch, j, sc.W, &c.cursor[c.n], &c.col0[c.n], &c.tainted[c.n]))
        c.tainted[c.n] = -1;                      // window start unknown: the read falls back if this cluster survives
    c.tainted2[c.n] = -1; c.col02[c.n] = 0; c.cursor2[c.n] = 0;
    if (sc.W2 > 0) {
        // every refill holds at most ch kept samples: W2 columns back is at least W2 / ch refills back -- start the walk there
        const int back = sc.W2 / sc.ch;
        const int from = sc.n_ref - 1 - back > 0 ? sc.n_ref - 1 - back : 0;
        if (!sqk_lb_window_start(sc.ck, sc.n_ref, sc.cursor0, sc.ch, j, sc.W2, &c.cursor2[c.n], &c.col02[c.n], &c.tainted2[c.n], from))
            c.tainted2[c.n] = -1;                 // no second attempt for this read
    }
    c.n++;
}

// End of read: keep the clusters that still hold a candidate under the final threshold.
SQK_HD void lbc_finish(LbClusters &c, float thr)
{
    int k = 0;
    for (int i = 0; i < c.n; i++) {
        if (c.mn[i] <= thr) {
            if (k != i) {
                c.lo[k] = c.lo[i]; c.hi[k] = c.hi[i]; c.mn[k] = c.mn[i];
                c.cursor[k] = c.cursor[i]; c.col0[k] = c.col0[i]; c.tainted[k] = c.tainted[i];
                c.cursor2[k] = c.cursor2[i]; c.col02[k] = c.col02[i]; c.tainted2[k] = c.tainted2[i];
            }
            k++;
        }
    }
    c.n = k;
}

// Combine the exact window results of one read (sqk_hit layout: start, end, dist).  Returns true when `best`
// is proven to be mlpy's result, false when the read must be re-run over its full length.
struct SqkHitLite { int32_t start, end; double dist; };
SQK_HD bool sqk_lb_decide(const LbRead &r, const SqkHitLite *res, SqkHitLite *best, bool *only_taint = nullptr)
{
    if (only_taint) *only_taint = false;
    if (r.flags != 0 || r.n_jobs <= 0) return false;
    SqkHitLite b = res[0];
    bool taint = (res[0].start == SQK_TAINT);
    for (int i = 1; i < r.n_jobs; i++) {
        if (res[i].start == SQK_TAINT) taint = true;
        if (res[i].dist < b.dist) b = res[i];     // clusters are in column order: strict < keeps the first minimum
    }
    if (taint) {                                  // a window was too short: wider windows can settle it
        if (only_taint) *only_taint = true;
        return false;
    }
    if (!(b.dist <= (double)r.thr)) return false; // also rejects NaN
    *best = b;
    return true;
}
