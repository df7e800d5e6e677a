/*
 * sqk_oracle.c -- CPU ORACLE for the MotifSeq / segmenter hot paths.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (squigglekit_b200/) never imports, links or calls anything in oracle/.
 *
 * Parity status: the DTW arithmetic of the reference lives in the third-party package
 * mlpy 3.5.0 (PyPI "machine-learning-py"; reference README.md:73-97, call site
 * MotifSeq.py:12,437), which is NOT vendored under /root/reference and cannot be
 * installed here (no network).  The DTW below is therefore a restatement of mlpy 3.5.0's
 * published algorithm (mlpy/dtw/cdtw.c: subsequence(), path(), subsequence_path();
 * mlpy/dtw/dtw.pyx: dtw_subsequence()), and the reference ships no golden vectors for it:
 * "PARITY UNPINNED" at the mlpy boundary.  It is cross-checked by (i) a second,
 * structurally different implementation in this file (rolling columns + forward start
 * pointers), (ii) a pure-numpy row-by-row implementation and (iii) brute-force path
 * enumeration on tiny inputs (tests/test_oracle.py).
 *
 * Everything else IS pinned against code that runs in the build container:
 *   - zscale   == sklearn.preprocessing.scale (MotifSeq.py:186-191), bit for bit;
 *   - medmad   == the numpy expression at MotifSeq.py:192-200, bit for bit;
 *   - get_segs == the reference's own segmenter.get_segs/test_segs (segmenter.py:399-494),
 *                 imported unmodified (oracle/refload.py), list for list;
 * see tests/golden/make_golden.py for the fixtures generated from those.
 *
 * Build: make -C oracle   (gcc -O2, -ffp-contract=off so no FMA contraction can change
 * a rounding).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

typedef struct {
    int32_t start;  /* path[1][0]  (MotifSeq.py:438) */
    int32_t end;    /* path[1][-1] (MotifSeq.py:439) */
    double dist;    /* cost[-1, argmin] */
} orc_hit;

/* ------------------------------------------------------------------------------------
 * mlpy 3.5.0 subsequence DTW, full cost matrix (what MotifSeq.py:437 executes).
 *   x = motif (n points), y = normalised signal (m points), cost is row-major n*m.
 *   cdtw.c subsequence(): Manhattan local cost, free start along y (row 0 has no
 *   accumulation), three-way min for the interior.
 * ---------------------------------------------------------------------------------- */
static inline double least3(double up, double diag, double left)
{
    /* same comparison order as mlpy's min3(a=up, b=diag, c=left) */
    double m = up;
    if (diag < m) m = diag;
    if (left < m) m = left;
    return m;
}

static void fill_cost_matrix(const double *x, const double *y, int n, int m, double *cost)
{
    cost[0] = fabs(x[0] - y[0]);
    for (int i = 1; i < n; i++)
        cost[(size_t)i * m] = fabs(x[i] - y[0]) + cost[(size_t)(i - 1) * m];
    for (int j = 1; j < m; j++)
        cost[j] = fabs(x[0] - y[j]);
    for (int i = 1; i < n; i++) {
        const double *above = cost + (size_t)(i - 1) * m;
        double *row = cost + (size_t)i * m;
        const double xi = x[i];
        for (int j = 1; j < m; j++)
            row[j] = fabs(xi - y[j]) + least3(above[j], above[j - 1], row[j - 1]);
    }
}

/* numpy argmin semantics on a float64 row: first minimum; a NaN wins and stops the scan. */
static int first_argmin(const double *v, int m)
{
    double best = v[0];
    int at = 0;
    for (int j = 0; j < m; j++) {
        if (!(v[j] >= best)) {
            best = v[j];
            at = j;
            if (isnan(best)) break;
        }
    }
    return at;
}

/*
 * Back-trace as cdtw.c path()/subsequence_path(): from (n-1, endcol) walk to (0,0); on the
 * border the move is forced (row 0: left, column 0: up); inside, diagonal if it equals the
 * three-way min, else left if it does, else up.  subsequence_path() then drops the leading
 * run of row-0 points except the last one, so path[1][0] is the column where the walk
 * last sits in row 0.  Returns that column; optionally records the (trimmed) path.
 */
static int trace_back(const double *cost, int n, int m, int endcol,
                      int32_t *px, int32_t *py, int64_t *plen)
{
    int i = n - 1, j = endcol;
    int64_t k = 0;
    int start_col = -1;
    /* we record in reverse, stopping at the first time row 0 is reached: everything
       after that in the reversed walk is the row-0 run that subsequence_path() trims. */
    for (;;) {
        if (px) { px[k] = i; py[k] = j; }
        k++;
        if (i == 0) { start_col = j; break; }
        if (j == 0) {
            i--;
        } else {
            const double up = cost[(size_t)(i - 1) * m + j];
            const double dg = cost[(size_t)(i - 1) * m + (j - 1)];
            const double lf = cost[(size_t)i * m + (j - 1)];
            const double mc = least3(up, dg, lf);
            if (dg == mc) { i--; j--; }
            else if (lf == mc) { j--; }
            else { i--; }
        }
    }
    if (px) {
        for (int64_t a = 0, b = k - 1; a < b; a++, b--) {
            int32_t t = px[a]; px[a] = px[b]; px[b] = t;
            t = py[a]; py[a] = py[b]; py[b] = t;
        }
    }
    if (plen) *plen = k;
    return start_col;
}

/*
 * dtw_subsequence(x, y) -> dist, cost, path.   cost (n*m doubles) may be NULL (allocated and
 * freed internally, as mlpy allocates it per call); px/py (capacity n+m) may be NULL.
 * Returns 0, or -1 on bad sizes / allocation failure.
 */
ORC_API int orc_dtw_subsequence(const double *x, int n, const double *y, int m,
                                double *cost, orc_hit *hit,
                                int32_t *px, int32_t *py, int64_t *plen)
{
    if (n < 1 || m < 1) return -1;
    double *c = cost ? cost : (double *)malloc((size_t)n * m * sizeof(double));
    if (!c) return -1;
    fill_cost_matrix(x, y, n, m, c);
    const double *last = c + (size_t)(n - 1) * m;
    const int endcol = first_argmin(last, m);
    hit->dist = last[endcol];
    hit->end = endcol;
    hit->start = trace_back(c, n, m, endcol, px, py, plen);
    if (!cost) free(c);
    return 0;
}

/* ------------------------------------------------------------------------------------
 * Independent second implementation: two rolling columns + forward-propagated start
 * pointers (no matrix, no back-trace).  Used to cross-check the restatement above and for
 * large parity sets.  S[0][j]=j, S[i][0]=0; interior: pointer of the predecessor the
 * back-trace would pick (diag on equality, then left, then up).
 * ---------------------------------------------------------------------------------- */
ORC_API int orc_dtw_subsequence_rolling(const double *x, int n, const double *y, int m,
                                        orc_hit *hit, double *last_row /* m or NULL */)
{
    if (n < 1 || m < 1) return -1;
    double *col = (double *)malloc((size_t)n * sizeof(double));
    int32_t *src = (int32_t *)malloc((size_t)n * sizeof(int32_t));
    if (!col || !src) { free(col); free(src); return -1; }

    double best = 0.0; int best_at = 0, best_src = 0, stop = 0;
    for (int j = 0; j < m; j++) {
        const double yj = y[j];
        double diag_c = 0.0; int32_t diag_s = 0;     /* C[i-1][j-1], S[i-1][j-1] */
        if (j == 0) {
            col[0] = fabs(x[0] - yj); src[0] = 0;
            for (int i = 1; i < n; i++) { col[i] = fabs(x[i] - yj) + col[i - 1]; src[i] = 0; }
        } else {
            diag_c = col[0]; diag_s = src[0];
            col[0] = fabs(x[0] - yj); src[0] = j;
            for (int i = 1; i < n; i++) {
                const double left_c = col[i]; const int32_t left_s = src[i];
                const double up_c = col[i - 1]; const int32_t up_s = src[i - 1];
                const double mc = least3(up_c, diag_c, left_c);
                int32_t s;
                if (diag_c == mc) s = diag_s; else if (left_c == mc) s = left_s; else s = up_s;
                col[i] = fabs(x[i] - yj) + mc; src[i] = s;
                diag_c = left_c; diag_s = left_s;
            }
        }
        const double v = col[n - 1];
        if (last_row) last_row[j] = v;
        if (!stop && (j == 0 || !(v >= best))) {
            best = v; best_at = j; best_src = src[n - 1];
            if (isnan(best)) stop = 1;
        }
    }
    hit->dist = best; hit->end = best_at; hit->start = best_src;
    free(col); free(src);
    return 0;
}

/* ------------------------------------------------------------------------------------
 * numpy's pairwise summation for a contiguous float64 vector (numpy/_core/src/umath
 * loops_utils.h.src, DOUBLE_pairwise_sum): <8 serial; <=128: eight strided accumulators,
 * folded ((0+1)+(2+3))+((4+5)+(6+7)), then the ragged tail serially; otherwise split at
 * n/2 rounded down to a multiple of 8.  np.sum / np.mean / np.std reduce a 1-D contiguous
 * float64 array in ONE such call (pinned in tests against numpy itself).
 * ---------------------------------------------------------------------------------- */
static double np_pairwise(const double *a, int64_t n)
{
    if (n < 8) {
        double r = 0.0;
        for (int64_t i = 0; i < n; i++) r += a[i];
        return r;
    }
    if (n <= 128) {
        double r[8];
        for (int k = 0; k < 8; k++) r[k] = a[k];
        int64_t i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; k++) r[k] += a[i + k];
        double s = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) s += a[i];
        return s;
    }
    int64_t h = n / 2;
    h -= h % 8;
    return np_pairwise(a, h) + np_pairwise(a + h, n - h);
}

ORC_API double orc_np_sum(const double *a, int64_t n) { return np_pairwise(a, n); }

/* scale_outliers (MotifSeq.py:317-324, segmenter.py:311-318): keep lo < s < hi, strict. */
ORC_API int64_t orc_scale_outliers_i16(const int16_t *s, int64_t n, int lo, int hi, double *out)
{
    int64_t k = 0;
    for (int64_t i = 0; i < n; i++)
        if (s[i] > lo && s[i] < hi) out[k++] = (double)s[i];
    return k;
}

/*
 * sklearn.preprocessing.scale on a 1-D float64 vector (MotifSeq.py:186-191):
 *   mean = nanmean(x) = pairwise(x)/n ; std = nanstd(x) = sqrt(pairwise((x-mean)^2)/n) ;
 *   std==0 -> 1 ; x -= mean ; x /= std.   (sklearn's two "mean not close to zero"
 *   corrections cannot trigger for |x| < 2^15: |mean_1| <= ~4e-12 << 1e-8; the Python
 *   wrapper asserts this against the real sklearn output.)   In place; returns 0.
 */
ORC_API int orc_zscale(double *v, int64_t n, double *mean_out, double *std_out)
{
    if (n < 1) return -1;
    const double mean = np_pairwise(v, n) / (double)n;
    double *sq = (double *)malloc((size_t)n * sizeof(double));
    if (!sq) return -1;
    for (int64_t i = 0; i < n; i++) { const double d = v[i] - mean; sq[i] = d * d; }
    double sd = sqrt(np_pairwise(sq, n) / (double)n);
    free(sq);
    if (sd == 0.0) sd = 1.0;
    for (int64_t i = 0; i < n; i++) { v[i] -= mean; v[i] /= sd; }
    if (mean_out) *mean_out = mean;
    if (std_out) *std_out = sd;
    return 0;
}

static int cmp_f64(const void *a, const void *b)
{
    const double p = *(const double *)a, q = *(const double *)b;
    return (p > q) - (p < q);
}

/* np.median on a NaN-free vector: middle element, or the mean of the two middle ones. */
static double np_median(const double *v, int64_t n, double *scratch)
{
    memcpy(scratch, v, (size_t)n * sizeof(double));
    qsort(scratch, (size_t)n, sizeof(double), cmp_f64);
    if (n & 1) return scratch[(n - 1) / 2];
    return (scratch[n / 2 - 1] + scratch[n / 2]) / 2.0;
}

ORC_API double orc_np_median(const double *v, int64_t n)
{
    double *t = (double *)malloc((size_t)(n > 0 ? n : 1) * sizeof(double));
    const double r = np_median(v, n, t);
    free(t);
    return r;
}

/*
 * med-MAD scaling (MotifSeq.py:192-200): med = median(x); mad = median(|x-med|);
 * x = (x - med) / (mad * 1.4826).  mad == 0 gives inf/nan exactly as numpy does.
 */
ORC_API int orc_medmad(double *v, int64_t n, double *med_out, double *mad_out)
{
    if (n < 1) return -1;
    double *t = (double *)malloc((size_t)n * sizeof(double));
    double *d = (double *)malloc((size_t)n * sizeof(double));
    if (!t || !d) { free(t); free(d); return -1; }
    const double med = np_median(v, n, t);
    for (int64_t i = 0; i < n; i++) d[i] = fabs(v[i] - med);
    const double mad = np_median(d, n, t);
    const double scaled = mad * 1.4826;
    for (int64_t i = 0; i < n; i++) v[i] = (v[i] - med) / scaled;
    free(t); free(d);
    if (med_out) *med_out = med;
    if (mad_out) *mad_out = mad;
    return 0;
}

/* ------------------------------------------------------------------------------------
 * segmenter.get_segs (segmenter.py:399-470) restated; sig is the post-outlier signal as
 * float64.  Writes up to max_segs [start,end] pairs, returns the number of segments found
 * (may exceed max_segs: caller sees the overflow), 0 == the reference's `False`.
 * ---------------------------------------------------------------------------------- */
typedef struct {
    int32_t error, corrector, window, seg_dist;
    double std_scale, stall_len;
} orc_seg_cfg;

ORC_API int orc_get_segs(const double *sig, int64_t n, const orc_seg_cfg *cfg,
                         int32_t *segs, int max_segs, double *thr /* top,bot,median,stdev or NULL */)
{
    if (n < 1) return 0;
    double *t = (double *)malloc((size_t)n * sizeof(double));
    if (!t) return -1;
    const double median = np_median(sig, n, t);
    const double mean = np_pairwise(sig, n) / (double)n;          /* np.std: arrmean */
    for (int64_t i = 0; i < n; i++) { const double d = sig[i] - mean; t[i] = d * d; }
    const double stdev = sqrt(np_pairwise(t, n) / (double)n);
    free(t);
    const double top = median + stdev * cfg->std_scale;
    const double bot = median - stdev * cfg->std_scale;
    if (thr) { thr[0] = top; thr[1] = bot; thr[2] = median; thr[3] = stdev; }

    int open = 0;                 /* prev */
    int64_t err = 0, run_err = 0; /* err, prev_err */
    int64_t c = 0, w = cfg->corrector;
    int64_t start = 0;
    int nseg = 0;
    int32_t last_start = 0, last_end = 0;
    const double first_min = (double)cfg->window * cfg->stall_len;
    for (int64_t i = 0; i < n; i++) {
        const double a = sig[i];
        if (a < top && a > bot) {
            if (!open) { start = i; open = 1; }
            c++; w++;
            run_err = 0;
            if (c >= cfg->window && c >= w && (c % w) == 0) err--;
        } else if (open && err < cfg->error) {
            c++; err++; run_err++;
            if (c >= cfg->window && c >= w && (c % w) == 0) err--;
        } else if (open && (c >= cfg->window || (nseg == 0 && (double)c >= first_min))) {
            const int64_t end = i - run_err;
            open = 0;
            if (nseg > 0 && start - last_end < cfg->seg_dist) {
                last_end = (int32_t)end;
            } else {
                nseg++;
                last_start = (int32_t)start; last_end = (int32_t)end;
            }
            if (nseg <= max_segs) { segs[2 * (nseg - 1)] = last_start; segs[2 * (nseg - 1) + 1] = last_end; }
            c = 0; err = 0; run_err = 0;
        } else if (open) {
            open = 0; c = 0; err = 0; run_err = 0;
        }
    }
    return nseg;
}

/* ------------------------------------------------------------------------------------
 * dRNA_segmenter.py, slow5 branch (dRNA_segmenter.py:86-176): the adapter-stall finder.  sig is the
 * post-outlier signal (scale_outliers with the fixed window (0, 1200), :331-334) as float64.
 *   median, stdev over sig[t_start:t_end] (:104-105; an empty slice gives NaN and nothing is ever in range);
 *   top = median + stdev * std_scale (:106);  in range  <=>  a < top  (one-sided, :110);
 *   state machine :108-166 -- unlike get_segs: err is reset when a run opens (:115), tolerated
 *   out-of-range samples only count as errors from position no_err_thresh on (:127-129), w is a constant,
 *   and the scan stops once a closed segment lies more than seg_dist behind (:154-161).
 * Only the first segment is printed (:171-174): returns 1 and writes out[0..1] = start, end, else 0.
 * ---------------------------------------------------------------------------------- */
typedef struct {
    int32_t error, no_err_thresh, corrector, window, seg_dist, t_start, t_end;
    double std_scale;
} orc_adapter_cfg;

ORC_API int orc_adapter_seg(const double *sig, int64_t n, const orc_adapter_cfg *cfg, int32_t *out,
                            double *thr /* top, median, stdev or NULL */)
{
    int64_t s0 = cfg->t_start < n ? cfg->t_start : n, s1 = cfg->t_end < n ? cfg->t_end : n;
    if (s0 < 0) s0 = 0;
    if (s1 < s0) s1 = s0;
    const int64_t ns = s1 - s0;
    double top = NAN, median = NAN, stdev = NAN;
    if (ns > 0) {
        double *t = (double *)malloc((size_t)ns * sizeof(double));
        if (!t) return -1;
        median = np_median(sig + s0, ns, t);
        const double mean = np_pairwise(sig + s0, ns) / (double)ns;
        for (int64_t i = 0; i < ns; i++) { const double d = sig[s0 + i] - mean; t[i] = d * d; }
        stdev = sqrt(np_pairwise(t, ns) / (double)ns);
        free(t);
        top = median + stdev * cfg->std_scale;
    }
    if (thr) { thr[0] = top; thr[1] = median; thr[2] = stdev; }

    int open = 0, have = 0;
    int64_t err = 0, run_err = 0, c = 0, start = 0;
    int64_t first_start = 0, first_end = 0, last_end = 0;
    int nseg = 0;
    const int64_t w = cfg->corrector;
    for (int64_t i = 0; i < n; i++) {
        const double a = sig[i];
        if (a < top) {
            if (!open) { start = i; open = 1; err = 0; }
            c++;
            run_err = 0;
            if (c >= cfg->window && c >= w && w != 0 && (c % w) == 0) err--;
        } else if (open && err < cfg->error) {
            c++;
            if (i >= cfg->no_err_thresh) { err++; run_err++; }
            if (c >= cfg->window && c >= w && w != 0 && (c % w) == 0) err--;
        } else if (open) {
            if (c >= cfg->window) {
                const int64_t end = i - run_err;
                if (nseg > 0 && start - last_end < cfg->seg_dist) {
                    last_end = end;
                    if (nseg == 1) first_end = end;
                } else {
                    nseg++;
                    last_end = end;
                    if (nseg == 1) { first_start = start; first_end = end; have = 1; }
                }
            }
            open = 0; c = 0; err = 0; run_err = 0;
        } else if (nseg > 0 && i - last_end > cfg->seg_dist) {
            break;
        }
    }
    if (have) { out[0] = (int32_t)first_start; out[1] = (int32_t)first_end; }
    return have;
}

/* per read: scale_outliers (0, 1200) -> orc_adapter_seg.  found[r] = 1/0, segs[r] = (start, end). */
ORC_API int orc_adapter_batch(const int16_t *signals, const int64_t *offsets, int64_t n_reads,
                              const orc_adapter_cfg *cfg, int lim_lo, int lim_hi, int n_threads,
                              int32_t *segs, int32_t *found)
{
    int failed = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t r = 0; r < n_reads; r++) {
        const int64_t len = offsets[r + 1] - offsets[r];
        int got = 0;
        segs[2 * r] = 0; segs[2 * r + 1] = 0;
        if (len > 0) {
            double *y = (double *)malloc((size_t)len * sizeof(double));
            if (!y) failed = 1;
            else {
                const int64_t kept = orc_scale_outliers_i16(signals + offsets[r], len, lim_lo, lim_hi, y);
                got = orc_adapter_seg(y, kept, cfg, segs + 2 * r, NULL);
                if (got < 0) { failed = 1; got = 0; }
                free(y);
            }
        }
        found[r] = got;
    }
    return failed ? -1 : 0;
}

/* ------------------------------------------------------------------------------------
 * dRNA_segmenter.py, TSV branch (dRNA_segmenter.py:272-326): the rolling-mean adapter finder.  sig is the
 * post-outlier signal (scale_outliers (0, 1200), :331-334).  As shipped the branch raises NameError (`w` is only
 * defined in a comment, :81 "# w = 2000"); w is a parameter here, default 2000 as that comment says.
 *   t   = pd.Series(sig).rolling(window=w).mean()     (:281-282)  pandas 3.0 roll_mean (aggregations.pyx): the first
 *         w-1 outputs are NaN; sum_x is kept by Kahan-compensated add (compensation_add) / remove (compensation_remove)
 *         per step, output sum_x / nobs; when the last nobs values were identical the output is that value; a negative
 *         result over all-non-negative values is clamped to 0 (and vice versa).
 *   mn  = t.mean()  -> nanops.nanmean: NaN -> 0, numpy pairwise sum over ALL n slots, divided by the count of non-NaN
 *   std = t.std()   -> nanops.nanvar(ddof=1): avg as above; sqr = (avg - t)**2 with NaN slots set to 0; pairwise sum
 *         over all n slots / (count - 1); sqrt.  count <= 1 -> NaN.
 *   bot = mn - std * std_factor                        (:287, std_factor = 0.5)
 *   run detector :291-313 (comparisons with NaN are False; a value equal to bot does nothing; a run with a single
 *   sample below keeps end = 0; no flush at the end), merge rule with seg_dist, then the first segment with
 *   lo_thresh <= b - a <= hi_thresh is printed as (a - shift, b - shift) (:315-323).
 * Returns 1 and writes out[0..1], else 0.
 * ---------------------------------------------------------------------------------- */
typedef struct {
    int32_t w, seg_dist, lo_thresh, hi_thresh, shift;
    double std_factor;
} orc_rollmean_cfg;

/* pandas roll_mean, fixed window w, min_periods = w; out[i] for i < n */
static void pd_roll_mean(const double *v, int64_t n, int64_t w, double *out)
{
    double sum_x = 0.0, comp_add = 0.0, comp_rem = 0.0, prev_value = n > 0 ? v[0] : 0.0;
    int64_t nobs = 0, neg_ct = 0, same = 0;
    for (int64_t i = 0; i < n; i++) {
        if (i >= w) {                                         /* remove_mean(v[i - w]) */
            const double val = v[i - w];
            if (val == val) {
                nobs--;
                const double y = -val - comp_rem;
                const double t = sum_x + y;
                comp_rem = t - sum_x - y;
                sum_x = t;
                if (signbit(val)) neg_ct--;
            }
        }
        {                                                     /* add_mean(v[i]) */
            const double val = v[i];
            if (val == val) {
                nobs++;
                const double y = val - comp_add;
                const double t = sum_x + y;
                comp_add = t - sum_x - y;
                sum_x = t;
                if (signbit(val)) neg_ct++;
                if (val == prev_value) same++; else same = 1;
                prev_value = val;
            }
        }
        double r = NAN;                                       /* calc_mean, minp = w */
        if (nobs >= w && nobs > 0) {
            r = sum_x / (double)nobs;
            if (same >= nobs) r = prev_value;
            else if (neg_ct == 0 && r < 0) r = 0;
            else if (neg_ct == nobs && r > 0) r = 0;
        }
        out[i] = r;
    }
}

ORC_API int orc_rollmean_seg(const double *sig, int64_t n, const orc_rollmean_cfg *cfg, int32_t *out,
                             double *thr /* bot, mn, std or NULL */)
{
    double bot = NAN, mn = NAN, sd = NAN;
    double *t = (double *)malloc((size_t)(n > 0 ? n : 1) * sizeof(double));
    double *z = (double *)malloc((size_t)(n > 0 ? n : 1) * sizeof(double));
    if (!t || !z) { free(t); free(z); return -1; }
    pd_roll_mean(sig, n, cfg->w, t);
    int64_t count = 0;
    for (int64_t i = 0; i < n; i++) { const int ok = t[i] == t[i]; count += ok; z[i] = ok ? t[i] : 0.0; }
    if (count > 0) {
        mn = np_pairwise(z, n) / (double)count;
        if (count > 1) {
            for (int64_t i = 0; i < n; i++) { const double d = mn - z[i]; z[i] = (t[i] == t[i]) ? d * d : 0.0; }
            sd = sqrt(np_pairwise(z, n) / (double)(count - 1));
        }
    }
    bot = mn - sd * cfg->std_factor;
    if (thr) { thr[0] = bot; thr[1] = mn; thr[2] = sd; }

    int begin = 0, have_last = 0, found = 0;
    int64_t start = 0, end = 0, last_a = 0, last_b = 0;
    for (int64_t i = 0; i <= n && !found; i++) {
        int close = 0;
        if (i < n) {
            const double x = t[i];
            if (x < bot && !begin) { start = i; begin = 1; }
            else if (x < bot) end = i;
            else if (x > bot && begin) close = 1;
        }
        /* a list entry is final once the next one is appended (or at the end of the read): test it then */
        const int append = close && !(have_last && start - last_b < cfg->seg_dist);
        if ((append || i == n) && have_last) {
            const int64_t d = last_b - last_a;
            if (!(d > cfg->hi_thresh) && !(d < cfg->lo_thresh)) {
                out[0] = (int32_t)(last_a - cfg->shift); out[1] = (int32_t)(last_b - cfg->shift); found = 1;
            }
        }
        if (close) {
            if (append) { last_a = start; last_b = end; have_last = 1; }
            else last_b = end;
            start = 0; end = 0; begin = 0;
        }
    }
    free(t); free(z);
    return found;
}

/* per read: scale_outliers (lim_lo, lim_hi) -> orc_rollmean_seg.  found[r] = 1/0, segs[r] = (x, y). */
ORC_API int orc_rollmean_batch(const int16_t *signals, const int64_t *offsets, int64_t n_reads,
                               const orc_rollmean_cfg *cfg, int lim_lo, int lim_hi, int n_threads,
                               int32_t *segs, int32_t *found)
{
    int failed = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t r = 0; r < n_reads; r++) {
        const int64_t len = offsets[r + 1] - offsets[r];
        int got = 0;
        segs[2 * r] = 0; segs[2 * r + 1] = 0;
        if (len > 0) {
            double *y = (double *)malloc((size_t)len * sizeof(double));
            if (!y) failed = 1;
            else {
                const int64_t kept = orc_scale_outliers_i16(signals + offsets[r], len, lim_lo, lim_hi, y);
                got = orc_rollmean_seg(y, kept, cfg, segs + 2 * r, NULL);
                if (got < 0) { failed = 1; got = 0; }
                free(y);
            }
        }
        found[r] = got;
    }
    return failed ? -1 : 0;
}

/* ------------------------------------------------------------------------------------
 * Batch drivers (OpenMP over reads) -- the "reference arm" / cpu_baseline of bench.py and
 * the large-set parity checker.  Per read they do exactly what the reference's main loop
 * does: scale_outliers -> normalise -> dtw_subsequence -> (start,end,dist).
 *   scale_mode: 0 zscale, 1 medmad, 2 none.   full_matrix!=0 uses the mlpy-shaped
 *   n*m matrix + back-trace (the faithful CPU cost), 0 the rolling cross-check.
 * ---------------------------------------------------------------------------------- */
ORC_API int orc_motifseq_batch(const int16_t *signals, const int64_t *offsets, int64_t n_reads,
                               const double *model, int n_model, int lo, int hi, int scale_mode,
                               int full_matrix, int n_threads,
                               orc_hit *hits, int32_t *n_kept)
{
    int failed = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t r = 0; r < n_reads; r++) {
        const int64_t len = offsets[r + 1] - offsets[r];
        orc_hit h; h.start = -1; h.end = -1; h.dist = NAN;
        int64_t kept = 0;
        if (len > 0) {
            double *y = (double *)malloc((size_t)len * sizeof(double));
            if (!y) { failed = 1; }
            else {
                kept = orc_scale_outliers_i16(signals + offsets[r], len, lo, hi, y);
                if (kept > 0) {
                    if (scale_mode == 0) orc_zscale(y, kept, NULL, NULL);
                    else if (scale_mode == 1) orc_medmad(y, kept, NULL, NULL);
                    int rc = full_matrix
                        ? orc_dtw_subsequence(model, n_model, y, (int)kept, NULL, &h, NULL, NULL, NULL)
                        : orc_dtw_subsequence_rolling(model, n_model, y, (int)kept, &h, NULL);
                    if (rc) failed = 1;
                }
                free(y);
            }
        }
        hits[r] = h;
        if (n_kept) n_kept[r] = (int32_t)kept;
    }
    return failed ? -1 : 0;
}

/* segmenter main-loop body per read: sig[:Num] (Num=0 -> drop last sample,
 * segmenter.py:104-105,207) -> scale_outliers -> get_segs.  n_segs[r] = count (0 = False). */
ORC_API int orc_segmenter_batch(const int16_t *signals, const int64_t *offsets, int64_t n_reads,
                                const orc_seg_cfg *cfg, int lim_lo, int lim_hi, int num,
                                int max_segs, int n_threads, int32_t *segs, int32_t *n_segs)
{
    int failed = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t r = 0; r < n_reads; r++) {
        int64_t len = offsets[r + 1] - offsets[r];
        /* python slice semantics of sig[:Num] with Num = num ? num : -1 */
        int64_t use;
        if (num == 0) use = len - 1;
        else if (num > 0) use = num < len ? num : len;
        else use = len + num;
        if (use < 0) use = 0;
        int cnt = 0;
        if (use > 0) {
            double *y = (double *)malloc((size_t)use * sizeof(double));
            if (!y) failed = 1;
            else {
                const int64_t kept = orc_scale_outliers_i16(signals + offsets[r], use, lim_lo, lim_hi, y);
                if (kept > 0) cnt = orc_get_segs(y, kept, cfg, segs + (size_t)r * max_segs * 2, max_segs, NULL);
                free(y);
            }
        }
        n_segs[r] = cnt;
    }
    return failed ? -1 : 0;
}

/* convert_to_pA_numpy + np.round(.., 2) (segmenter.py:515-517, 345-349): (d + offset) * raw_unit, then
 * numpy's round = rint(x * 100) / 100. */
static inline double pa_value(int d, double offset, double raw_unit)
{
    const double x = ((double)d + offset) * raw_unit;
    return rint(x * 100.0) / 100.0;
}

ORC_API void orc_convert_to_pa(const int16_t *s, int64_t n, double offset, double raw_unit, double *out)
{
    for (int64_t i = 0; i < n; i++) out[i] = pa_value(s[i], offset, raw_unit);
}

/* fast5 default path of segmenter.py: pA conversion -> sig[:Num] -> scale_outliers on pA -> get_segs. */
ORC_API int orc_segmenter_batch_pa(const int16_t *signals, const int64_t *offsets, int64_t n_reads,
                                   const double *pa_offset, const double *pa_scale,
                                   const orc_seg_cfg *cfg, int lim_lo, int lim_hi, int num,
                                   int max_segs, int n_threads, int32_t *segs, int32_t *n_segs)
{
    int failed = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t r = 0; r < n_reads; r++) {
        int64_t len = offsets[r + 1] - offsets[r];
        int64_t use;
        if (num == 0) use = len - 1;
        else if (num > 0) use = num < len ? num : len;
        else use = len + num;
        if (use < 0) use = 0;
        int cnt = 0;
        if (use > 0) {
            double *y = (double *)malloc((size_t)use * sizeof(double));
            if (!y) failed = 1;
            else {
                int64_t kept = 0;
                for (int64_t i = 0; i < use; i++) {
                    const double v = pa_value(signals[offsets[r] + i], pa_offset[r], pa_scale[r]);
                    if (v > lim_lo && v < lim_hi) y[kept++] = v;
                }
                if (kept > 0) cnt = orc_get_segs(y, kept, cfg, segs + (size_t)r * max_segs * 2, max_segs, NULL);
                free(y);
            }
        }
        n_segs[r] = cnt;
    }
    return failed ? -1 : 0;
}

/* float64-signal variants of the two batch drivers: what the reference's `-s` path does with a TSV of floats
 * (MotifSeq.py:270-298; segmenter.py:198-211). */
ORC_API int orc_motifseq_batch_f64(const double *signals, const int64_t *offsets, int64_t n_reads,
                                   const double *model, int n_model, int lo, int hi, int scale_mode,
                                   int full_matrix, int n_threads, orc_hit *hits, int32_t *n_kept)
{
    int failed = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t r = 0; r < n_reads; r++) {
        const int64_t len = offsets[r + 1] - offsets[r];
        orc_hit h; h.start = -1; h.end = -1; h.dist = NAN;
        int64_t kept = 0;
        if (len > 0) {
            double *y = (double *)malloc((size_t)len * sizeof(double));
            if (!y) failed = 1;
            else {
                for (int64_t i = 0; i < len; i++) {
                    const double v = signals[offsets[r] + i];
                    if (v > lo && v < hi) y[kept++] = v;
                }
                if (kept > 0) {
                    if (scale_mode == 0) orc_zscale(y, kept, NULL, NULL);
                    else if (scale_mode == 1) orc_medmad(y, kept, NULL, NULL);
                    int rc = full_matrix
                        ? orc_dtw_subsequence(model, n_model, y, (int)kept, NULL, &h, NULL, NULL, NULL)
                        : orc_dtw_subsequence_rolling(model, n_model, y, (int)kept, &h, NULL);
                    if (rc) failed = 1;
                }
                free(y);
            }
        }
        hits[r] = h;
        if (n_kept) n_kept[r] = (int32_t)kept;
    }
    return failed ? -1 : 0;
}

ORC_API int orc_segmenter_batch_f64(const double *signals, const int64_t *offsets, int64_t n_reads,
                                    const orc_seg_cfg *cfg, int lim_lo, int lim_hi, int num,
                                    int max_segs, int n_threads, int32_t *segs, int32_t *n_segs)
{
    int failed = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t r = 0; r < n_reads; r++) {
        int64_t len = offsets[r + 1] - offsets[r];
        int64_t use;
        if (num == 0) use = len - 1;
        else if (num > 0) use = num < len ? num : len;
        else use = len + num;
        if (use < 0) use = 0;
        int cnt = 0;
        if (use > 0) {
            double *y = (double *)malloc((size_t)use * sizeof(double));
            if (!y) failed = 1;
            else {
                int64_t kept = 0;
                for (int64_t i = 0; i < use; i++) {
                    const double v = signals[offsets[r] + i];
                    if (v > lim_lo && v < lim_hi) y[kept++] = v;
                }
                if (kept > 0) cnt = orc_get_segs(y, kept, cfg, segs + (size_t)r * max_segs * 2, max_segs, NULL);
                free(y);
            }
        }
        n_segs[r] = cnt;
    }
    return failed ? -1 : 0;
}

ORC_API int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
