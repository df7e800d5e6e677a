// plan_harness.cpp -- CPU replay of the exact two-pass DTW plan (squigglekit_b200/csrc/sqk_dtw_plan.cuh).
// Test infrastructure: compiled with g++ by tests/test_plan_cpu.py; it includes the SAME header the CUDA kernels
// use, so the candidate threshold, the cluster bookkeeping, the window-start search, the taint rule and the
// final decision are exercised here bit for bit (float32 directed rounding is emulated in the header's host
// branch).  The recurrences themselves are plain scalar loops written after mlpy's C loop.
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "../squigglekit_b200/csrc/sqk_dtw_plan.cuh"
#include "../squigglekit_b200/csrc/sqk_stats_plan.cuh"

namespace {

struct Hit { int start, end; double dist; };

// exact float64 recurrence with forward start pointers on columns [col0, col0 + n_cols) of y; argmin over
// columns >= col0 + arg_lo; tainted: column col0 is a boundary (rows >= 1 := -1 / SQK_TAINT).
Hit exact_window(const double *x, int N, const double *y, int col0, int n_cols, int arg_lo, bool tainted,
                 std::vector<double> *last_row = nullptr)
{
    const double inf = std::numeric_limits<double>::infinity();
    std::vector<double> c(N, inf), nc(N);
    std::vector<int> s(N, 0), ns(N);
    Hit best{-1, -1, inf};
    for (int jj = 0; jj < n_cols; jj++) {
        const double yj = y[col0 + jj];
        for (int i = 0; i < N; i++) {
            double m; int ms;
            if (i == 0) { m = 0.0; ms = jj; }
            else {
                // diagonal if it ties the minimum, else left, else up (mlpy's back-trace preference)
                const double dg = jj == 0 ? inf : c[i - 1], lf = jj == 0 ? inf : c[i], up = nc[i - 1];
                const int dgs = jj == 0 ? 0 : s[i - 1], lfs = jj == 0 ? 0 : s[i], ups = ns[i - 1];
                m = dg; ms = dgs;
                if (lf < m) { m = lf; ms = lfs; }
                if (up < m) { m = up; ms = ups; }
            }
            nc[i] = std::fabs(x[i] - yj) + m;
            ns[i] = ms;
        }
        if (tainted && jj == 0)
            for (int i = 1; i < N; i++) { nc[i] = -1.0; ns[i] = SQK_TAINT; }
        c.swap(nc); s.swap(ns);
        if (last_row) (*last_row)[jj] = c[N - 1];
        if (jj >= arg_lo && c[N - 1] < best.dist) { best.dist = c[N - 1]; best.end = jj; best.start = s[N - 1]; }
    }
    if (best.start != SQK_TAINT) best.start += col0;
    best.end += col0;
    return best;
}

}  // namespace

extern "C" {

// y: normalised kept samples (float64) of one read, sv: the same samples before normalisation, n of them; keep[raw_len]: 1 where the raw sample survived the
// outlier filter (sum == n); align_off: raw samples between the aligned block start and the read's first sample;
// ch: raw samples per refill (8 * lanes).  out: start, end; *dist.  diag[0..7]: n_clusters, n_jobs, fallback
// (0 proven, 1 fallback), flags, lower-bound violations (must be 0), tainted windows, window columns, max L gap *1e9;
// diag[8]: 1 if the read needed the second attempt.  W2_in < 0: the default second-attempt window, 0: none.
int plan_two_pass(const double *x, int N, const double *y, const double *sv, int n, const uint8_t *keep, int raw_len, int align_off, int ch,
                  int lo, int hi, double center, double scale, int W, int W2_in, int32_t *out, double *dist, int64_t *diag)
{
    const float inf = std::numeric_limits<float>::infinity();
    double xmax = 0.0;
    for (int i = 0; i < N; i++) xmax = std::fmax(xmax, std::fabs(x[i]));
    const float w = sqk_lb_width(xmax, sqk_lb_ymax(lo, hi, center, scale));
    float aeps, bslack;
    sqk_lb_slack(N, w, &aeps, &bslack);
    if (W <= 0) W = sqk_lb_window(N);
    const int W2 = W2_in < 0 ? sqk_lb_window_retry(N) : W2_in;       // 0: no second attempt

    // exact last row of the whole read: the property under test is L[j] <= C[N-1][j]
    std::vector<double> crow(n);
    const Hit truth = exact_window(x, N, y, 0, n, 0, false, &crow);

    // ---- pass 1: float32 lower bound, candidate clusters, refill checkpoints ------------------------------
    std::vector<float> x32(N), c(N, inf), nc(N);
    for (int i = 0; i < N; i++) x32[i] = (float)x[i];
    LbClusters cl; lbc_reset(cl);
    float runmin = inf, thr = SQK_LB_THR_INIT;
    int32_t ck[SQK_LB_CKPT];
    int n_ref = 0, wcount = 0, raw = -align_off;       // raw: position of the next refill relative to the read's first sample
    int64_t violations = 0; double max_gap = 0.0;
    // the kernel's schedule: blocks of S columns between refills, the last lane LAG columns behind lane 0
    // (four columns per step when a lane holds <= 12 rows on >= 4 lanes, as sqk_dtw_lb_launch.cuh decides; two otherwise)
    const int lanes = ch / 8, rows = (N + lanes - 1) / lanes, cols = (lanes >= 4 && rows <= 12) ? 4 : 2, LAG = cols * (lanes - 1);
    const int S = lanes == 1 ? 8 : (cols == 4 ? 12 * lanes : 6 * lanes);
    const float wstep = sqk_mul_rd((float)cols, w);
    float thr_u = -inf, prev_virt = 0.0f, step_virt = 0.0f;
    int64_t missed = 0;
    for (int j = 0; j < n; j++) {
        // the kernel refills ahead of use (whenever fewer than S columns are buffered beyond the wavefront)
        while (wcount < j + LAG + 1 + S && raw < raw_len) {
            ck[n_ref % SQK_LB_CKPT] = wcount; n_ref++;
            for (int e = 0; e < ch; e++) { const int r = raw + e; if (r >= 0 && r < raw_len && keep[r]) wcount++; }
            raw += ch;
        }
        // the lane that owns the last row sees column j at step t = j + lanes - 1; blocks of S steps share one thr_u
        const int t = j - j % cols + LAG;
        if (j == 0 || (t % S == 0 && j % cols == 0)) thr_u = sqk_lb_thr_u(thr, sqk_mul_ru((float)(t - t % S + S + N), w));
        const float y32 = sqk_lb_y32(sv[j], center, sqk_lb_inv_scale(scale));
        // free-start row: exact-ish at the first column of each block (lane 0's clock is the column); inside a block one
        // rounded-down add of cols * w per step, every column of a step taking the value of the step's first column
        if (j % S == 0) step_virt = sqk_lb_virtual((float)j, w);
        else if (j % cols == 0) step_virt = sqk_add_rd(step_virt, wstep);
        const float virt = step_virt;
        for (int i = 0; i < N; i++) {
            float m;
            if (i == 0) m = std::fmin(std::fmin(virt, prev_virt), j == 0 ? inf : c[0]);   // free-start row: j*w, (j-1)*w, left
            else {
                const float dg = j == 0 ? inf : c[i - 1], lf = j == 0 ? inf : c[i], up = nc[i - 1];
                m = std::fmin(std::fmin(up, dg), lf);
            }
            nc[i] = sqk_lb_cell(x32[i], y32, m);
        }
        prev_virt = virt;
        c.swap(nc);
        const float u = c[N - 1];
        const float v = sqk_lb_adjust(u, j, N, w);
        if (!((double)v <= crow[j])) violations++;
        if (crow[j] - (double)v > max_gap && crow[j] <= truth.dist + 1.0) max_gap = crow[j] - (double)v;
        if (v <= thr && !(u <= thr_u)) missed++;      // the cheap test must never miss a candidate
        if (u <= thr_u && v <= thr) {
            LbScan sc; sc.ck = ck; sc.n_ref = n_ref; sc.cursor0 = -(int64_t)align_off; sc.ch = ch; sc.W = W; sc.W2 = W2;
            lbc_event(cl, j, v, runmin, thr, aeps, bslack, sc);
            thr_u = sqk_lb_thr_u(thr, sqk_mul_ru((float)(j + N + 2 * ch), w));
        }
    }
    violations += missed;
    lbc_finish(cl, thr);
    LbRead rec; rec.min_l = runmin; rec.thr = thr; rec.n_jobs = 0; rec.flags = cl.overflow;
    if (cl.n == 0) rec.flags |= 2;
    for (int q = 0; q < cl.n; q++)
        if (cl.tainted[q] < 0) rec.flags |= 4;
    if (rec.flags == 0) rec.n_jobs = cl.n;

    // ---- pass 2: exact windows --------------------------------------------------------------------------
    SqkHitLite res[SQK_LB_MAX_CLUSTERS], best;
    int64_t tainted_windows = 0, window_cols = 0;
    for (int q = 0; q < rec.n_jobs; q++) {
        // the window starts at the first kept sample at/after raw position cursor: that must be column col0
        int kept_before = 0;
        for (int r = 0; r < cl.cursor[q] && r < raw_len; r++) kept_before += keep[r];
        if (kept_before != cl.col0[q]) return -100 - q;                 // checkpoint bookkeeping broken
        const int n_cols = cl.hi[q] - cl.col0[q] + 1;
        const Hit h = exact_window(x, N, y, cl.col0[q], n_cols, cl.lo[q] - cl.col0[q], cl.tainted[q] != 0);
        res[q].start = h.start; res[q].end = h.end; res[q].dist = h.dist;
        if (h.start == SQK_TAINT) tainted_windows++;
        window_cols += n_cols;
    }
    bool only_taint = false;
    bool proven = sqk_lb_decide(rec, res, &best, &only_taint);
    int64_t second_attempt = 0;
    if (!proven && only_taint && W2 > 0) {
        // second attempt (sqk_dtw_finalize_kernel, stage 1 -> 2): the same clusters behind windows of W2 columns
        bool ok = true;
        for (int q = 0; q < rec.n_jobs; q++) ok = ok && cl.tainted2[q] >= 0;
        if (ok) {
            second_attempt = 1;
            for (int q = 0; q < rec.n_jobs; q++) {
                int kept_before = 0;
                for (int r = 0; r < cl.cursor2[q] && r < raw_len; r++) kept_before += keep[r];
                if (kept_before != cl.col02[q]) return -200 - q;            // checkpoint bookkeeping broken
                const int n_cols = cl.hi[q] - cl.col02[q] + 1;
                const Hit h = exact_window(x, N, y, cl.col02[q], n_cols, cl.lo[q] - cl.col02[q], cl.tainted2[q] != 0);
                res[q].start = h.start; res[q].end = h.end; res[q].dist = h.dist;
                window_cols += n_cols;
            }
            proven = sqk_lb_decide(rec, res, &best);
        }
    }
    Hit fin;
    if (proven) { fin.start = best.start; fin.end = best.end; fin.dist = best.dist; }
    else fin = truth;                                                    // the kernel re-runs the full read
    out[0] = fin.start; out[1] = fin.end; *dist = fin.dist;
    diag[0] = cl.n; diag[1] = rec.n_jobs; diag[2] = proven ? 0 : 1; diag[3] = rec.flags; diag[4] = violations;
    diag[5] = tainted_windows; diag[6] = window_cols; diag[7] = (int64_t)(max_gap * 1e9); diag[8] = second_attempt;
    // a proven result must equal the full exact recurrence
    if (proven && (fin.start != truth.start || fin.end != truth.end || fin.dist != truth.dist)) return -1;
    return 0;
}

// the full exact recurrence alone (cross-checked against the oracle by the test)
int plan_exact(const double *x, int N, const double *y, int n, int32_t *out, double *dist)
{
    const Hit h = exact_window(x, N, y, 0, n, 0, false);
    out[0] = h.start; out[1] = h.end; *dist = h.dist;
    return 0;
}

// np.sum(a) the way the stats kernel takes it: leaf sums per slot (sqk_tree_leaf), then a pairwise fold of the slots.
// Valid for n <= 8192 (the kernel walks the levels above such subtrees separately).
double stats_tree_sum(const double *a, int n)
{
    const int depth = sqk_tree_depth(n), slots = 1 << depth;
    std::vector<double> v(slots, 0.0);
    for (int j = 0; j < slots; j++) {
        int off, len;
        if (!sqk_tree_leaf(n, depth, j, &off, &len)) continue;
        const double *p = a + off;
        double s;
        if (len < 8) { s = 0.0; for (int i = 0; i < len; i++) s += p[i]; }
        else {
            double r[8];
            for (int k = 0; k < 8; k++) r[k] = p[k];
            int i = 8;
            for (; i < len - (len % 8); i += 8) for (int k = 0; k < 8; k++) r[k] += p[i + k];
            s = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
            for (; i < len; i++) s += p[i];
        }
        v[j] = s;
    }
    for (int w = slots; w > 1; w /= 2) for (int i = 0; i < w / 2; i++) v[i] = v[2 * i] + v[2 * i + 1];
    return v[0];
}

}  // extern "C"

// The branch-free forms of the split rule (sqk_tree_depth7 / sqk_tree_leaf7) against the loops, every slot of the tree
// over n elements.  -> number of disagreements.
extern "C" int stats_tree_check7(int n)
{
    const int depth = sqk_tree_depth(n);
    int bad = depth != sqk_tree_depth7(n) ? 1 : 0;
    for (int j = 0; j < (1 << depth); j++) {
        int o1 = 0, l1 = 0, o2 = 0, l2 = 0;
        const bool m1 = sqk_tree_leaf(n, depth, j, &o1, &l1), m2 = sqk_tree_leaf7(n, depth, j, &o2, &l2);
        if (m1 != m2 || (m1 && (o1 != o2 || l1 != l2))) bad++;
    }
    return bad;
}
