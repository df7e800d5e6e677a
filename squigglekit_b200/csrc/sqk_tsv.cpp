// sqk_tsv.cpp -- host-side text I/O of the SquigglePull signal format, the front door of both command lines:
//
//     fast5 <TAB> readID [<TAB> digitisation <TAB> offset <TAB> range <TAB> sampling_rate] <TAB> s0 <TAB> s1 ...
//
// written by SquigglePull.py:243-253 (print_data) and consumed by MotifSeq.py:252-298 (signal from column 8) and
// segmenter.py:179-230 (signal from column 4), where every field goes through float() / int() in a Python list
// comprehension.  Here a buffer of text is cut into lines, the signal fields of every line are counted and then parsed --
// both in parallel over the lines (OpenMP) -- straight into the int16 batch the GPU path consumes.  Lines that are not
// plain int16 integers (pA output, exponents, empty fields) are flagged, not guessed at: the caller sends them through the
// float path.  No GPU work in this file; it is part of libsqk.so so that one ctypes handle serves the whole drop-in.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <charconv>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif
#include <sched.h>

#include "../../include/sqk.h"

// host threads this process may run on (NOT OMP_NUM_THREADS: launchers such as torch.distributed.run set that to 1)
static int host_threads(int asked)
{
    if (asked > 0) return asked;
    cpu_set_t set;
    CPU_ZERO(&set);
    int n = 0;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) n = CPU_COUNT(&set);
    if (n < 1) n = 1;
    return n > 64 ? 64 : n;
}

// memcpy on all host threads (1 MiB pieces): what sqk_api.cu uses to bring a PAGEABLE caller buffer into its pinned
// staging buffer -- the driver's own staging of pageable memory runs on one thread
extern "C" void sqk_parallel_memcpy(void *dst, const void *src, size_t bytes, int n_threads)
{
    const size_t piece = (size_t)1 << 20;
    const int64_t n = (int64_t)((bytes + piece - 1) / piece);
    int nt = host_threads(n_threads);
    if (nt > 16) nt = 16;                                       // memory bandwidth, not cores, limits this
    if (n <= 1 || nt <= 1) { memcpy(dst, src, bytes); return; }
#pragma omp parallel for schedule(static) num_threads(nt)
    for (int64_t i = 0; i < n; i++) {
        const size_t at = (size_t)i * piece;
        memcpy((char *)dst + at, (const char *)src + at, bytes - at < piece ? bytes - at : piece);
    }
}

// parse one line's signal part [p, e) into out[0..n_fields); returns the status bits.  `sentinel`: the byte at e can be read
// and is not a digit (a newline: every line but an unterminated last one) -- the digit loop of the common case (a plain
// non-negative number followed by a tab or the end of the line) then needs no bounds test.
static int parse_fields_i16(const char *p, const char *e, int16_t *out, int64_t n_fields, bool sentinel)
{
    int status = 0;
    unsigned any_nonzero = 0;
    int64_t k = 0;
    while (p <= e && k < n_fields) {
        if (sentinel && p < e) {
            const unsigned char *q = (const unsigned char *)p;
            unsigned v = (unsigned)*q - '0';
            if (v <= 9u) {
                unsigned d;
                q++;
                while ((d = (unsigned)*q - '0') <= 9u) { v = v * 10u + d; q++; }   // (stops at e at the latest)
                const int nd = (int)((const char *)q - p);
                if (nd <= 5 && v <= 32767u && (*q == '\t' || (const char *)q == e)) {
                    out[k++] = (int16_t)v;
                    any_nonzero |= v;
                    p = (const char *)q + 1;
                    continue;
                }
            }
        }
        // everything else: signs, CRLF, junk, out-of-range values, the unterminated last line
        const char *q = p;
        bool neg = false;
        if (q < e && (*q == '-' || *q == '+')) { neg = *q == '-'; q++; }
        int v = 0, nd = 0;
        while (q < e && (unsigned)(*q - '0') <= 9u) {
            if (nd < 7) v = v * 10 + (*q - '0');
            nd++; q++;
        }
        if (q < e && *q == '\r' && q + 1 == e) q++;             // CRLF files
        const bool ends = q == e || *q == '\t';
        if (nd == 0 || nd > 6 || !ends) {
            status |= SQK_TSV_NOT_INT16;
            // skip to the end of this field
            while (q < e && *q != '\t') q++;
            out[k] = 0;
        } else {
            if (neg) v = -v;
            if (v < -32768 || v > 32767) { status |= SQK_TSV_NOT_INT16; out[k] = 0; }
            else out[k] = (int16_t)v;
            any_nonzero |= (unsigned)(v != 0);
        }
        k++;
        p = q + 1;
    }
    if (!(status & SQK_TSV_NOT_INT16) && !any_nonzero) status |= SQK_TSV_ALL_ZERO;
    return status;
}

// positions of the newlines in text[lo, hi), in order, found by all threads (each a contiguous piece)
static void find_newlines(const char *text, int64_t lo, int64_t hi, int nt, std::vector<int64_t> &out)
{
    if (hi <= lo) return;
    if (nt > 1 && hi - lo < (1 << 20)) nt = 1;
    std::vector<std::vector<int64_t>> found((size_t)nt);
#pragma omp parallel for schedule(static, 1) num_threads(nt)
    for (int t = 0; t < nt; t++) {
        const int64_t a = lo + (hi - lo) * t / nt, b = lo + (hi - lo) * (t + 1) / nt;
        std::vector<int64_t> &v = found[(size_t)t];
        const char *p = text + a, *e = text + b;
        while (p < e) {
            const char *nl = (const char *)memchr(p, '\n', (size_t)(e - p));
            if (!nl) break;
            v.push_back(nl - text);
            p = nl + 1;
        }
    }
    for (int t = 0; t < nt; t++) out.insert(out.end(), found[(size_t)t].begin(), found[(size_t)t].end());
}

extern "C" {

int sqk_tsv_parse(const char *text, int64_t n_bytes, int is_final, int start_col, int64_t max_lines, int64_t max_samples,
                  int n_threads, int16_t *samples, int64_t *offsets, int64_t *line_begin, int64_t *sig_begin,
                  int32_t *status, int64_t *n_lines_out, int64_t *consumed_out)
{
    if (!text || !samples || !offsets || !line_begin || !sig_begin || !status || !n_lines_out || !consumed_out) return SQK_ERR_ARG;
    if (n_bytes < 0 || max_lines < 1 || start_col < 0) return SQK_ERR_ARG;
    // ---- 1. cut into lines, find where the signal columns start ------------------------------------------------------
    // The newlines of a window of the text sized for max_lines lines like the first one are found by all threads; the
    // window grows while it holds fewer lines than asked for and there is text left.
#ifdef _OPENMP
    const int nt = host_threads(n_threads);
#else
    const int nt = 1; (void)n_threads;
#endif
    std::vector<int64_t> nlpos;
    {
        const char *first = n_bytes > 0 ? (const char *)memchr(text, '\n', (size_t)n_bytes) : nullptr;
        const int64_t l1 = first ? (first - text) + 1 : n_bytes;
        // lines the sample buffer can take if they all look like the first one (the batch ends there anyway)
        int64_t tabs1 = 0;
        for (int64_t i = 0; i < l1; i++) tabs1 += text[i] == '\t';
        const int64_t fields1 = std::max<int64_t>(1, tabs1 + 1 - start_col);
        const int64_t lines_est = std::min<int64_t>(max_lines, max_samples / fields1 + 2);
        const double want = (double)l1 * (double)lines_est * 1.125 + 65536.0;
        const int64_t step = want > 4e18 ? n_bytes : std::max<int64_t>((int64_t)want, 1 << 20);
        int64_t scanned = 0;
        while (scanned < n_bytes && (int64_t)nlpos.size() < max_lines) {
            const int64_t hi = (n_bytes - scanned <= step) ? n_bytes : scanned + (scanned == 0 ? step : std::max<int64_t>(step / 4, 1 << 20));
            find_newlines(text, scanned, hi, nt, nlpos);
            scanned = hi;
        }
    }
    std::vector<int64_t> line_end;
    int64_t pos = 0, n = 0;
    {
        const int64_t n_nl = std::min<int64_t>((int64_t)nlpos.size(), max_lines);
        line_end.resize((size_t)n_nl);
        for (int64_t i = 0; i < n_nl; i++) {
            line_begin[i] = pos;
            line_end[(size_t)i] = nlpos[(size_t)i];
            pos = nlpos[(size_t)i] + 1;
        }
        n = n_nl;
        if (n < max_lines && (int64_t)nlpos.size() <= n && pos < n_bytes && is_final) {   // unterminated last line
            line_begin[n] = pos;
            line_end.push_back(n_bytes);
            pos = n_bytes;
            n++;
        }
        // (not final: an incomplete last line stays where it is -- the caller brings it back)
    }
    std::vector<int64_t> n_fields((size_t)n);
#pragma omp parallel for schedule(static) num_threads(nt)
    for (int64_t i = 0; i < n; i++) {
        const char *p = text + line_begin[i], *e = text + line_end[(size_t)i];
        int col = 0;
        while (col < start_col && p < e) {
            const char *t = (const char *)memchr(p, '\t', (size_t)(e - p));
            if (!t) { p = e + 1; break; }
            p = t + 1; col++;
        }
        int st = 0;
        int64_t nf = 0;
        if (col < start_col || p > e || (p == e)) { st = SQK_TSV_NO_SIGNAL; p = e; }
        else {
            nf = 1;
            for (const char *q = p; q < e; q++) nf += *q == '\t';
            if (e > p && e[-1] == '\t') nf--;                    // a trailing tab does not open a field
            if (nf <= 0) { st = SQK_TSV_NO_SIGNAL; nf = 0; }
        }
        sig_begin[i] = p - text;
        n_fields[(size_t)i] = nf;
        status[i] = st;
    }
    // ---- 2. sample offsets; stop in front of the line that no longer fits ----------------------------------------------
    int64_t total = 0, used = 0;
    offsets[0] = 0;
    for (; used < n; used++) {
        if (total + n_fields[(size_t)used] > max_samples) break;
        total += n_fields[(size_t)used];
        offsets[used + 1] = total;
    }
    if (used == 0 && n > 0) return SQK_ERR_NOMEM;               // one line is larger than the sample buffer
    // ---- 3. parse -----------------------------------------------------------------------------------------------------
#pragma omp parallel for schedule(dynamic, 16) num_threads(nt)
    for (int64_t i = 0; i < used; i++) {
        if (status[i] & SQK_TSV_NO_SIGNAL) continue;
        status[i] |= parse_fields_i16(text + sig_begin[i], text + line_end[(size_t)i], samples + offsets[i], n_fields[(size_t)i],
                                      /*sentinel=*/line_end[(size_t)i] < n_bytes);
    }
    line_begin[used] = used < n ? line_begin[used] : pos;
    *n_lines_out = used;
    *consumed_out = line_begin[used];
    return SQK_OK;
}

// The first n_cols columns of every parsed line, one line each ("col0 <TAB> col1 <NL>"), gathered into `out`: the caller
// decodes and splits the whole batch in one go instead of slicing line by line.  Returns the bytes written, or minus the
// bytes needed when cap is too small.
int64_t sqk_tsv_heads(const char *text, const int64_t *line_begin, const int64_t *sig_begin, int64_t n_lines, int n_cols,
                      char *out, int64_t cap)
{
    if (!text || !line_begin || !sig_begin || n_lines < 0 || n_cols < 1) return 0;
    int64_t need = 0;
    for (int pass = 0; pass < 2; pass++) {
        char *w = out;
        for (int64_t i = 0; i < n_lines; i++) {
            const char *p = text + line_begin[i];
            const char *e = text + (sig_begin[i] > line_begin[i] ? sig_begin[i] - 1 : line_begin[i]);   // the tab in front of the signal
            const char *q = p;
            int col = 0;
            while (q < e) {
                if (*q == '\t' && ++col == n_cols) break;
                q++;
            }
            while (q > p && (q[-1] == '\r' || q[-1] == '\n')) q--;
            const int64_t len = q - p;
            if (pass == 0) need += len + 1;
            else { memcpy(w, p, (size_t)len); w += len; *w++ = '\n'; }
        }
        if (pass == 0 && (!out || need > cap)) return -need;
    }
    return need;
}

// Python's repr() of a float (= str() of a numpy float64): the shortest digit string that reads back as the same double
// (the closest to it among the shortest), fixed notation with at least ".0" when 1e-4 <= |x| < 1e16, otherwise d.ddde+XX
// with at least two exponent digits (CPython: PyOS_double_to_string(x, 'r', 0, Py_DTSF_ADD_DOT_0), float_repr_style
// "short").  std::to_chars gives exactly those digits (shortest round trip); only the layout is Python's.  Returns the
// length written; buf needs 32 bytes.  tests/test_host.py compares with repr() on ~10^6 doubles of every magnitude.
static int fmt_repr(double x, char *buf)
{
    if (x != x) { memcpy(buf, "nan", 3); return 3; }
    if (isinf(x)) { if (x > 0) { memcpy(buf, "inf", 3); return 3; } memcpy(buf, "-inf", 4); return 4; }
    char *w = buf;
    if (signbit(x)) { *w++ = '-'; x = -x; }
    if (x == 0.0) { memcpy(w, "0.0", 3); return (int)(w - buf) + 3; }
    char e[40];
    {
        // shortest digits that read back as x, closest to x among the shortest (what repr prints), in d.ddde+XX form
        const std::to_chars_result tr = std::to_chars(e, e + sizeof(e) - 1, x, std::chars_format::scientific);
        *tr.ptr = 0;
    }
    char *ep = strchr(e, 'e');
    const int exp10 = atoi(ep + 1);
    char dig[20];
    int nd = 0;
    for (const char *q = e; q < ep; q++) if (*q != '.') dig[nd++] = *q;
    while (nd > 1 && dig[nd - 1] == '0') nd--;
    if (exp10 >= -4 && exp10 < 16) {
        if (exp10 < 0) {                                      // 0.000ddd
            *w++ = '0'; *w++ = '.';
            for (int i = 0; i < -exp10 - 1; i++) *w++ = '0';
            memcpy(w, dig, (size_t)nd); w += nd;
        } else {
            for (int i = 0; i <= exp10; i++) *w++ = i < nd ? dig[i] : '0';
            *w++ = '.';
            if (nd > exp10 + 1) { memcpy(w, dig + exp10 + 1, (size_t)(nd - exp10 - 1)); w += nd - exp10 - 1; }
            else *w++ = '0';
        }
    } else {
        *w++ = dig[0];
        if (nd > 1) { *w++ = '.'; memcpy(w, dig + 1, (size_t)(nd - 1)); w += nd - 1; }
        *w++ = 'e';
        *w++ = exp10 < 0 ? '-' : '+';
        const int ae = exp10 < 0 ? -exp10 : exp10;
        if (ae >= 100) *w++ = (char)('0' + ae / 100);
        *w++ = (char)('0' + (ae / 10) % 10);
        *w++ = (char)('0' + ae % 10);
    }
    return (int)(w - buf);
}

static int fmt_int(int v, char *buf)
{
    char tmp[12];
    int nd = 0, n = 0;
    unsigned u = v < 0 ? (unsigned)(-(int64_t)v) : (unsigned)v;
    if (v < 0) buf[n++] = '-';
    do { tmp[nd++] = (char)('0' + u % 10); u /= 10; } while (u);
    while (nd) buf[n++] = tmp[--nd];
    return n;
}

// ---- the score columns of a MotifSeq row (MotifSeq.py:441-445) ----------------------------------------------------------
// Z = (dist - mod_mean) / mod_stdev; p = scipy.stats.norm.cdf(Z); hit_P = (1 - p) * 100.  scipy evaluates norm.cdf with
// its ndtr (scipy/special: the Cephes ndtr.c algorithm, now in its xsf headers): x = Z / sqrt(2), 0.5 + 0.5 erf(x) for
// |x| < 1, else 0.5 erfc(|x|) reflected; erf and erfc are Cephes' rational approximations (coefficient tables P/Q for
// 1 <= x < 8, R/S beyond, T/U for erf below 1), not libm's.  Restated here so that the command line does not have to import
// scipy (0.35-0.75 s per process); tests/test_tsv.py checks bit equality with scipy.special.ndtr on millions of values.
// exp() is libm's, as in scipy's build; no FMA contraction (the Makefile passes -ffp-contract=off for this file).
static const double ND_P[] = {2.46196981473530512524E-10, 5.64189564831068821977E-1, 7.46321056442269912687E0,
                              4.86371970985681366614E1, 1.96520832956077098242E2, 5.26445194995477358631E2,
                              9.34528527171957607540E2, 1.02755188689515710272E3, 5.57535335369399327526E2};
static const double ND_Q[] = {1.32281951154744992508E1, 8.67072140885989742329E1, 3.54937778887819891062E2,
                              9.75708501743205489753E2, 1.82390916687909736289E3, 2.24633760818710981792E3,
                              1.65666309194161350182E3, 5.57535340817727675546E2};
static const double ND_R[] = {5.64189583547755073984E-1, 1.27536670759978104416E0, 5.01905042251180477414E0,
                              6.16021097993053585195E0, 7.40974269950448939160E0, 2.97886665372100240670E0};
static const double ND_S[] = {2.26052863220117276590E0, 9.39603524938001434673E0, 1.20489539808096656605E1,
                              1.70814450747565897222E1, 9.60896809063285878198E0, 3.36907645100081516050E0};
static const double ND_T[] = {9.60497373987051638749E0, 9.00260197203842689217E1, 2.23200534594684319226E3,
                              7.00332514112805075473E3, 5.55923013010394962768E4};
static const double ND_U[] = {3.35617141647503099647E1, 5.21357949780152679795E2, 4.59432382970980127987E3,
                              2.26290000613890934246E4, 4.92673942608635921086E4};
static inline double nd_polevl(double x, const double *c, int n) { double a = c[0]; for (int i = 1; i <= n; i++) a = a * x + c[i]; return a; }
static inline double nd_p1evl(double x, const double *c, int n) { double a = x + c[0]; for (int i = 1; i < n; i++) a = a * x + c[i]; return a; }
static double nd_erf(double x);
static double nd_erfc(double a)
{
    if (a != a) return a;
    const double x = a < 0.0 ? -a : a;
    if (x < 1.0) return 1.0 - nd_erf(a);
    double z = -a * a;
    if (z < -7.09782712893383996843E2) return a < 0 ? 2.0 : 0.0;      // exp underflows
    z = exp(z);
    double p, q;
    if (x < 8.0) { p = nd_polevl(x, ND_P, 8); q = nd_p1evl(x, ND_Q, 8); }
    else { p = nd_polevl(x, ND_R, 5); q = nd_p1evl(x, ND_S, 6); }
    double y = (z * p) / q;
    if (a < 0) y = 2.0 - y;
    if (y == 0.0) return a < 0 ? 2.0 : 0.0;
    return y;
}
static double nd_erf(double x)
{
    if (x != x) return x;
    if (x < 0.0) return -nd_erf(-x);
    if (x > 1.0) return 1.0 - nd_erfc(x);
    const double z = x * x;
    return x * nd_polevl(z, ND_T, 4) / nd_p1evl(z, ND_U, 5);
}
static double nd_ndtr(double a)
{
    if (a != a) return a;
    const double x = a * 7.07106781186547524401E-1, z = fabs(x);
    if (z < 1.0) return 0.5 + 0.5 * nd_erf(x);
    const double y = 0.5 * nd_erfc(z);
    return x > 0 ? 1.0 - y : y;
}

void sqk_ndtr(const double *z, int64_t n, double *out)
{
    if (!z || !out) return;
    for (int64_t i = 0; i < n; i++) out[i] = nd_ndtr(z[i]);
}

// Z-score, p-value and hit probability of every (read, model) record: hits [n_reads][n_models] sqk_hit, mod_mean / mod_stdev
// per model; zs / ps / hps [n_reads][n_models].
void sqk_score_hits(const void *hits_v, int64_t n_reads, int n_models, const double *mod_mean, const double *mod_stdev,
                    int n_threads, double *zs, double *ps, double *hps)
{
    if (!hits_v || !mod_mean || !mod_stdev || !zs || !ps || !hps || n_reads <= 0 || n_models <= 0) return;
    const sqk_hit *hits = (const sqk_hit *)hits_v;
    // (~20 ns per record: a team of threads only pays for itself on large batches)
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(host_threads(n_threads), n_reads * n_models / 65536));
    (void)nt;
#pragma omp parallel for schedule(static) num_threads(nt)
    for (int64_t r = 0; r < n_reads; r++)
        for (int m = 0; m < n_models; m++) {
            const int64_t i = r * n_models + m;
            const double Z = (hits[i].dist - mod_mean[m]) / mod_stdev[m];
            const double p = nd_ndtr(Z);
            zs[i] = Z; ps[i] = p; hps[i] = (1.0 - p) * 100.0;
        }
}

// The rows get_region_multi prints (MotifSeq.py:441-449) for a batch, as text:
//   head <TAB> name <TAB> start <TAB> end <TAB> end-start <TAB> dist <TAB> consts <TAB> Z <TAB> p <TAB> hit_P <NL>
// heads: "fast5 <TAB> readID <NL>" per read (sqk_tsv_heads); names / consts ("mod_mean <TAB> mod_stdev" already as text):
// per model, NUL-separated; hits [n_reads][n_models] (start, end, dist); zs / ps / hps [n_reads][n_models].  Reads whose
// hit is a status (start < 0) print nothing.  Floats are written as Python's repr() writes them.  Returns the bytes
// written, or minus an upper bound of the bytes needed when cap is too small.
int64_t sqk_tsv_format_rows(const char *heads, int64_t n_reads, const void *hits_v, int n_models, const char *names,
                            const char *consts, const double *zs, const double *ps, const double *hps, int n_threads,
                            char *out, int64_t cap)
{
    struct Hit { int32_t start, end; double dist; };
    const Hit *hits = (const Hit *)hits_v;
    if (!heads || !hits || !names || !consts || !zs || !ps || !hps || n_reads < 0 || n_models < 1) return 0;
    std::vector<const char *> nm((size_t)n_models), cs((size_t)n_models);
    std::vector<size_t> nml((size_t)n_models), csl((size_t)n_models);
    { const char *p = names, *q = consts;
      for (int m = 0; m < n_models; m++) { nm[(size_t)m] = p; nml[(size_t)m] = strlen(p); p += nml[(size_t)m] + 1;
                                           cs[(size_t)m] = q; csl[(size_t)m] = strlen(q); q += csl[(size_t)m] + 1; } }
    // where each read's head starts, and an upper bound of each read's text
    std::vector<int64_t> hb((size_t)n_reads + 1), at((size_t)n_reads + 1);
    { const char *p = heads;
      for (int64_t r = 0; r < n_reads; r++) { hb[(size_t)r] = p - heads; const char *nl = strchr(p, '\n'); p = nl ? nl + 1 : p + strlen(p); }
      hb[(size_t)n_reads] = p - heads; }
    at[0] = 0;
    for (int64_t r = 0; r < n_reads; r++) {
        int64_t need = 0;
        for (int m = 0; m < n_models; m++) need += (hb[(size_t)r + 1] - hb[(size_t)r]) + (int64_t)nml[(size_t)m] + (int64_t)csl[(size_t)m] + 3 * 12 + 4 * 26 + 12;
        at[(size_t)r + 1] = at[(size_t)r] + need;
    }
    if (!out || at[(size_t)n_reads] > cap) return -at[(size_t)n_reads];
    std::vector<int64_t> len((size_t)n_reads);
    // (a row costs ~0.3 us: a team of threads only pays for itself on batches of several hundred thousand rows)
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(host_threads(n_threads), n_reads * n_models / 131072));
#pragma omp parallel for schedule(static) num_threads(nt)
    for (int64_t r = 0; r < n_reads; r++) {
        char *w = out + at[(size_t)r];
        const int64_t hl = hb[(size_t)r + 1] - hb[(size_t)r] - 1;        // without the newline
        for (int m = 0; m < n_models; m++) {
            const Hit &h = hits[r * n_models + m];
            if (h.start < 0) continue;
            memcpy(w, heads + hb[(size_t)r], (size_t)(hl > 0 ? hl : 0)); w += hl > 0 ? hl : 0;
            *w++ = '\t'; memcpy(w, nm[(size_t)m], nml[(size_t)m]); w += nml[(size_t)m];
            *w++ = '\t'; w += fmt_int(h.start, w);
            *w++ = '\t'; w += fmt_int(h.end, w);
            *w++ = '\t'; w += fmt_int(h.end - h.start, w);
            *w++ = '\t'; w += fmt_repr(h.dist, w);
            *w++ = '\t'; memcpy(w, cs[(size_t)m], csl[(size_t)m]); w += csl[(size_t)m];
            *w++ = '\t'; w += fmt_repr(zs[r * n_models + m], w);
            *w++ = '\t'; w += fmt_repr(ps[r * n_models + m], w);
            *w++ = '\t'; w += fmt_repr(hps[r * n_models + m], w);
            *w++ = '\n';
        }
        len[(size_t)r] = w - (out + at[(size_t)r]);
    }
    // close the gaps (each read was given an upper bound of room)
    int64_t total = 0;
    for (int64_t r = 0; r < n_reads; r++) {
        if (at[(size_t)r] != total) memmove(out + total, out + at[(size_t)r], (size_t)len[(size_t)r]);
        total += len[(size_t)r];
    }
    return total;
}

// The rows segmenter.py prints (segmenter.py:130-146: name <TAB> s0,e0,s1,e1,...) for the reads with keep[r] != 0.
// heads: "name <NL>" per read (sqk_tsv_heads with one column); segs [n_reads][max_segs][2], n_segs [n_reads].  Returns the
// bytes written, or minus an upper bound of the bytes needed when cap is too small (out may be NULL to ask).
int64_t sqk_tsv_format_segs(const char *heads, int64_t n_reads, const int32_t *segs, const int32_t *n_segs, int max_segs,
                            const unsigned char *keep, int n_threads, char *out, int64_t cap)
{
    if (!heads || !segs || !n_segs || !keep || n_reads < 0 || max_segs < 1) return 0;
    std::vector<int64_t> hb((size_t)n_reads + 1), at((size_t)n_reads + 1), len((size_t)n_reads);
    { const char *p = heads;
      for (int64_t r = 0; r < n_reads; r++) { hb[(size_t)r] = p - heads; const char *nl = strchr(p, '\n'); p = nl ? nl + 1 : p + strlen(p); }
      hb[(size_t)n_reads] = p - heads; }
    at[0] = 0;
    for (int64_t r = 0; r < n_reads; r++) {
        const int n = n_segs[r] < max_segs ? n_segs[r] : max_segs;
        at[(size_t)r + 1] = at[(size_t)r] + (keep[r] ? (hb[(size_t)r + 1] - hb[(size_t)r]) + 2 + 24ll * (n > 0 ? n : 0) : 0);
    }
    if (!out || at[(size_t)n_reads] > cap) return -at[(size_t)n_reads];
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(host_threads(n_threads), n_reads / 131072));
#pragma omp parallel for schedule(static) num_threads(nt)
    for (int64_t r = 0; r < n_reads; r++) {
        char *w = out + at[(size_t)r];
        if (keep[r]) {
            const int64_t hl = hb[(size_t)r + 1] - hb[(size_t)r] - 1;
            memcpy(w, heads + hb[(size_t)r], (size_t)(hl > 0 ? hl : 0)); w += hl > 0 ? hl : 0;
            *w++ = '\t';
            const int n = n_segs[r] < max_segs ? n_segs[r] : max_segs;
            for (int i = 0; i < n; i++) {
                if (i) *w++ = ',';
                w += fmt_int(segs[(r * max_segs + i) * 2], w);
                *w++ = ',';
                w += fmt_int(segs[(r * max_segs + i) * 2 + 1], w);
            }
            *w++ = '\n';
        }
        len[(size_t)r] = w - (out + at[(size_t)r]);
    }
    int64_t total = 0;
    for (int64_t r = 0; r < n_reads; r++) {
        if (len[(size_t)r] && at[(size_t)r] != total) memmove(out + total, out + at[(size_t)r], (size_t)len[(size_t)r]);
        total += len[(size_t)r];
    }
    return total;
}

// "fast5 \t readID \t s0 \t s1 ...\n" per read (SquigglePull.py:251-253).  heads: the text in front of the signal columns of
// each read ("fast5\treadID" or with the four extra_info columns), concatenated, head_offsets[n_reads + 1].  Returns the
// bytes written, or the (negative) bytes needed when `cap` is too small.
int64_t sqk_tsv_format(const int16_t *samples, const int64_t *offsets, int64_t n_reads, const char *heads,
                       const int64_t *head_offsets, int n_threads, char *out, int64_t cap)
{
    if (!samples || !offsets || !heads || !head_offsets || n_reads < 0) return 0;
    std::vector<int64_t> at((size_t)n_reads + 1);
    at[0] = 0;
#ifdef _OPENMP
    const int nt = host_threads(n_threads);
#else
    const int nt = 1; (void)n_threads;
#endif
    // exact length of every line first
#pragma omp parallel for schedule(static) num_threads(nt)
    for (int64_t r = 0; r < n_reads; r++) {
        int64_t len = head_offsets[r + 1] - head_offsets[r] + 1;    // head + newline
        for (int64_t i = offsets[r]; i < offsets[r + 1]; i++) {
            int v = samples[i];
            int d = 1 + (v < 0);                                  // tab + sign ... (the tab in front of every sample)
            if (v < 0) v = -v;
            d += v >= 10000 ? 5 : v >= 1000 ? 4 : v >= 100 ? 3 : v >= 10 ? 2 : 1;
            len += d;
        }
        at[(size_t)r + 1] = len;
    }
    for (int64_t r = 0; r < n_reads; r++) at[(size_t)r + 1] += at[(size_t)r];
    const int64_t need = at[(size_t)n_reads];
    if (!out || need > cap) return -need;
#pragma omp parallel for schedule(static) num_threads(nt)
    for (int64_t r = 0; r < n_reads; r++) {
        char *w = out + at[(size_t)r];
        const int64_t hl = head_offsets[r + 1] - head_offsets[r];
        memcpy(w, heads + head_offsets[r], (size_t)hl);
        w += hl;
        for (int64_t i = offsets[r]; i < offsets[r + 1]; i++) {
            int v = samples[i];
            *w++ = '\t';
            if (v < 0) { *w++ = '-'; v = -v; }
            char tmp[6];
            int nd = 0;
            do { tmp[nd++] = (char)('0' + v % 10); v /= 10; } while (v);
            while (nd) *w++ = tmp[--nd];
        }
        *w++ = '\n';
    }
    return need;
}

}   // extern "C"
