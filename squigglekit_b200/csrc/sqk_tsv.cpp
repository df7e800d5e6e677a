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

// parse one line's signal part [p, e) into out[0..n_fields); returns the status bits
static int parse_fields_i16(const char *p, const char *e, int16_t *out, int64_t n_fields)
{
    int status = 0;
    bool any_nonzero = false;
    int64_t k = 0;
    while (p <= e && k < n_fields) {
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
            any_nonzero |= v != 0;
        }
        k++;
        p = q + 1;
    }
    if (!(status & SQK_TSV_NOT_INT16) && !any_nonzero) status |= SQK_TSV_ALL_ZERO;
    return status;
}

extern "C" {

int sqk_tsv_parse(const char *text, int64_t n_bytes, int is_final, int start_col, int64_t max_lines, int64_t max_samples,
                  int n_threads, int16_t *samples, int64_t *offsets, int64_t *line_begin, int64_t *sig_begin,
                  int32_t *status, int64_t *n_lines_out, int64_t *consumed_out)
{
    if (!text || !samples || !offsets || !line_begin || !sig_begin || !status || !n_lines_out || !consumed_out) return SQK_ERR_ARG;
    if (n_bytes < 0 || max_lines < 1 || start_col < 0) return SQK_ERR_ARG;
    // ---- 1. cut into lines, find where the signal columns start ------------------------------------------------------
    std::vector<int64_t> line_end;
    line_end.reserve((size_t)std::min<int64_t>(max_lines, 1 << 20));
    int64_t pos = 0, n = 0;
    while (pos < n_bytes && n < max_lines) {
        const char *nl = (const char *)memchr(text + pos, '\n', (size_t)(n_bytes - pos));
        int64_t end;
        if (nl) end = nl - text;
        else if (is_final) end = n_bytes;
        else break;                                             // incomplete last line: the caller brings it back
        line_begin[n] = pos;
        line_end.push_back(end);
        pos = nl ? end + 1 : end;
        n++;
    }
#ifdef _OPENMP
    const int nt = host_threads(n_threads);
#else
    const int nt = 1; (void)n_threads;
#endif
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
        status[i] |= parse_fields_i16(text + sig_begin[i], text + line_end[(size_t)i], samples + offsets[i], n_fields[(size_t)i]);
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

// Python's repr() of a float (= str() of a numpy float64): the shortest digit string that reads back as the same double,
// fixed notation with at least ".0" when 1e-4 <= |x| < 1e16, otherwise d.ddde+XX with at least two exponent digits
// (CPython: PyOS_double_to_string(x, 'r', 0, Py_DTSF_ADD_DOT_0), float_repr_style "short").  The shortest string has at
// most 17 significant digits.  If it has <= 15, the correctly rounded 15-digit decimal IS that string padded with zeros
// (a normal double lies within 1.2e-16 relative of it, less than half a unit of the 15th digit), so: print 15 digits, read
// back, strip zeros; else the correctly rounded 16-digit decimal if it reads back (the closest of the 16-digit candidates,
// which is the one repr picks), else the correctly rounded 17-digit one.  Returns the length written; buf needs 32 bytes.
static int fmt_repr(double x, char *buf)
{
    if (x != x) { memcpy(buf, "nan", 3); return 3; }
    if (isinf(x)) { if (x > 0) { memcpy(buf, "inf", 3); return 3; } memcpy(buf, "-inf", 4); return 4; }
    char *w = buf;
    if (signbit(x)) { *w++ = '-'; x = -x; }
    if (x == 0.0) { memcpy(w, "0.0", 3); return (int)(w - buf) + 3; }
    char e[40];
    if (x < 2.3e-308) {                                       // subnormal: few significant bits, try every length
        for (int p = 0; p <= 16; p++) { snprintf(e, sizeof(e), "%.*e", p, x); if (strtod(e, nullptr) == x) break; }
    } else {
        snprintf(e, sizeof(e), "%.14e", x);                   // d.dddddddddddddde+XX
        if (strtod(e, nullptr) != x) {
            snprintf(e, sizeof(e), "%.15e", x);
            if (strtod(e, nullptr) != x) snprintf(e, sizeof(e), "%.16e", x);
        }
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
    const int nt = host_threads(n_threads);
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
    const int nt = host_threads(n_threads);
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
