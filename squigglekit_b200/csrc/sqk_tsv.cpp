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
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "../../include/sqk.h"

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
    const int nt = n_threads > 0 ? n_threads : omp_get_max_threads();
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
    const int nt = n_threads > 0 ? n_threads : omp_get_max_threads();
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
