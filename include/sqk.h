/*
 * sqk.h -- C ABI of libsqk.so, the B200 (sm_100a) implementation of SquiggleKit's two
 * signal-analysis hot paths.  Plain C: pointers and sizes only, no C++/torch types, no
 * exceptions across the boundary.  Every entry point cites the reference code it replaces
 * (paths relative to the Psy-Fer/SquiggleKit tree).
 *
 * The reference has no FFI of its own: the boundary is two Python call sites,
 *     dist, cost, path = dtw_subsequence(model[name], sig)        MotifSeq.py:437
 *     segs = get_segs(sig, args)                                   segmenter.py:128,159,211,242,273
 * plus the per-read preparation in front of them (outlier removal, normalisation,
 * truncation).  libsqk batches those calls over reads.  INTEGRATION.md shows the ctypes stub
 * a maintainer of the reference would add.
 *
 * Conventions
 *   - signals: all reads concatenated, int16 raw DAC samples; offsets[n_reads+1] are sample
 *     indices into it (ragged reads, zero-length reads allowed); offsets[0] need not be 0.
 *   - mem = SQK_MEM_HOST:   every pointer is host memory; the call is synchronous; reads are
 *                           streamed to the GPU in chunks, copies overlapped with compute
 *                           (pinned host memory -- sqk_host_alloc -- makes the overlap real).
 *     mem = SQK_MEM_DEVICE: every pointer is device memory on the ctx's device; the call only
 *                           enqueues work on the ctx stream (sqk_ctx_set_stream / sqk_ctx_sync).
 *   - return value: 0 = ok, <0 = sqk_status; message via sqk_last_error() (thread-local).
 *   - a ctx is bound to one device and is not thread-safe; use one ctx per host thread.  Device-mode calls
 *     share the ctx's scratch memory in stream order; the library orders calls that arrive on different streams
 *     (and host-mode calls behind device-mode ones) with events.
 *   - the caller owns every buffer passed in; the library keeps no reference after return
 *     (device mode: after the stream work has completed).
 *   - there is NO CPU fallback: without a CUDA device sqk_ctx_create fails.
 */
#ifndef SQK_H
#define SQK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SQK_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define SQK_API __attribute__((visibility("default")))
#else
#define SQK_API
#endif

enum sqk_status {
    SQK_OK = 0,
    SQK_ERR_ARG = -1,         /* bad argument */
    SQK_ERR_CUDA = -2,        /* CUDA runtime error (message has the cudaError string) */
    SQK_ERR_NOMEM = -3,       /* host or device allocation failed */
    SQK_ERR_UNSUPPORTED = -4, /* valid request this build cannot serve (e.g. motif too long) */
    SQK_ERR_OVERFLOW = -5     /* an output capacity given by the caller was exceeded */
};

enum sqk_mem { SQK_MEM_HOST = 0, SQK_MEM_DEVICE = 1 };

/* MotifSeq.py:96  --scale {zscale, medmad};  "none" feeds already-centred data through */
enum sqk_scale { SQK_SCALE_ZSCALE = 0, SQK_SCALE_MEDMAD = 1, SQK_SCALE_NONE = 2 };

/* DTW arithmetic.  FP64 reproduces mlpy's float64 recurrence bit for bit (indices and dist).  FP32 asks for "dist within
 * 1e-4, indices may differ on near-ties" (BASELINE.json north_star); since the exact two-pass plan (float32 lower-bound
 * scan + float64 windows) is faster than a float32 recurrence with start pointers ever was, FP32 requests are served by
 * the exact path as well: same results as FP64, which satisfies the looser contract. */
enum sqk_precision { SQK_PREC_FP64 = 0, SQK_PREC_FP32 = 1 };

typedef struct sqk_ctx sqk_ctx;

SQK_API int sqk_version(void);
SQK_API const char *sqk_last_error(void);

SQK_API int sqk_ctx_create(int device, sqk_ctx **out);
SQK_API int sqk_ctx_destroy(sqk_ctx *ctx);
/* Run device-mode calls on the caller's CUDA stream (a cudaStream_t, e.g. torch's current stream).  NULL is a
 * stream like any other: the legacy default stream (what torch's default stream is).  sqk_ctx_reset_stream goes
 * back to the ctx's own stream.  Calls arriving on a different stream than the previous device-mode call are
 * ordered behind it with an event, so the shared ctx scratch is never raced. */
SQK_API int sqk_ctx_set_stream(sqk_ctx *ctx, void *cuda_stream);
SQK_API int sqk_ctx_reset_stream(sqk_ctx *ctx);
SQK_API int sqk_ctx_sync(sqk_ctx *ctx);
SQK_API int sqk_device_count(int *count);
/* multiProcessorCount etc. of the ctx device: props[0]=SMs, [1]=max smem/block (bytes),
 * [2]=SM clock kHz, [3]=L2 bytes, [4]=compute capability major*10+minor */
SQK_API int sqk_ctx_device_props(sqk_ctx *ctx, int64_t props[5]);

/* Pinned host memory for mem=SQK_MEM_HOST callers that want copy/compute overlap. */
SQK_API int sqk_host_alloc(uint64_t bytes, void **out);
SQK_API int sqk_host_free(void *p);

/* ---------------------------------------------------------------------------------------
 * MotifSeq hot path.  Per read, per model -- exactly the body of the reference's main loop:
 *     sig = scale_outliers(sig)                 MotifSeq.py:317-324  keep lo < s < hi
 *     sig = zscale | medmad                     MotifSeq.py:186-200
 *     dist, cost, path = dtw_subsequence(model, sig)                 MotifSeq.py:437
 *     start = path[1][0]; end = path[1][-1]                          MotifSeq.py:438-439
 * Indices are in the post-outlier-removal index space, as in the reference.
 * ------------------------------------------------------------------------------------- */
typedef struct {
    int32_t scale_mode; /* enum sqk_scale */
    int32_t lo, hi;     /* --scale_low / --scale_hi, MotifSeq.py:118-121 (defaults 0, 1200) */
    int32_t precision;  /* enum sqk_precision */
} sqk_motif_params;

typedef struct {
    int32_t start; /* path[1][0];  -1: read empty after outlier removal; -2: scale undefined (MAD == 0);
                      -3: read longer than the max_read_len the caller declared (not processed) */
    int32_t end;   /* path[1][-1] == np.argmin(cost[-1, :]) */
    double dist;   /* cost[-1, end]; NaN when start < 0 */
} sqk_hit;

/*
 * models: n_models expanded motifs (float64, what read_synth_model MotifSeq.py:354-379 returns),
 * concatenated; model_offsets[n_models+1].  models, model_offsets and params are ALWAYS host
 * pointers (they are a few hundred bytes); `mem` describes signals, offsets, hits and n_kept.  hits is [n_reads][n_models] (read-major, the order
 * get_region_multi MotifSeq.py:436 prints).  n_kept (optional, may be NULL): post-outlier length
 * of each read.  max_read_len: longest read in samples if the caller knows it, 0 = let the
 * library find out (costs one device sync in device mode).
 */
SQK_API int sqk_motifseq(sqk_ctx *ctx, const int16_t *signals, const int64_t *offsets, int64_t n_reads,
                 int64_t max_read_len, const double *models, const int32_t *model_offsets, int32_t n_models,
                 const sqk_motif_params *params, int mem, sqk_hit *hits, int32_t *n_kept);

/* cost[-1, :] of ONE read against ONE model (what view_region plots, MotifSeq.py:507-509) and its
 * normalised signal (what -x prints, MotifSeq.py:447).  Host pointers.  last_row / norm_sig have
 * capacity `cap` doubles each (either may be NULL); *n_out = post-outlier length. */
SQK_API int sqk_motifseq_trace(sqk_ctx *ctx, const int16_t *signal, int64_t n_samples, const double *model,
                       int32_t n_model, const sqk_motif_params *params, double *last_row, double *norm_sig,
                       int64_t cap, int64_t *n_out, sqk_hit *hit);

/* ---------------------------------------------------------------------------------------
 * segmenter hot path.  Per read -- the body of the reference's main loop:
 *     sig = sig[:Num]   (Num == 0 -> -1: drops the last sample)      segmenter.py:104-105,124,207
 *     sig = scale_outliers(sig)                                      segmenter.py:311-318
 *     segs = get_segs(sig, args)                                     segmenter.py:399-470
 * test_segs (segmenter.py:473-494) is a two-comparison filter on the result and stays on the host.
 * ------------------------------------------------------------------------------------- */
typedef struct {
    int32_t error;      /* -e, default 5    */
    int32_t corrector;  /* -c, default 50   */
    int32_t window;     /* -w, default 150  */
    int32_t seg_dist;   /* -d, default 50   */
    double std_scale;   /* -t, default 0.75 */
    double stall_len;   /* -l, default 0.25 */
    int32_t lim_lo;     /* -lim_low, default 0   */
    int32_t lim_hi;     /* -lim_hi,  default 900 */
    int32_t num;        /* -n: 0 = all (reference quirk: drops the last sample), >0 first n, <0 python slice */
    int32_t max_segs;   /* capacity of segs per read */
} sqk_seg_params;

/* segs: [n_reads][max_segs][2] int32 (start, end); n_segs[n_reads]: segments found (0 == the
 * reference's `False`); a count above max_segs means the row was truncated to max_segs; -1 = read longer than the
 * declared max_read_len (not processed). */
SQK_API int sqk_segmenter(sqk_ctx *ctx, const int16_t *signals, const int64_t *offsets, int64_t n_reads,
                  int64_t max_read_len, const sqk_seg_params *params, int mem, int32_t *segs, int32_t *n_segs);

/* Same, on picoamperes: the reference's default for fast5 input converts each read before segmenting,
 *     pA = np.round((raw + offset) * (range / digitisation), 2)        segmenter.py:345-349, 366-370, 515-517
 * (range itself pre-rounded to 2 decimals, :344).  pa_offset[n_reads] = channel offset, pa_scale[n_reads] =
 * range / digitisation, both float64, following `mem`; pa_scale must be > 0.  lim_lo / lim_hi and the
 * thresholds then apply to the pA values; segment positions are identical to the reference's. */
SQK_API int sqk_segmenter_pa(sqk_ctx *ctx, const int16_t *signals, const int64_t *offsets, int64_t n_reads,
                             int64_t max_read_len, const double *pa_offset, const double *pa_scale,
                             const sqk_seg_params *params, int mem, int32_t *segs, int32_t *n_segs);

/* ---------------------------------------------------------------------------------------
 * dRNA adapter finder -- the slow5 branch of dRNA_segmenter.py (dRNA_segmenter.py:86-176).  Per read:
 *     sig = scale_outliers(signal)                    keep 0 < s < 1200              :88, :331-334
 *     median, stdev of sig[t_start:t_end]; top = median + stdev * 0.8               :104-106
 *     one-sided run detector  a < top  with the constants of :89-100                 :108-166
 *     print the first segment                                                        :171-174
 * The reference hard-codes every constant; they are parameters here with those defaults
 * (SQK_ADAPTER_DEFAULTS).  segs: [n_reads][2] int32 (start, end) in post-outlier coordinates; found[n_reads]:
 * 1 = a segment was found, 0 = none (the reference prints nothing), -1 = read longer than the declared
 * max_read_len.  (The TSV branch of that script, :272-326, cannot run as shipped -- `w` is undefined -- and is
 * not provided.)
 * ------------------------------------------------------------------------------------- */
typedef struct {
    int32_t error;          /* 5    :90  */
    int32_t no_err_thresh;  /* 2500 :91  tolerated samples count as errors only from this position on */
    int32_t corrector;      /* 1200 :96  (w) */
    int32_t window;         /* 100  :97  */
    int32_t seg_dist;       /* 1200 :99  */
    int32_t t_start, t_end; /* 1000, 5000 :82-83: the slice the threshold statistics come from */
    double std_scale;       /* 0.8  :106 */
    int32_t lim_lo, lim_hi; /* 0, 1200 :333 */
} sqk_adapter_params;
#define SQK_ADAPTER_DEFAULTS {5, 2500, 1200, 100, 1200, 1000, 5000, 0.8, 0, 1200}

SQK_API int sqk_adapter(sqk_ctx *ctx, const int16_t *signals, const int64_t *offsets, int64_t n_reads,
                        int64_t max_read_len, const sqk_adapter_params *params, int mem, int32_t *segs, int32_t *found);

/* ---------------------------------------------------------------------------------------
 * The rolling-mean adapter finder of dRNA_segmenter.py's TSV branch (dRNA_segmenter.py:272-326):
 *   scale_outliers (lim_lo < s < lim_hi, :278, :331-334) -> t = pd.Series(sig).rolling(window=w).mean() (:281-282)
 *   -> bot = t.mean() - t.std() * std_factor (:283-287, pandas nanops: NaN slots skipped, ddof = 1)
 *   -> runs of t < bot closed by t > bot, merged when closer than seg_dist (:291-313)
 *   -> the first segment with lo_thresh <= end - start <= hi_thresh, printed as (start - shift, end - shift) (:315-323).
 * As shipped the branch raises NameError: `w` only exists in the comment `# w = 2000` (:81); it is a parameter here.
 * segs[r] = (start - shift, end - shift), found[r] = 1, or (0, 0) and 0.  max_read_len as for sqk_segmenter.
 * w must be in [1, 65536].
 * ------------------------------------------------------------------------------------- */
typedef struct sqk_rollmean_params {
    int32_t w;              /* 2000   :81  */
    int32_t seg_dist;       /* 1500   :292 */
    int32_t lo_thresh;      /* 2000   :294 */
    int32_t hi_thresh;      /* 200000 :293 */
    int32_t shift;          /* 1000   :320 */
    int32_t lim_lo, lim_hi; /* 0, 1200 :333 */
    int32_t reserved;
    double std_factor;      /* 0.5    :287 */
} sqk_rollmean_params;
#define SQK_ROLLMEAN_DEFAULTS {2000, 1500, 2000, 200000, 1000, 0, 1200, 0, 0.5}

SQK_API int sqk_rollmean(sqk_ctx *ctx, const int16_t *signals, const int64_t *offsets, int64_t n_reads,
                         int64_t max_read_len, const sqk_rollmean_params *params, int mem, int32_t *segs, int32_t *found);

/* ---------------------------------------------------------------------------------------
 * float64 signals.  The reference's `-s` path hands both tools whatever the TSV holds (MotifSeq.py:270
 * `float(i) for i in l[8:]`; segmenter.py:198-199), e.g. SquigglePull's pA output.  Same semantics and outputs as
 * sqk_motifseq / sqk_segmenter with `signals` as float64: outlier window, numpy-exact float statistics (pairwise
 * mean / sigma, exact medians), then the same DTW / state-machine kernels.  Device mode needs one small
 * synchronising copy (to size its scratch).  This path costs 18 bytes of HBM traffic per sample instead of 2.
 * ------------------------------------------------------------------------------------- */
SQK_API int sqk_motifseq_f64(sqk_ctx *ctx, const double *signals, const int64_t *offsets, int64_t n_reads,
                             const double *models, const int32_t *model_offsets, int32_t n_models,
                             const sqk_motif_params *params, int mem, sqk_hit *hits, int32_t *n_kept);
SQK_API int sqk_segmenter_f64(sqk_ctx *ctx, const double *signals, const int64_t *offsets, int64_t n_reads,
                              const sqk_seg_params *params, int mem, int32_t *segs, int32_t *n_segs);

/* ---------------------------------------------------------------------------------------
 * SquigglePull signal text (host side, no GPU work): the `-s` input of both command lines.
 *   fast5 <TAB> readID [<TAB> digitisation <TAB> offset <TAB> range <TAB> sampling_rate] <TAB> s0 <TAB> s1 ...
 * written by SquigglePull.py:243-253 (print_data), read by MotifSeq.py:252-298 (signal from column 8) and
 * segmenter.py:179-230 (signal from column 4) one float() / int() per field.
 *
 * sqk_tsv_parse cuts `text` into lines and parses the signal fields of every line (from column start_col on) into the
 * int16 batch layout of sqk_motifseq / sqk_segmenter, in parallel over the lines.  Per line i: line_begin[i] (byte offset
 * of the line), sig_begin[i] (byte offset of the first signal field: the head columns are text[line_begin[i] ..
 * sig_begin[i] - 1)), offsets[i .. i+1] (its samples), status[i] (flags below; a flagged line's samples are not valid and
 * the caller sends the line through its float path).  Stops in front of an incomplete last line (unless is_final), after
 * max_lines lines, or in front of the line that would overflow max_samples.  *n_lines = lines parsed, *consumed = bytes
 * used.  n_threads <= 0: all host threads.
 *
 * sqk_tsv_format writes "<head> <TAB> s0 <TAB> s1 ... <NL>" per read (heads concatenated, head_offsets[n_reads + 1]).
 * Returns the bytes written, or minus the bytes needed when cap is too small (out may be NULL to ask).
 * ------------------------------------------------------------------------------------- */
#define SQK_TSV_NO_SIGNAL 1   /* fewer than start_col + 1 columns                                  */
#define SQK_TSV_NOT_INT16 2   /* a field is not a plain integer in [-32768, 32767] (pA output ...) */
#define SQK_TSV_ALL_ZERO 4    /* every sample is 0: the reference prints "No Signal found"           */
SQK_API int sqk_tsv_parse(const char *text, int64_t n_bytes, int is_final, int start_col, int64_t max_lines,
                          int64_t max_samples, int n_threads, int16_t *samples, int64_t *offsets, int64_t *line_begin,
                          int64_t *sig_begin, int32_t *status, int64_t *n_lines, int64_t *consumed);
/* first n_cols columns of every parsed line, "c0 <TAB> c1 <NL>" each, gathered into out (bytes written, or minus the bytes
 * needed when cap is too small / out is NULL) */
SQK_API int64_t sqk_tsv_heads(const char *text, const int64_t *line_begin, const int64_t *sig_begin, int64_t n_lines,
                              int n_cols, char *out, int64_t cap);
/* The rows get_region_multi prints (MotifSeq.py:441-449) for a batch, floats written as Python's repr() writes them:
 * heads = "fast5 <TAB> readID <NL>" per read (sqk_tsv_heads); names / consts = per model, NUL-separated (consts: the text of
 * "mod_mean <TAB> mod_stdev"); hits [n_reads][n_models] sqk_hit; zs / ps / hps [n_reads][n_models] = Z-score, p-value,
 * hit probability.  Reads whose hit is a status (start < 0) print nothing.  Returns the bytes written, or minus an upper
 * bound of the bytes needed when cap is too small (out may be NULL to ask). */
SQK_API int64_t sqk_tsv_format_rows(const char *heads, int64_t n_reads, const void *hits, int n_models, const char *names,
                                    const char *consts, const double *zs, const double *ps, const double *hps, int n_threads,
                                    char *out, int64_t cap);
/* The rows segmenter.py prints (segmenter.py:130-146: name <TAB> s0,e0,s1,e1,...) for the reads with keep[r] != 0; heads =
 * "name <NL>" per read (sqk_tsv_heads, one column); segs [n_reads][max_segs][2], n_segs [n_reads] as sqk_segmenter returns
 * them.  Returns the bytes written, or minus an upper bound of the bytes needed when cap is too small / out is NULL. */
SQK_API int64_t sqk_tsv_format_segs(const char *heads, int64_t n_reads, const int32_t *segs, const int32_t *n_segs, int max_segs,
                                    const unsigned char *keep, int n_threads, char *out, int64_t cap);
SQK_API int64_t sqk_tsv_format(const int16_t *samples, const int64_t *offsets, int64_t n_reads, const char *heads,
                               const int64_t *head_offsets, int n_threads, char *out, int64_t cap);
/* The score columns of those rows (MotifSeq.py:441-445) on the host, without scipy: Z = (dist - mod_mean) / mod_stdev,
 * p = scipy.stats.norm.cdf(Z) bit for bit (its ndtr: the Cephes algorithm and tables, libm's exp), hit_P = (1 - p) * 100, for
 * hits [n_reads][n_models]; mod_mean / mod_stdev per model.  sqk_ndtr is the cdf alone. */
SQK_API void sqk_score_hits(const void *hits, int64_t n_reads, int n_models, const double *mod_mean, const double *mod_stdev,
                            int n_threads, double *zs, double *ps, double *hps);
SQK_API void sqk_ndtr(const double *z, int64_t n, double *out);

/* ---------------------------------------------------------------------------------------
 * Instrumentation (bench.py): per-kernel device time measured with cudaEvents recorded on the
 * launching stream around each launch.  Off by default.  Reading the counters synchronises.
 * ------------------------------------------------------------------------------------- */
enum sqk_kernel_id {
    SQK_K_STATS = 0,
    SQK_K_DTW = 1,      /* single-pass float64 / float32 DTW kernel */
    SQK_K_SEG_FSM = 2,
    SQK_K_DTW_LB = 3,   /* two-pass plan, pass 1: float32 lower-bound scan of every read */
    SQK_K_DTW_WIN = 4,  /* two-pass plan, pass 2: exact float64 windows + finalize + full-length fallback */
    SQK_K_COUNT = 5
};
typedef struct {
    int64_t launches[SQK_K_COUNT];
    double ms[SQK_K_COUNT];
} sqk_timing;
SQK_API int sqk_ctx_enable_timing(sqk_ctx *ctx, int on);
SQK_API int sqk_ctx_get_timing(sqk_ctx *ctx, sqk_timing *out, int reset);

/* Tuning knobs.  dtw_lanes: force lanes-per-read of the DTW kernel (0 = automatic).  chunk_samples: host mode,
 * samples per in-flight chunk of the H2D | compute | D2H pipeline (0 = default 64 Mi; the first two chunks
 * are 1/4 and 1/2 of it). */
SQK_API int sqk_ctx_set_chunk_samples(sqk_ctx *ctx, int64_t samples);
SQK_API int sqk_ctx_set_dtw_lanes(sqk_ctx *ctx, int lanes);

/* How SQK_PREC_FP64 requests of sqk_motifseq are executed -- the results are identical bit for bit:
 *   SINGLE_PASS  the float64 recurrence with start pointers over every column of every read;
 *   TWO_PASS     a float32, rounded-down, cost-only scan proves which columns can hold mlpy's
 *                np.argmin(cost[-1, :]) (MotifSeq.py:437-439); the float64 recurrence then runs only on
 *                windows around those columns, and on the whole read whenever the proof does not close;
 *   AUTO         TWO_PASS when the longest read is at least 4 windows long (default). */
enum sqk_dtw_plan { SQK_PLAN_AUTO = 0, SQK_PLAN_SINGLE_PASS = 1, SQK_PLAN_TWO_PASS = 2 };
SQK_API int sqk_ctx_set_dtw_plan(sqk_ctx *ctx, int plan);
/* Diagnostics of the most recent two-pass launch of slot 0 (device-mode calls; the last chunk in host mode), first
 * model: out[0] = exact windows run, out[1] = reads re-run over their full length.  Synchronises. */
SQK_API int sqk_ctx_get_plan_counters(sqk_ctx *ctx, int64_t out[2]);
/* the same plus out[2] = second-attempt windows run (reads whose first windows were too short), out[3] reserved */
SQK_API int sqk_ctx_get_plan_counters_ex(sqk_ctx *ctx, int64_t out[4]);

/* Kernels launched by this ctx since the last reset (every launch is counted where it is made). */
SQK_API int sqk_ctx_get_launches(sqk_ctx *ctx, int64_t *out, int reset);

/* Which statistics kernel (K1) runs: 0 = automatic (second generation -- bulk-copy staging, no compaction pass -- for
 * reads of up to 8176 samples, the first generation beyond that and for the reads the second hands back), 1 = first
 * generation only.  Results are identical bit for bit; the knob exists for A/B measurements and tests. */
SQK_API int sqk_ctx_set_stats_generation(sqk_ctx *ctx, int generation);

/* ---------------------------------------------------------------------------------------
 * Multi-GPU: one process per GPU, reads sharded over the ranks.  The only exchange on the path is the
 * gather of the 16-byte hit records.  Instead of a collective after the kernels, the kernels that PRODUCE a record store
 * it into the gathered buffer of every peer GPU through P2P-mapped pointers (NVLink): a fused compute + all-gather.
 *
 *   every rank:   sqk_device_alloc(gathered buffer [world * n_local * n_models] sqk_hit, flag array)
 *                 sqk_ipc_export -> exchange the 64-byte handles by any means (torch.distributed, MPI, a file)
 *                 sqk_ipc_open on every peer's handles
 *                 sqk_ctx_set_hit_peers(peers' buffer bases, n, first_record = rank * n_local)
 *                 sqk_ctx_set_flag_peers(all ranks' flag arrays incl. the own one, world, rank)
 *   per step:     sqk_motifseq(..., SQK_MEM_DEVICE, hits = own block of the own gathered buffer, ...)
 *                 sqk_peer_signal(step)       -- after the step's kernels, stream-ordered
 *   to read:      sqk_peer_wait(step)         -- stream-ordered: returns (on the stream) once every rank signalled `step`
 *
 * Only device-mode sqk_motifseq publishes, and sqk_ctx_set_hit_peers arms ONE call: the next device-mode sqk_motifseq
 * stores its records into the peers' buffers, calls after it do not until the peers are set again (a call that is not
 * part of the exchange -- a comparison run, another batch -- must never write into a peer's memory).  A buffer may be
 * rewritten by a later step as soon as that step runs: the caller rotates buffers (sqk_ctx_set_hit_peers per step) and
 * reads a buffer before the ranks reuse it.
 * ------------------------------------------------------------------------------------- */
SQK_API int sqk_device_alloc(sqk_ctx *ctx, uint64_t bytes, void **dev_ptr);     /* zero-filled; IPC-exportable */
SQK_API int sqk_device_free(sqk_ctx *ctx, void *dev_ptr);
SQK_API int sqk_ipc_export(sqk_ctx *ctx, void *dev_ptr, unsigned char handle[64]);
SQK_API int sqk_ipc_open(sqk_ctx *ctx, const unsigned char handle[64], void **dev_ptr);
SQK_API int sqk_ipc_close(sqk_ctx *ctx, void *dev_ptr);
/* peers[n_peers]: bases of the OTHER ranks' gathered buffers (device pointers valid on this GPU); record r of this rank
 * is stored at peers[p][(first_record + r) * n_models + m].  n_peers = 0 turns publication off. */
SQK_API int sqk_ctx_set_hit_peers(sqk_ctx *ctx, void *const *peers, int n_peers, int64_t first_record);
/* the same with the size of a gathered buffer in records: a call whose block [first_record, first_record + n_reads) does not
 * fit is refused (SQK_ERR_ARG) instead of storing past the end of a peer's buffer */
SQK_API int sqk_ctx_set_hit_peers_ex(sqk_ctx *ctx, void *const *peers, int n_peers, int64_t first_record,
                                     int64_t capacity_records);
/* flag_arrays[n_ranks]: every rank's flag array (uint64[16], zero-filled), the own one at index my_rank. */
SQK_API int sqk_ctx_set_flag_peers(sqk_ctx *ctx, void *const *flag_arrays, int n_ranks, int my_rank);
SQK_API int sqk_peer_signal(sqk_ctx *ctx, uint64_t value);
SQK_API int sqk_peer_wait(sqk_ctx *ctx, uint64_t value);

#ifdef __cplusplus
}
#endif
#endif /* SQK_H */
