// sqk_dtw.cuh -- K2: subsequence DTW of one motif against many reads (mlpy.dtw_subsequence as
// called at MotifSeq.py:437, + start/end extraction MotifSeq.py:438-439), fused with the
// int16 -> float64 outlier compaction and normalisation in front of it (MotifSeq.py:186-200,
// 317-324).  No tensor cores: the recurrence has no contraction; it is bound by the FP64/ALU
// issue rate, not by HBM.
//
// Mapping.  A read is owned by a group of L lanes (32/L reads per warp); lane l keeps K
// consecutive motif rows of the current DTW column in registers: cost (fp64) and the start
// pointer (column where the best path into that cell left row 0).  Lanes run a skewed
// wavefront: at step t lane l handles signal column t-l; the bottom cell of lane l-1 reaches lane
// l by one __shfl_up per step.  Only one column ever exists -> any read length, no N x M matrix.
// The motif rows are spread so that the last row always sits in register K-1 of lane L-1: when
// L*K > N the last (L*K-N) lanes carry a pass-through slot in register 0.
//
// Exactness (FP64 mode).  Every cost is  fl(|x_i - y_j| + min3)  with the same operand order as
// mlpy's C loop; min3 and the pointer follow the back-trace's preference: diagonal if it ties
// the minimum, else left, else up.  Forward pointers equal the back-trace by induction because
// each choice is a function of the same three neighbours.  The last row keeps a running
// first-argmin (np.argmin).  Row 0 (free start) is produced by feeding lane 0 a virtual row of
// cost 0 whose pointer is j+1: min3 = 0 via the diagonal, pointer j, cost |x_0 - y_j| + 0.
//
// Streaming.  Raw int16 is read from HBM exactly once, 16 bytes per lane, filtered lo < s < hi,
// normalised ((s - center) / scale, both roundings as numpy) and written to a per-group ring in
// shared memory holding 16*L doubles; the DTW consumes one ring entry per step.  Groups pull
// reads from a global atomic counter and re-arm independently, so ragged read lengths need no
// sorting and the tail is one read per group.
//
// JOBS variant (pass 2 of the exact two-pass plan, sqk_dtw_plan.cuh).  The work items are DtwJob records --
// a column window of a read, or a whole read -- instead of reads.  A tainted job treats its first column as a
// boundary: after computing it, rows >= 1 are overwritten with cost -1 and the pointer SQK_TAINT, a strict lower
// bound of whatever the columns in front of the window would have delivered.  The last-row argmin only looks at
// columns >= arg_lo (the candidate cluster); if the minimum carries the taint the job reports start = SQK_TAINT.
#pragma once
#include "sqk_common.cuh"
#include "sqk_dtw_plan.cuh"

#define SQK_DTW_WARPS 4
#define SQK_DTW_THREADS (SQK_DTW_WARPS * 32)
// resident CTAs per SM the register allocator must allow for (experiments: -DSQK_DTW_MINB_SMALLK=5)
#ifndef SQK_DTW_MINB_SMALLK
#define SQK_DTW_MINB_SMALLK 1
#endif
#define SQK_DTW_MINB(K) ((K) <= 10 ? SQK_DTW_MINB_SMALLK : 1)

// one cell of a boundary row between two row blocks: cost and start pointer as the recurrence carries them
struct __align__(16) BndCell { double c; int s; int pad; };

struct DtwArgs {
    const int16_t *base;      // base[i] = absolute sample i
    int64_t alloc_lo, alloc_hi;
    const int64_t *offsets;   // absolute
    int64_t read0;
    int n_reads;
    const ReadStats *stats;   // [n_reads]
    const double *model;      // N motif points
    int N;
    int lo, hi;
    sqk_hit *hits;            // hits[i * hit_stride]
    int hit_stride;
    unsigned int *counter;    // work queue head, zeroed before launch
    const double *prenorm;    // float64 front end: prenorm[offsets[r] + i] = i-th normalised kept sample (else null)
    // JOBS variant only
    const DtwJob *jobs;       // work items
    const unsigned int *n_jobs;   // how many (device memory: produced by the previous kernel)
    sqk_hit *job_out;         // job_out[job.out * job_out_stride]
    int job_out_stride;
    // BND variant only (motifs longer than one pass holds: the rows are cut into blocks, sqk_api.cu): row r0 - 1 of the
    // previous block for every column of every read comes in, the block's last row goes out
    const BndCell *bnd_in;    // [n_reads][bnd_stride] or null: this is the first block (free-start row above)
    BndCell *bnd_out;         // [n_reads][bnd_stride] or null: this is the last block (the hit is written)
    int64_t bnd_stride;
};

template <typename T> struct DtwNum;
template <> struct DtwNum<double> {
    static __device__ __forceinline__ double inf() { return SQK_INF_D; }
    static __device__ __forceinline__ double shfl_up(double v, int w) { return shfl_up_f64(v, 1, w); }
    // |x - y| + m, two roundings (no contraction possible: no multiply)
    static __device__ __forceinline__ double step(double x, double y, double m) { return __dadd_rn(fabs(__dsub_rn(x, y)), m); }
};
template <> struct DtwNum<float> {
    static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
    static __device__ __forceinline__ float shfl_up(float v, int w) { return __shfl_up_sync(SQK_FULL_MASK, v, 1, w); }
    static __device__ __forceinline__ float step(float x, float y, float m) { return __fadd_rn(fabsf(__fsub_rn(x, y)), m); }
};

// One wavefront step of one lane: column (t - l) for this lane's K rows.  (ci, si) hold the previous
// column of these rows, (co, so) receive the new one.
template <typename T, int K, int L, bool RAGGED, bool JOBS, bool BND = false>
__device__ __forceinline__ void dtw_step(const T (&ci)[K], const int (&si)[K], T (&co)[K], int (&so)[K],
                                         const T (&x)[K], const T *ring, int l, bool pass0, int t, int n_last,
                                         T &bot_c, int &bot_s, T &prev_up_c, int &prev_up_s,
                                         T &best, int &best_j, int &best_s, int arg_lo, bool tainted,
                                         const BndCell *bin = nullptr, BndCell *bout = nullptr, int n_cols = 0,
                                         BndCell *pre = nullptr)
{
    using Num = DtwNum<T>;
    constexpr int RC = 16 * L;
    T up_c = Num::shfl_up(bot_c, L);
    int up_s = __shfl_up_sync(SQK_FULL_MASK, bot_s, 1, L);
    if (l == 0) { up_c = (T)0; up_s = t + 1; }          // virtual row above row 0: free start
    if constexpr (BND) {
        if (l == 0 && bin != nullptr) {
            // the row above this block's first row: column t was fetched during the previous step, t + 1 is requested now
            up_c = (T)pre->c; up_s = pre->s;
            if (t + 1 < n_cols) *pre = bin[t + 1];
        }
    }
    const T y = ring[(t - l) & (RC - 1)];
    T dg_c = prev_up_c; int dg_s = prev_up_s;
    prev_up_c = up_c; prev_up_s = up_s;
    T u_c = up_c; int u_s = up_s;
#pragma unroll
    for (int k = 0; k < K; k++) {
        const T lf_c = ci[k]; const int lf_s = si[k];
        const bool p = lf_c < dg_c;                 // left beats diagonal only if strictly smaller
        const T m1_c = p ? lf_c : dg_c; const int m1_s = p ? lf_s : dg_s;
        const bool q = u_c < m1_c;                  // up only if strictly smaller than both
        const T m_c = q ? u_c : m1_c; int m_s = q ? u_s : m1_s;
        T nc = Num::step(x[k], y, m_c);
        if (RAGGED && k == 0 && pass0) { nc = up_c; m_s = up_s; }   // only when L*K != N
        dg_c = lf_c; dg_s = lf_s;
        u_c = nc; u_s = m_s;
        co[k] = nc; so[k] = m_s;
    }
    bot_c = u_c; bot_s = u_s;
    if constexpr (BND) {
        if (bout != nullptr && l == L - 1 && (unsigned)(t - (L - 1)) < (unsigned)n_cols) {
            BndCell o; o.c = (double)bot_c; o.s = bot_s; o.pad = 0;
            bout[t - (L - 1)] = o;
        }
    }
    if constexpr (JOBS) {
        if (tainted && t == l) {
            // this lane just computed the window's boundary column: rows >= 1 become the strict lower bound -1 with
            // the taint pointer (row 0 keeps its exact value |x_0 - y|: it has no left neighbour in mlpy's recurrence)
#pragma unroll
            for (int k = 0; k < K; k++)
                if (k > 0 || l > 0) { co[k] = (T)-1; so[k] = SQK_TAINT; }
            if (K > 1 || l > 0) { bot_c = (T)-1; bot_s = SQK_TAINT; }
        }
    }
    // running first-argmin of the last row: n_last is the read length in lane L-1 and 0 elsewhere, so only
    // the lane that owns the last motif row can fire.  New minima are rare (O(log M) per read), so the update
    // sits behind a warp vote instead of costing four selects every step.
    const int j = t - (L - 1);
    // n_last counts the columns from arg_lo on (JOBS) / from 0 on
    const bool better = (unsigned)(JOBS ? j - arg_lo : j) < (unsigned)n_last && bot_c < best;
    if (__any_sync(SQK_FULL_MASK, better)) {
        if (better) { best = bot_c; best_j = j; best_s = bot_s; }
    }
}

template <typename T, int K, int L, bool RAGGED, bool JOBS = false, bool BND = false>
__global__ void __launch_bounds__(SQK_DTW_THREADS, SQK_DTW_MINB(K)) sqk_dtw_kernel(const DtwArgs a)
{
    constexpr int G = 32 / L;          // reads per warp
    constexpr int RC = 16 * L;         // ring capacity (entries), power of two
    constexpr int S = (L == 1) ? 8 : 7 * L;   // DTW steps between ring refills: even, S <= RC - 9L + 2
    constexpr int CH = 8 * L;          // raw samples fetched per refill
    using Num = DtwNum<T>;

    __shared__ T ring_all[SQK_DTW_WARPS * G * RC];

    int64_t alloc_lo = a.alloc_lo, alloc_hi = a.alloc_hi;
    resolve_bounds(a.offsets, a.read0, a.n_reads, alloc_lo, alloc_hi);
    const unsigned n_items = JOBS ? *a.n_jobs : (unsigned)a.n_reads;

    const int lane = threadIdx.x & 31;
    const int l = lane % L;            // lane within the group
    const int g = lane / L;            // group within the warp
    T *ring = ring_all + ((threadIdx.x >> 5) * G + g) * RC;
    for (int q = l; q < RC; q += L) ring[q] = (T)0;   // never leave non-finite garbage in the ring
    __syncwarp();
    const BndCell *bin = nullptr;      // BND: this read's boundary rows
    BndCell *bout = nullptr;
    BndCell pre; pre.c = 0.0; pre.s = 0; pre.pad = 0;

    // motif rows of this lane
    const int P = L * K - a.N;                         // pass-through slots, 0 <= P < L
    const bool pass0 = (l >= L - P);                   // register 0 is a pass-through slot
    const int row0 = pass0 ? (L - P) * K + (l - (L - P)) * (K - 1) - 1 : l * K;   // motif row of register 0
    T x[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        const int row = row0 + k;
        x[k] = (row >= 0 && row < a.N) ? (T)a.model[row] : (T)0;
    }

    T c[K], c2[K];
    int s[K], s2[K];
    T bot_c = Num::inf(), prev_up_c = Num::inf(), best = Num::inf();
    int bot_s = 0, prev_up_s = 0, best_j = -1, best_s = -1;
    int n = 0, n_last = 0, t = 0, wcount = 0, my_read = -1;
    int col0 = 0, arg_lo = 0;          // JOBS: first column of the window, first column the argmin looks at
    bool tainted = false;
    int64_t begin = 0, end = 0, cursor = 0;
    double center = 0.0, scale = 1.0;
    bool done = true, exhausted = false;
#pragma unroll
    for (int k = 0; k < K; k++) { c[k] = Num::inf(); s[k] = 0; }

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
                if (idx >= n_items) {
                    exhausted = true;
                } else if constexpr (JOBS) {
                    const DtwJob jb = a.jobs[idx];
                    my_read = jb.out;
                    const int64_t r = a.read0 + jb.read;
                    begin = a.offsets[r];
                    end = a.offsets[r + 1];
                    const ReadStats st = a.stats[jb.read];
                    center = st.center; scale = st.scale;
                    n = jb.n_cols; col0 = jb.col0; arg_lo = jb.arg_lo; tainted = jb.tainted != 0;
                    if (n > arg_lo && arg_lo >= 0) {       // always true for jobs built by pass 1 / finalize
                        done = false;
                        n_last = (l == L - 1) ? n - arg_lo : 0;
                        t = 0; wcount = 0;
                        cursor = jb.cursor;
#pragma unroll
                        for (int k = 0; k < K; k++) { c[k] = Num::inf(); s[k] = 0; }
                        bot_c = Num::inf(); bot_s = 0;
                        prev_up_c = (l == 0) ? (T)0 : Num::inf(); prev_up_s = 0;
                        best = Num::inf(); best_j = -1; best_s = -1;
                    }
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
                    if (n <= 0) {
                        // empty after outlier removal (-1) / undefined scale (-2) / longer than declared (-3): the
                        // reference skips or prints NaN; report a status instead and pull again next round
                        if (l == L - 1) {
                            sqk_hit h; h.start = n == 0 ? -1 : (n == -1 ? -2 : -3); h.end = h.start; h.dist = __longlong_as_double(0x7ff8000000000000LL);
                            a.hits[(int64_t)idx * a.hit_stride] = h;
                        }
                    } else {
                        done = false;
                        n_last = (l == L - 1) ? n : 0;
                        t = 0; wcount = 0;
                        cursor = aligned_block_start(a.base, begin);
#pragma unroll
                        for (int k = 0; k < K; k++) { c[k] = Num::inf(); s[k] = 0; }
                        bot_c = Num::inf(); bot_s = 0;
                        prev_up_c = (l == 0) ? (T)0 : Num::inf(); prev_up_s = 0;
                        if constexpr (BND) {
                            bin = a.bnd_in ? a.bnd_in + (int64_t)idx * a.bnd_stride : nullptr;
                            bout = a.bnd_out ? a.bnd_out + (int64_t)idx * a.bnd_stride : nullptr;
                            if (bin != nullptr) {
                                prev_up_c = Num::inf();                    // no column in front of the first one
                                if (l == 0) pre = bin[0];
                            }
                            if (bout != nullptr) n_last = 0;               // not the last block: no argmin, no hit
                        }
                        best = Num::inf(); best_j = -1; best_s = -1;
                    }
                }
            }
        }
        if (__all_sync(SQK_FULL_MASK, exhausted)) break;   // exhausted implies done; others re-pull above

        // ---- refill the rings: raw int16 -> filter -> normalise -> shared memory ---------------
        if (!JOBS && a.prenorm != nullptr) {
            // float64 front end (sqk_f64.cuh): the read is already compacted and normalised in a global row
            for (;;) {
                const bool want = !done && (wcount < t + S) && (wcount < n);
                if (!__any_sync(SQK_FULL_MASK, want)) break;
                if (want) {
                    const double *row = a.prenorm + begin;
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        const int idx = wcount + l * 8 + e;
                        if (idx < n) ring[idx & (RC - 1)] = (T)row[idx];
                    }
                    wcount = (n - wcount < CH) ? n : wcount + CH;
                }
            }
        } else
        for (;;) {
            const bool want = !done && (wcount < t + S) && (cursor < end);
            if (!__any_sync(SQK_FULL_MASK, want)) break;
            unsigned keep = 0;
            Samples8 smp;
            const int64_t blk = cursor + l * 8;
            if (want && blk < end) {
                smp = load_block8(a.base, blk, alloc_lo, alloc_hi);
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    const int v = smp.get(e);
                    const int64_t idx = blk + e;
                    if (idx >= begin && idx < end && v > a.lo && v < a.hi) keep |= 1u << e;
                }
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
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    if (keep & (1u << e)) {
                        const double y = __ddiv_rn(__dsub_rn((double)smp.get(e), center), scale);
                        ring[pos & (RC - 1)] = (T)y;
                        pos++;
                    }
                }
            }
            if (want) { wcount += group_total; cursor += CH; }
        }
        __syncwarp();

        // ---- S wavefront steps (two per trip: the column ping-pongs between two register sets,
        //      so no register-to-register copies are needed to keep the previous column alive) ----
#pragma unroll 1
        for (int it = 0; it < S; it += 2) {
            dtw_step<T, K, L, RAGGED, JOBS, BND>(c, s, c2, s2, x, ring, l, pass0, t, n_last, bot_c, bot_s, prev_up_c, prev_up_s, best, best_j, best_s, arg_lo, tainted, bin, bout, n, &pre);
            t++;
            dtw_step<T, K, L, RAGGED, JOBS, BND>(c2, s2, c, s, x, ring, l, pass0, t, n_last, bot_c, bot_s, prev_up_c, prev_up_s, best, best_j, best_s, arg_lo, tainted, bin, bout, n, &pre);
            t++;
        }
        __syncwarp();

        if (!done && t >= n + L - 1) {
            if (l == L - 1) {
                sqk_hit h; h.start = best_s; h.end = best_j; h.dist = (double)best;
                if constexpr (JOBS) {
                    h.start = best_s == SQK_TAINT ? SQK_TAINT : col0 + best_s;   // window columns -> read columns
                    h.end = col0 + best_j;
                    a.job_out[(int64_t)my_read * a.job_out_stride] = h;
                } else {
                    if (!BND || bout == nullptr) a.hits[(int64_t)my_read * a.hit_stride] = h;
                }
            }
            done = true;
        }
    }
}
