// sqk_f64.cuh -- float64-signal front end: the reference's `-s` path feeds both tools
// whatever numbers the TSV holds (`float(i) for i in l[8:]`, MotifSeq.py:270; the float branch of
// segmenter.py:198-199), e.g. SquigglePull's pA output.  Arbitrary doubles have no int16 form, so this
// path runs ONE extra kernel per batch that does, per read (one CTA):
//     A  scale_outliers                  keep lo < v < hi, compact the survivors into a global scratch row
//     B  statistics, bit-identical to numpy on float64: pairwise mean and sigma (stats_sum), medians by an
//        8-pass radix select on order-preserving 64-bit keys (exact k-th order statistics)
//     C  MotifSeq: normalise the row in place, (v - center) / scale  -> the DTW kernel streams it as is
//        segmenter: write a 0/1 int16 code per kept sample (bot < v < top) -> the int16 FSM kernel runs on
//        the code row with window [1,1]
// so K2 and K3 are reused unchanged apart from where they fetch a read from.  This path is about coverage,
// not speed: it moves 8+8(+2) bytes per sample through HBM instead of 2.
#pragma once
#include "sqk_stats.cuh"

struct F64Args {
    const double *base;       // base[i] = absolute sample i
    const int64_t *offsets;   // absolute
    int64_t read0, n_reads;
    ReadStats *stats;         // [n_reads]
    int32_t *n_kept_out;      // [n_reads] or null
    int mode, lo, hi, num;    // SQK_STATS_* ; num: segmenter truncation
    double std_scale;
    double *ynorm;            // ynorm[i]: scratch row of the read that starts at absolute sample i (compacted)
    int16_t *codes;           // segmenter: 0/1 code per kept sample, same indexing (or null)
};

__device__ __forceinline__ unsigned long long f64_key(double v)
{
    const long long b = __double_as_longlong(v);
    return (unsigned long long)b ^ ((unsigned long long)(b >> 63) | 0x8000000000000000ull);
}

__device__ __forceinline__ double f64_from_key(unsigned long long k)
{
    const unsigned long long b = (k & 0x8000000000000000ull) ? (k ^ 0x8000000000000000ull) : ~k;
    return __longlong_as_double((long long)b);
}

// rank-th smallest (0-based) of val(i), i < n: 8 radix passes over order-preserving keys, most significant byte
// first; every thread returns the value.
template <class ValFn>
__device__ double f64_select(ValFn val, int n, int rank, StatsShared &sh)
{
    const int tid = threadIdx.x;
    unsigned long long prefix = 0;
#pragma unroll 1
    for (int pass = 0; pass < 8; pass++) {
        const int shift = 56 - 8 * pass;
        for (int b = tid; b < 256; b += SQK_STATS_THREADS) sh.hist[b] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += SQK_STATS_THREADS) {
            const unsigned long long k = f64_key(val(i));
            if (pass == 0 || (k >> (shift + 8)) == prefix) atomicAdd(&sh.hist[(unsigned)(k >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid < 32) {
            uint32_t mine = 0;
            for (int b = 0; b < 8; b++) mine += sh.hist[tid * 8 + b];
            uint32_t incl = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(SQK_FULL_MASK, incl, d);
                if (tid >= d) incl += t;
            }
            uint32_t cum = incl - mine;
            if ((uint32_t)rank >= cum && (uint32_t)rank < incl) {
                for (int b = 0; b < 8; b++) {
                    const uint32_t h = sh.hist[tid * 8 + b];
                    if ((uint32_t)rank < cum + h) { sh.sel[0] = tid * 8 + b; sh.sel[1] = rank - cum; break; }
                    cum += h;
                }
            }
        }
        __syncthreads();
        prefix = (prefix << 8) | sh.sel[0];
        rank = (int)sh.sel[1];
        __syncthreads();
    }
    return f64_from_key(prefix);
}

// np.median of val(0..n-1): middle element, or the mean of the two middle ones ((a + b) / 2, as np.mean does)
template <class ValFn>
__device__ double f64_median(ValFn val, int n, StatsShared &sh)
{
    if (n & 1) return f64_select(val, n, (n - 1) / 2, sh);
    const double a = f64_select(val, n, n / 2 - 1, sh);
    const double b = f64_select(val, n, n / 2, sh);
    return __ddiv_rn(__dadd_rn(a, b), 2.0);
}

__global__ void __launch_bounds__(SQK_STATS_THREADS) sqk_f64_front_kernel(const F64Args a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    StatsShared &sh = *reinterpret_cast<StatsShared *>(smem_raw);
    __shared__ int s_total;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int WARPS = SQK_STATS_THREADS / 32;
    const double lo = (double)a.lo, hi = (double)a.hi;

    for (int64_t i = blockIdx.x; i < a.n_reads; i += gridDim.x) {
        const int64_t r = a.read0 + i;
        const int64_t begin = a.offsets[r];
        int64_t len = a.offsets[r + 1] - begin;
        if (a.mode == SQK_STATS_SEGMENTER) len = sqk_truncate_len(len, a.num);
        double *row = a.ynorm + begin;

        // ---- A: outlier removal + compaction (order preserved) ---------------------------------
        if (tid == 0) s_total = 0;
        __syncthreads();
        for (int64_t b0 = 0; b0 < len; b0 += SQK_STATS_THREADS) {
            const int64_t q = b0 + tid;
            const double v = q < len ? a.base[begin + q] : 0.0;
            const bool keep = q < len && v > lo && v < hi;
            const unsigned bal = __ballot_sync(SQK_FULL_MASK, keep);
            if (lane == 0) sh.warp_tot[warp] = __popc(bal);
            __syncthreads();
            int pos = s_total + __popc(bal & ((1u << lane) - 1u)), all = 0;
#pragma unroll
            for (int w = 0; w < WARPS; w++) { if (w < warp) pos += sh.warp_tot[w]; all += sh.warp_tot[w]; }
            if (keep) row[pos] = v;
            __syncthreads();
            if (tid == 0) s_total += all;
            __syncthreads();
        }
        const int n = s_total;
        __threadfence_block();
        __syncthreads();

        ReadStats out;
        out.center = 0.0; out.scale = 1.0; out.n_kept = n; out.flags = 0;
        out.seg_lo = 1; out.seg_hi = 1; out.out_lo = 0; out.out_hi = 1;     // window of the 0/1 code row
        auto val = [row](int q) -> double { return row[q]; };

        // ---- B: statistics ----------------------------------------------------------------------
        double sd = 0.0;
        if (n > 0 && (a.mode == SQK_STATS_ZSCALE || a.mode == SQK_STATS_SEGMENTER)) {
            const double mean = __ddiv_rn(stats_sum<SQK_STATS_THREADS>(val, n, sh), (double)n);
            auto sq = [row, mean](int q) -> double { const double d = __dsub_rn(row[q], mean); return __dmul_rn(d, d); };
            sd = __dsqrt_rn(__ddiv_rn(stats_sum<SQK_STATS_THREADS>(sq, n, sh), (double)n));
            if (a.mode == SQK_STATS_ZSCALE) {
                if (sd == 0.0) sd = 1.0;
                out.center = mean; out.scale = sd;
            }
        }
        double top = 0.0, bot = 0.0;
        if (n > 0 && (a.mode == SQK_STATS_MEDMAD || a.mode == SQK_STATS_SEGMENTER)) {
            const double med = f64_median(val, n, sh);
            if (a.mode == SQK_STATS_MEDMAD) {
                auto dev = [row, med](int q) -> double { return fabs(__dsub_rn(row[q], med)); };
                const double mad = f64_median(dev, n, sh);
                const double scaled = __dmul_rn(mad, 1.4826);
                out.center = med; out.scale = scaled;
                if (!(scaled > 0.0) && !(scaled < 0.0)) out.flags |= SQK_FLAG_DEGENERATE;   // 0 or NaN
            } else {
                const double spread = __dmul_rn(sd, a.std_scale);
                top = __dadd_rn(med, spread);
                bot = __dsub_rn(med, spread);
                out.center = top; out.scale = bot;
            }
        }

        // ---- C: hand the read to K2 / K3 --------------------------------------------------------
        if (a.mode == SQK_STATS_SEGMENTER) {
            int16_t *code = a.codes + begin;
            for (int q = tid; q < n; q += SQK_STATS_THREADS) {
                const double v = row[q];
                code[q] = (v < top && v > bot) ? (int16_t)1 : (int16_t)0;
            }
        } else if (!(out.flags & SQK_FLAG_DEGENERATE)) {
            const double c = out.center, s = out.scale;
            for (int q = tid; q < n; q += SQK_STATS_THREADS) row[q] = __ddiv_rn(__dsub_rn(row[q], c), s);
        }
        if (tid == 0) {
            a.stats[i] = out;
            if (a.n_kept_out) a.n_kept_out[i] = n;
        }
        __syncthreads();
    }
}
