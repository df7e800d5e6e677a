// sqk_common.cuh -- shared device-side types and helpers for the libsqk kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sqk.h"

#define SQK_FULL_MASK 0xffffffffu
#define SQK_INF_D __longlong_as_double(0x7ff0000000000000LL)

// Per-read summary written by the stats kernel (K1) and consumed by the DTW / FSM kernels.
//   MotifSeq:  y = (x - center) / scale          (zscale: mean, std | medmad: med, mad*1.4826)
//   segmenter: in-range  <=>  seg_lo <= x <= seg_hi  (integer form of  bot < x < top)
struct __align__(8) ReadStats {
    double center;
    double scale;
    int32_t n_kept;   // samples surviving  lo < s < hi  (after truncation for segmenter)
    int32_t flags;    // bit0: degenerate scale (MAD == 0)
    int32_t seg_lo;
    int32_t seg_hi;
    int32_t out_lo;   // outlier window on the raw sample, inclusive: kept  <=>  out_lo <= s <= out_hi
    int32_t out_hi;   //   (raw mode: lo+1 .. hi-1; pA mode: the raw values whose pA lies in (lim_low, lim_hi))
};
static_assert(sizeof(ReadStats) == 40, "ReadStats layout");

#define SQK_FLAG_DEGENERATE 1
#define SQK_FLAG_TOO_LONG 2     // read longer than the max_read_len the caller declared: not processed
#define SQK_FLAG_NO_MASK 4      // segmenter: no in-range bit mask was emitted for this read (sqk_fsm_kernel takes it)

// convert_to_pA_numpy + np.round(.., 2)  (segmenter.py:515-517, 345-349): every op rounds once, as numpy's
//   (d + offset) * raw_unit ; multiply by 100 ; rint ; divide by 100.
// Monotone non-decreasing in d for raw_unit > 0, so windows on pA are windows on d.
__device__ __forceinline__ double sqk_pa_value(int d, double offset, double raw_unit)
{
    const double x = __dmul_rn(__dadd_rn((double)d, offset), raw_unit);
    return __ddiv_rn(rint(__dmul_rn(x, 100.0)), 100.0);
}

// python  sig[:Num]  with Num = num ? num : -1   (segmenter.py:104-105,207)
__host__ __device__ __forceinline__ int64_t sqk_truncate_len(int64_t len, int num)
{
    int64_t use;
    if (num == 0) use = len - 1;
    else if (num > 0) use = num < len ? num : len;
    else use = len + num;
    return use < 0 ? 0 : use;
}

// 8 consecutive int16 samples fetched as one 16-byte word.
struct Samples8 {
    int4 v;
    __device__ __forceinline__ int get(int e) const {
        const int w = (e < 2) ? v.x : (e < 4) ? v.y : (e < 6) ? v.z : v.w;
        return (e & 1) ? (w >> 16) : (int)(short)(w & 0xffff);
    }
};

// Streaming 16-byte load (read once, do not pollute L1).
__device__ __forceinline__ int4 ld_stream_16(const void *p)
{
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// Cached 16-byte load (the line stays in L1: for kernels whose lanes each walk their own read, so that the
// other half of every 32-byte sector -- and the rest of the 128-byte line -- is not fetched from L2/HBM again).
__device__ __forceinline__ int4 ld_cached_16(const void *p)
{
    return __ldg(reinterpret_cast<const int4 *>(p));
}

// Load the 8-sample block starting at sample index `blk` (absolute index into the signal
// array whose element 0 is at `base`); `blk` is such that the address is 16-byte aligned.
// Blocks that stick out of the allocation [alloc_lo, alloc_hi) are read sample by sample.
template <bool STREAM = true>
__device__ __forceinline__ Samples8 load_block8(const int16_t *base, int64_t blk, int64_t alloc_lo,
                                                int64_t alloc_hi)
{
    Samples8 s;
    if (blk >= alloc_lo && blk + 8 <= alloc_hi) {
        s.v = STREAM ? ld_stream_16(base + blk) : ld_cached_16(base + blk);
    } else {
        int w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const int64_t i = blk + e;
            const int val = (i >= alloc_lo && i < alloc_hi) ? (int)base[i] : 0;
            w[e >> 1] |= (val & 0xffff) << ((e & 1) * 16);
        }
        s.v = make_int4(w[0], w[1], w[2], w[3]);
    }
    return s;
}

// First 8-sample block (absolute sample index, may be < begin) covering `begin`, chosen so that
// the block's byte address is 16-byte aligned.
__device__ __forceinline__ int64_t aligned_block_start(const int16_t *base, int64_t begin)
{
    const uintptr_t addr = reinterpret_cast<uintptr_t>(base + begin);
    return begin - (int64_t)((addr & 15u) >> 1);
}

// Device-mode callers hand us pointers whose allocation bounds we do not know: there the first and
// last offsets of the launch delimit what may be touched (alloc_lo > alloc_hi asks for that).
__device__ __forceinline__ void resolve_bounds(const int64_t *offsets, int64_t read0, int64_t n_reads,
                                               int64_t &alloc_lo, int64_t &alloc_hi)
{
    if (alloc_lo > alloc_hi) {
        alloc_lo = offsets[read0];
        alloc_hi = offsets[read0 + n_reads];
    }
}

__device__ __forceinline__ double shfl_up_f64(double v, int delta, int width)
{
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_up_sync(SQK_FULL_MASK, lo, delta, width);
    hi = __shfl_up_sync(SQK_FULL_MASK, hi, delta, width);
    return __hiloint2double(hi, lo);
}

__device__ __forceinline__ double shfl_xor_f64(double v, int mask, int width)
{
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(SQK_FULL_MASK, lo, mask, width);
    hi = __shfl_xor_sync(SQK_FULL_MASK, hi, mask, width);
    return __hiloint2double(hi, lo);
}

__device__ __forceinline__ double shfl_idx_f64(double v, int src, int width)
{
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(SQK_FULL_MASK, lo, src, width);
    hi = __shfl_sync(SQK_FULL_MASK, hi, src, width);
    return __hiloint2double(hi, lo);
}
