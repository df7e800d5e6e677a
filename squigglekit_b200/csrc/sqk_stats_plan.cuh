// sqk_stats_plan.cuh -- numpy's pairwise-summation split rule as scalar functions shared by the stats kernel
// (sqk_stats.cuh) and the CPU test harness (tests/plan_harness.cpp).
//
// numpy (DOUBLE_pairwise_sum): n <= 128 is a leaf (eight strided accumulators); otherwise the left half has
// floor(n/2) rounded down to a multiple of 8 elements, the right half the rest, recursively.  The right child is
// never the smaller one, so the right-most path has the depth D of the whole tree; leaves above level D exist
// ("early leaves").  Slot j in [0, 2^D) addresses the node reached by reading j's bits from the top.
#pragma once

#if defined(__CUDACC__)
#define SQK_SP_HD __host__ __device__ __forceinline__
#else
#define SQK_SP_HD inline
#endif

// depth of the pairwise tree over n elements (0: a single leaf)
SQK_SP_HD int sqk_tree_depth(int n)
{
    int depth = 0;
    for (int len = n; len > 128; depth++) { int half = len / 2; half -= half % 8; len -= half; }
    return depth;
}

// The leaf that slot j of a depth-`depth` tree over n elements stands for: elements [off, off+len).  Returns false for
// the slots under an early leaf other than its left-most one (they contribute 0.0).
SQK_SP_HD bool sqk_tree_leaf(int n, int depth, int j, int *off_out, int *len_out)
{
    int off = 0, len = n, d = 0;
    while (len > 128) {
        int half = len / 2; half -= half % 8;
        if ((j >> (depth - 1 - d)) & 1) { off += half; len -= half; } else len = half;
        d++;
    }
    *off_out = off; *len_out = len;
    return !(d < depth && (j & ((1 << (depth - d)) - 1)) != 0);
}

// The same two functions with a fixed trip count (n <= 8192: at most 7 levels) and no data-dependent branch, for the
// warp-per-read kernel (sqk_stats3.cuh); checked against the loops above for every n <= 8192 (tests/test_plan_cpu.py).
SQK_SP_HD int sqk_tree_depth7(int n)
{
    int depth = 0, len = n;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int lv = 0; lv < 7; lv++) {
        const bool go = len > 128;
        const int half = (len >> 1) & ~7;
        len = go ? len - half : len;
        depth += go ? 1 : 0;
    }
    return depth;
}

SQK_SP_HD bool sqk_tree_leaf7(int n, int depth, int j, int *off_out, int *len_out)
{
    int off = 0, len = n, d = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int lv = 0; lv < 7; lv++) {
        const bool go = len > 128;
        const int half = (len >> 1) & ~7;
        const bool right = ((j >> ((depth - 1 - lv) & 31)) & 1) != 0;     // (only looked at when go: then lv < depth)
        off += (go && right) ? half : 0;
        len = go ? (right ? len - half : half) : len;
        d += go ? 1 : 0;
    }
    *off_out = off; *len_out = len;
    return !(d < depth && (j & ((1 << (depth - d)) - 1)) != 0);
}
