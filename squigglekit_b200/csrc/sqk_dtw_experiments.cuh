// sqk_dtw_experiments.cuh -- alternative formulations of the DTW step that were measured and NOT adopted.
// They are compiled only into the micro-benchmark (sqk_ubench.cu), whose numbers back the "alternatives
// measured and rejected" paragraph of DESIGN.md §4; libsqk.so does not contain them.
//   dtw_step2      two signal columns per step (shared shuffle / ring / argmin overhead, two dependency chains)
//   dtw_step_cost  cost-only step without start pointers (the main pass of a two-pass scheme)
#pragma once
#include "sqk_dtw.cuh"

// One DTW cell: candidates left (lf), diagonal (dg), up (u) with their start pointers; tie order of the
// reference's back-trace: diagonal, then left, then up.
template <typename T>
__device__ __forceinline__ void dtw_cell(T x, T y, T lf_c, int lf_s, T dg_c, int dg_s, T u_c, int u_s, T &nc, int &ns)
{
    const bool p = lf_c < dg_c;
    const T m1_c = p ? lf_c : dg_c; const int m1_s = p ? lf_s : dg_s;
    const bool q = u_c < m1_c;
    const T m_c = q ? u_c : m1_c;
    ns = q ? u_s : m1_s;
    nc = DtwNum<T>::step(x, y, m_c);
}

// Two signal columns per step: lane l handles columns 2(tp-l) and 2(tp-l)+1 for its K rows, updating the
// column registers in place (the old value of a row is consumed as the diagonal of the next row).  Per
// cell the work is the same as dtw_step; per step the shuffle / ring / argmin overhead is paid once for 2K
// cells and the two columns give the scheduler two interleaved dependency chains.
template <typename T, int K, int L, bool RAGGED>
__device__ __forceinline__ void dtw_step2(T (&c)[K], int (&s)[K], const T (&x)[K], const T *ring, int l, bool pass0,
                                          int tp, int n_last, T &botA_c, int &botA_s, T &botB_c, int &botB_s,
                                          T &prev_c, int &prev_s, T &best, int &best_j, int &best_s)
{
    using Num = DtwNum<T>;
    constexpr int RC = 16 * L;
    T upA_c = Num::shfl_up(botA_c, L), upB_c = Num::shfl_up(botB_c, L);
    int upA_s = __shfl_up_sync(SQK_FULL_MASK, botA_s, 1, L), upB_s = __shfl_up_sync(SQK_FULL_MASK, botB_s, 1, L);
    const int j0 = 2 * (tp - l);
    if (l == 0) { upA_c = (T)0; upA_s = j0 + 1; upB_c = (T)0; upB_s = j0 + 2; }   // virtual row: free start
    const int ri = j0 & (RC - 1);
    const T y0 = ring[ri], y1 = ring[ri + 1];
    T dg_c = prev_c; int dg_s = prev_s;             // C[row above][j0-1]
    prev_c = upB_c; prev_s = upB_s;
    T uA_c = upA_c, uB_c = upB_c; int uA_s = upA_s, uB_s = upB_s;
#pragma unroll
    for (int k = 0; k < K; k++) {
        const T old_c = c[k]; const int old_s = s[k];
        T a_c, b_c; int a_s, b_s;
        dtw_cell<T>(x[k], y0, old_c, old_s, dg_c, dg_s, uA_c, uA_s, a_c, a_s);
        if (RAGGED && k == 0 && pass0) { a_c = upA_c; a_s = upA_s; }
        dtw_cell<T>(x[k], y1, a_c, a_s, uA_c, uA_s, uB_c, uB_s, b_c, b_s);
        if (RAGGED && k == 0 && pass0) { b_c = upB_c; b_s = upB_s; }
        dg_c = old_c; dg_s = old_s;
        uA_c = a_c; uA_s = a_s; uB_c = b_c; uB_s = b_s;
        c[k] = b_c; s[k] = b_s;
    }
    botA_c = uA_c; botA_s = uA_s; botB_c = uB_c; botB_s = uB_s;
    // running first-argmin of the last row over both columns, in column order
    const int jl = 2 * (tp - (L - 1));
    const bool betA = (unsigned)jl < (unsigned)n_last && botA_c < best;
    const bool betB = (unsigned)(jl + 1) < (unsigned)n_last && botB_c < best;
    if (__any_sync(SQK_FULL_MASK, betA || betB)) {
        if (betA) { best = botA_c; best_j = jl; best_s = botA_s; }
        if ((unsigned)(jl + 1) < (unsigned)n_last && botB_c < best) { best = botB_c; best_j = jl + 1; best_s = botB_s; }
    }
}

// Cost-only variant of dtw_step: same values bit for bit (min3 is order-free), no start pointers --
// 4 instead of 6 ALU-pipe selects per cell.  Used by the two-pass scheme: this pass finds end and dist,
// a short pointer-carrying re-run from a saved column state recovers start.
template <typename T, int K, int L, bool RAGGED>
__device__ __forceinline__ void dtw_step_cost(const T (&ci)[K], T (&co)[K], const T (&x)[K], const T *ring, int l,
                                              bool pass0, int t, int n_last, T &bot_c, T &prev_up_c, T &best, int &best_j)
{
    using Num = DtwNum<T>;
    constexpr int RC = 16 * L;
    T up_c = Num::shfl_up(bot_c, L);
    if (l == 0) up_c = (T)0;
    const T y = ring[(t - l) & (RC - 1)];
    T dg_c = prev_up_c;
    prev_up_c = up_c;
    T u_c = up_c;
#pragma unroll
    for (int k = 0; k < K; k++) {
        const T lf_c = ci[k];
        const T m1_c = lf_c < dg_c ? lf_c : dg_c;
        const T m_c = u_c < m1_c ? u_c : m1_c;
        T nc = Num::step(x[k], y, m_c);
        if (RAGGED && k == 0 && pass0) nc = up_c;
        dg_c = lf_c;
        u_c = nc;
        co[k] = nc;
    }
    bot_c = u_c;
    const int j = t - (L - 1);
    const bool better = (unsigned)j < (unsigned)n_last && bot_c < best;
    if (__any_sync(SQK_FULL_MASK, better)) {
        if (better) { best = bot_c; best_j = j; }
    }
}

