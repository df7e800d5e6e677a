// sqk_dtw_lb_launch.cuh -- instantiation + launch plumbing for the lower-bound kernel (one TU per lanes-per-read).
#pragma once
#include "sqk_dtw_lb.cuh"
#include "sqk_dtw_launch.cuh"

template <int K, int L, bool RAGGED, int COLS>
static cudaError_t sqk_lb_launch_one(const LbArgs &a, int n_sms, cudaStream_t st)
{
    static int occ = 0;
    if (occ == 0) {
        int o = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, sqk_dtw_lb_kernel<K, L, RAGGED, COLS>, SQK_LB_THREADS, 0);
        if (e != cudaSuccess) return e;
        occ = o > 0 ? o : 1;
    }
    constexpr int reads_per_cta = SQK_LB_WARPS * (32 / L);
    long long want = ((long long)a.n_reads + reads_per_cta - 1) / reads_per_cta;
    long long grid = (long long)n_sms * occ;           // persistent: groups pull reads from the queue
    if (want < grid) grid = want;
    if (grid < 1) grid = 1;
    sqk_dtw_lb_kernel<K, L, RAGGED, COLS><<<(unsigned)grid, SQK_LB_THREADS, 0, st>>>(a);
    return cudaGetLastError();
}

template <int L, int K, int KMAX>
struct SqkLbDispatch {
    static cudaError_t go(int k, const LbArgs &a, int n_sms, cudaStream_t st)
    {
        if (k == K) {
            // columns per wavefront step: four when a lane holds few rows (the per-step work is spread over 4K cells); two
            // when the rows already amortise it or the ring of 32 L entries would not pay (a.cols: 0 = this rule)
            const int cols = a.cols ? a.cols : (K <= 12 && L >= 4 ? 4 : 2);
            if (cols == 4)
                return a.N == K * L ? sqk_lb_launch_one<K, L, false, 4>(a, n_sms, st) : sqk_lb_launch_one<K, L, true, 4>(a, n_sms, st);
            return a.N == K * L ? sqk_lb_launch_one<K, L, false, 2>(a, n_sms, st) : sqk_lb_launch_one<K, L, true, 2>(a, n_sms, st);
        }
        if constexpr (K < KMAX) return SqkLbDispatch<L, K + 1, KMAX>::go(k, a, n_sms, st);
        else return cudaErrorInvalidValue;
    }
};

#define SQK_DEFINE_LB_LAUNCHER(L, KMIN, KMAX)                                                      \
    cudaError_t sqk_launch_lb_l##L(int K, const LbArgs &a, int n_sms, cudaStream_t st)              \
    {                                                                                               \
        if (K < KMIN || K > KMAX) return cudaErrorInvalidValue;                                     \
        return SqkLbDispatch<L, KMIN, KMAX>::go(K, a, n_sms, st);                                   \
    }
