// sqk_dtw_launch.cuh -- instantiation + launch plumbing for the DTW kernel family.
// One translation unit per (precision, lanes-per-read) so the K = rows-per-lane variants
// compile in parallel (make -j).  Each TU exports one launcher taking K at run time.
#pragma once
#include "sqk_dtw.cuh"

typedef cudaError_t (*sqk_dtw_launcher)(int K, const DtwArgs &a, int n_sms, cudaStream_t st);

template <typename T, int K, int L, bool RAGGED, bool JOBS>
static cudaError_t sqk_dtw_launch_one(const DtwArgs &a, int n_sms, cudaStream_t st)
{
    static int occ = 0;   // resident CTAs per SM for this instantiation
    if (occ == 0) {
        int o = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, sqk_dtw_kernel<T, K, L, RAGGED, JOBS>, SQK_DTW_THREADS, 0);
        if (e != cudaSuccess) return e;
        occ = o > 0 ? o : 1;
    }
    constexpr int reads_per_cta = SQK_DTW_WARPS * (32 / L);
    // JOBS: the number of work items only exists on the device; a.n_reads is its upper bound for sizing the grid
    long long want = ((long long)a.n_reads + reads_per_cta - 1) / reads_per_cta;
    long long grid = (long long)n_sms * occ;           // persistent: every CTA resident, groups pull reads
    if (want < grid) grid = want;
    if (grid < 1) grid = 1;
    sqk_dtw_kernel<T, K, L, RAGGED, JOBS><<<(unsigned)grid, SQK_DTW_THREADS, 0, st>>>(a);
    return cudaGetLastError();
}

template <typename T> struct SqkDtwHasJobs { static constexpr bool value = false; };
template <> struct SqkDtwHasJobs<double> { static constexpr bool value = true; };   // pass 2 is float64 only

template <typename T, int L, int K, int KMAX>
struct SqkDtwDispatch {
    static cudaError_t go(int k, const DtwArgs &a, int n_sms, cudaStream_t st)
    {
        if (k == K) { // motif rows fill the lanes exactly (no pass-through slots) or not
            if constexpr (SqkDtwHasJobs<T>::value) {
                if (a.jobs != nullptr)
                    return a.N == K * L ? sqk_dtw_launch_one<T, K, L, false, true>(a, n_sms, st) : sqk_dtw_launch_one<T, K, L, true, true>(a, n_sms, st);
            }
            return a.N == K * L ? sqk_dtw_launch_one<T, K, L, false, false>(a, n_sms, st) : sqk_dtw_launch_one<T, K, L, true, false>(a, n_sms, st);
        }
        if constexpr (K < KMAX) return SqkDtwDispatch<T, L, K + 1, KMAX>::go(k, a, n_sms, st);
        else return cudaErrorInvalidValue;
    }
};

#define SQK_DEFINE_DTW_LAUNCHER(T, TAG, L, KMIN, KMAX)                                              \
    cudaError_t sqk_launch_dtw_##TAG##_l##L(int K, const DtwArgs &a, int n_sms, cudaStream_t st)    \
    {                                                                                               \
        if (K < KMIN || K > KMAX) return cudaErrorInvalidValue;                                     \
        return SqkDtwDispatch<T, L, KMIN, KMAX>::go(K, a, n_sms, st);                               \
    }

// Row-block variant (motifs longer than one pass holds): float64, 32 lanes, rows cut into blocks by sqk_api.cu; one TU.
template <int K, bool RAGGED>
static cudaError_t sqk_dtw_launch_bnd(const DtwArgs &a, int n_sms, cudaStream_t st)
{
    static int occ = 0;
    if (occ == 0) {
        int o = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, sqk_dtw_kernel<double, K, 32, RAGGED, false, true>, SQK_DTW_THREADS, 0);
        if (e != cudaSuccess) return e;
        occ = o > 0 ? o : 1;
    }
    long long want = ((long long)a.n_reads + SQK_DTW_WARPS - 1) / SQK_DTW_WARPS;
    long long grid = (long long)n_sms * occ;
    if (want < grid) grid = want;
    if (grid < 1) grid = 1;
    sqk_dtw_kernel<double, K, 32, RAGGED, false, true><<<(unsigned)grid, SQK_DTW_THREADS, 0, st>>>(a);
    return cudaGetLastError();
}
template <int K, int KMAX>
struct SqkDtwBndDispatch {
    static cudaError_t go(int k, const DtwArgs &a, int n_sms, cudaStream_t st)
    {
        if (k == K) return a.N == K * 32 ? sqk_dtw_launch_bnd<K, false>(a, n_sms, st) : sqk_dtw_launch_bnd<K, true>(a, n_sms, st);
        if constexpr (K < KMAX) return SqkDtwBndDispatch<K + 1, KMAX>::go(k, a, n_sms, st);
        else return cudaErrorInvalidValue;
    }
};
cudaError_t sqk_launch_dtw_f64_bnd_l32(int K, const DtwArgs &a, int n_sms, cudaStream_t st);

// launchers defined across sqk_dtw_*.cu; [KMIN, KMAX] per lanes-per-read must match sqk_api.cu
#define SQK_DTW_L1_KMIN 1
#define SQK_DTW_L1_KMAX 4
#define SQK_DTW_L4_KMIN 2
#define SQK_DTW_L4_KMAX 20
#define SQK_DTW_L8_KMIN 2
#define SQK_DTW_L8_KMAX 20
#define SQK_DTW_L16_KMIN 2
#define SQK_DTW_L16_KMAX 20
#define SQK_DTW_L32_KMIN 2
#define SQK_DTW_L32_KMAX 32

#define SQK_DECLARE_DTW_LAUNCHERS(TAG)                                                         \
    cudaError_t sqk_launch_dtw_##TAG##_l1(int, const DtwArgs &, int, cudaStream_t);            \
    cudaError_t sqk_launch_dtw_##TAG##_l4(int, const DtwArgs &, int, cudaStream_t);            \
    cudaError_t sqk_launch_dtw_##TAG##_l8(int, const DtwArgs &, int, cudaStream_t);            \
    cudaError_t sqk_launch_dtw_##TAG##_l16(int, const DtwArgs &, int, cudaStream_t);           \
    cudaError_t sqk_launch_dtw_##TAG##_l32(int, const DtwArgs &, int, cudaStream_t);
SQK_DECLARE_DTW_LAUNCHERS(f64)

// pass 1 of the two-pass plan (sqk_dtw_lb.cuh), same (lanes, rows-per-lane) grid; defined in sqk_dtw_lb_l*.cu
struct LbArgs;
typedef cudaError_t (*sqk_lb_launcher)(int K, const LbArgs &a, int n_sms, cudaStream_t st);
cudaError_t sqk_launch_lb_l4(int, const LbArgs &, int, cudaStream_t);
cudaError_t sqk_launch_lb_l8(int, const LbArgs &, int, cudaStream_t);
cudaError_t sqk_launch_lb_l16(int, const LbArgs &, int, cudaStream_t);
cudaError_t sqk_launch_lb_l32(int, const LbArgs &, int, cudaStream_t);
