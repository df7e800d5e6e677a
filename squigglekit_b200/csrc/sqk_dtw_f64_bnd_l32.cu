// float64 DTW kernel, row-block variant (boundary row in / out), 32 lanes per read: motifs of more than 1024 points
#include "sqk_dtw_launch.cuh"
cudaError_t sqk_launch_dtw_f64_bnd_l32(int K, const DtwArgs &a, int n_sms, cudaStream_t st)
{
    if (K < SQK_DTW_L32_KMIN || K > SQK_DTW_L32_KMAX) return cudaErrorInvalidValue;
    return SqkDtwBndDispatch<SQK_DTW_L32_KMIN, SQK_DTW_L32_KMAX>::go(K, a, n_sms, st);
}
