// DTW kernel instantiations: float costs, 32 lane(s) per read, K = SQK_DTW_L32_KMIN..SQK_DTW_L32_KMAX rows per lane.
#include "sqk_dtw_launch.cuh"
SQK_DEFINE_DTW_LAUNCHER(float, f32, 32, SQK_DTW_L32_KMIN, SQK_DTW_L32_KMAX)
