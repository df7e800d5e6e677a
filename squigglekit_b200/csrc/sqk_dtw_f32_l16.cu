// DTW kernel instantiations: float costs, 16 lane(s) per read, K = SQK_DTW_L16_KMIN..SQK_DTW_L16_KMAX rows per lane.
#include "sqk_dtw_launch.cuh"
SQK_DEFINE_DTW_LAUNCHER(float, f32, 16, SQK_DTW_L16_KMIN, SQK_DTW_L16_KMAX)
