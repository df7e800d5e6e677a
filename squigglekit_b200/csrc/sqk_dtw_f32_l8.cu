// DTW kernel instantiations: float costs, 8 lane(s) per read, K = SQK_DTW_L8_KMIN..SQK_DTW_L8_KMAX rows per lane.
#include "sqk_dtw_launch.cuh"
SQK_DEFINE_DTW_LAUNCHER(float, f32, 8, SQK_DTW_L8_KMIN, SQK_DTW_L8_KMAX)
