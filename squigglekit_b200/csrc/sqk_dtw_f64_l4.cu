// DTW kernel instantiations: double costs, 4 lane(s) per read, K = SQK_DTW_L4_KMIN..SQK_DTW_L4_KMAX rows per lane.
#include "sqk_dtw_launch.cuh"
SQK_DEFINE_DTW_LAUNCHER(double, f64, 4, SQK_DTW_L4_KMIN, SQK_DTW_L4_KMAX)
