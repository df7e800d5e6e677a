// DTW kernel instantiations: double costs, 1 lane(s) per read, K = SQK_DTW_L1_KMIN..SQK_DTW_L1_KMAX rows per lane.
#include "sqk_dtw_launch.cuh"
SQK_DEFINE_DTW_LAUNCHER(double, f64, 1, SQK_DTW_L1_KMIN, SQK_DTW_L1_KMAX)
