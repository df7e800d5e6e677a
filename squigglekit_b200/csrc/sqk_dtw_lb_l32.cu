// lower-bound kernel (pass 1 of the two-pass DTW plan), 32 lane(s) per read
#include "sqk_dtw_lb_launch.cuh"
SQK_DEFINE_LB_LAUNCHER(32, SQK_DTW_L32_KMIN, SQK_DTW_L32_KMAX)
