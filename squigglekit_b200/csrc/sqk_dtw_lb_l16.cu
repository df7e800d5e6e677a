// lower-bound kernel (pass 1 of the two-pass DTW plan), 16 lane(s) per read
#include "sqk_dtw_lb_launch.cuh"
SQK_DEFINE_LB_LAUNCHER(16, SQK_DTW_L16_KMIN, SQK_DTW_L16_KMAX)
