// lower-bound kernel (pass 1 of the two-pass DTW plan), 4 lane(s) per read
#include "sqk_dtw_lb_launch.cuh"
SQK_DEFINE_LB_LAUNCHER(4, SQK_DTW_L4_KMIN, SQK_DTW_L4_KMAX)
