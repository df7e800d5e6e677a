// lower-bound kernel (pass 1 of the two-pass DTW plan), 8 lane(s) per read
#include "sqk_dtw_lb_launch.cuh"
SQK_DEFINE_LB_LAUNCHER(8, SQK_DTW_L8_KMIN, SQK_DTW_L8_KMAX)
