// instantiations of the lane-family kernels for N in [12, 13] (see lane_kernels.cuh)
#include "lane_kernels.cuh"
LANE_INSTANTIATE(12)
LANE_INSTANTIATE(13)
