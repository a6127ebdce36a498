// instantiations of the lane-family kernels for N in [7, 8, 9] (see lane_kernels.cuh)
#include "lane_kernels.cuh"
LANE_INSTANTIATE(7)
LANE_INSTANTIATE(8)
LANE_INSTANTIATE(9)
