// instantiations of the lane-family kernels for N in [10, 11] (see lane_kernels.cuh)
#include "lane_kernels.cuh"
LANE_INSTANTIATE(10)
LANE_INSTANTIATE(11)
